#!/bin/bash
# development: rebuild the library, stop if that fails, then run the given command on a B200 under an inner timeout
# (a kernel that deadlocks must not eat the GPU budget)
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; echo BUILD FAILED; exit 1; }
printf '%s\n' "$*" > tools/.gpu_cmd.sh
/usr/local/graft/bin/gpurun --timeout ${GPU_TIMEOUT:-600} -- "timeout ${CMD_TIMEOUT:-240} bash tools/.gpu_cmd.sh"
