#!/bin/bash
# development: rebuild the library, stop if that fails, then run the given command on a B200
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; echo BUILD FAILED; exit 1; }
/usr/local/graft/bin/gpurun --timeout ${GPU_TIMEOUT:-900} -- "$@"
