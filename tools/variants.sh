#!/bin/bash
# development: build several variants of the library (one per NVCC_EXTRA string), time each on one GPU box
# usage: tools/variants.sh name1:"-DX" name2:"-DY -DZ" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
names=()
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  NVCC_EXTRA="$flags" python __graft_entry__.py > /tmp/build_$name.log 2>&1 || { tail -20 /tmp/build_$name.log; echo BUILD FAILED $name; exit 1; }
  cp bionumpy_b200/_lib/libbnpk.so gpurun_variants/libbnpk_$name.so
  names+=("$name")
done
python __graft_entry__.py > /tmp/build.log 2>&1
{
  echo "export TC_FIRST=1"
  echo "cp bionumpy_b200/_lib/libbnpk.so /tmp/libbnpk_orig.so"
  for round in 1 2; do
    echo "cp /tmp/libbnpk_orig.so bionumpy_b200/_lib/libbnpk.so; echo base: \$(python tools/time_count.py ${TC_ARGS:-} | tail -1)"
    for n in "${names[@]}" ${EXTRA_VARIANTS:-}; do
      echo "cp gpurun_variants/libbnpk_$n.so bionumpy_b200/_lib/libbnpk.so; echo $n: \$(timeout 60 python tools/time_count.py ${TC_ARGS:-} | tail -1)"
    done
  done
  echo "cp /tmp/libbnpk_orig.so bionumpy_b200/_lib/libbnpk.so"
} > tools/.gpu_cmd.sh
/usr/local/graft/bin/gpurun --timeout ${GPU_TIMEOUT:-600} -- "timeout ${CMD_TIMEOUT:-280} bash tools/.gpu_cmd.sh"
python - <<'PY'
import json; d=json.load(open('gpurun_out/.last_call.json')); print(d.get('status')); print(d.get('stdout_tail','')[-1500:])
PY
