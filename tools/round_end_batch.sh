set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for tool in memcheck racecheck synccheck; do
  echo "$tool: $(timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E 'sanitize target ok|SUMMARY' | tr '\n' ' ')"
done > gpurun_out/sanitize_r01f.txt 2>&1
cat gpurun_out/sanitize_r01f.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_r01_f.err; tail -c 2500 gpurun_out/bench_r01_f.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 3 --no-extra --e2e-steps 1 > gpurun_out/bench_under_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_tma_kernel -s 3 -c 1 -o gpurun_out/prof_r01f python tools/ncu_target.py 10000000 16384 2>&1 | tail -1
