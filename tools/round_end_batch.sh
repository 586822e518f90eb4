# round-end evidence in one gpurun call: parity tests, smoke, compute-sanitizer, the bench line, ncu launch list, ncu --set full
# of the dominant kernel.  Everything is written under gpurun_out/ and summarised into profiles/ on the CPU box afterwards.
set -x
R=${R:-r02}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for tool in memcheck racecheck synccheck; do
  echo "$tool: $(timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E 'sanitize target ok|SUMMARY' | tr '\n' ' ')"
done > gpurun_out/sanitize_$R.txt 2>&1
cat gpurun_out/sanitize_$R.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -c 1500 gpurun_out/bench_$R.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-extra --e2e-steps 1 > gpurun_out/bench_under_ncu_$R.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_ws_kernel -s 3 -c 1 -o gpurun_out/prof_$R python tools/ncu_target.py 10000000 16384 2>&1 | tail -1
