# round 2, call F: probe kernel split (warp-per-read probe, thread-per-read calling, warp-per-read slow path)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2f_tests.log | cut -c1-600
timeout 900 python bench.py --hbm-transcripts 0 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r2f_bench.err
NB200_BENCH_READS=2000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"probe_kernel|call_fast|call_slow" -s 9 -c 3 -o gpurun_out/r2f_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2f_ncufull.log 2>&1; echo "ncufull rc=$?"
