# round 2, call Q: sw_setup_kernel (thread per read) replaces the warp-per-read SW setup
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2q_tests.log | cut -c1-700
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2q_bench.err
NB200_BENCH_READS=2000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sw_setup|call_fast|call_deferred|dedupe_kernel|window_hash" -s 15 -c 5 -o gpurun_out/r2q_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2q_ncufull.log 2>&1; echo "ncufull rc=$?"
