# round 2, call R: thread-per-read deferred call (thread_call shared with call_fast), block-level list atomics, chunked item loads
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2r_tests.log | cut -c1-700
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2r_bench.err
timeout 900 python bench.py --workload cfg3 --reads 5000000 --steps 5 > gpurun_out/r2r_cfg3.json 2> gpurun_out/r2r_cfg3.err; echo "cfg3 rc=$?"
tail -2 gpurun_out/r2r_cfg3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc=$?"
