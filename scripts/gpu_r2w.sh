# round 2, call W: call_fast list atomics hidden behind the calling work, unrolled candidate scores in the deferred call, async table fetch
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2w_tests.log | cut -c1-500
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2w_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
