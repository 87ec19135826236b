# round 2, call B: probe v2 (bucketed cuckoo table, super-round probe) — parity suite, bench, ncu capture
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2b_tests.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -8 gpurun_out/r2b_bench.err
NB200_BENCH_READS=2000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"probe_kernel|sw_kernel|call_deferred_kernel|umi_kernel" -s 12 -c 4 -o gpurun_out/r2b_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2b_ncufull.log 2>&1; echo "ncufull rc=$?"
