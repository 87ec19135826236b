# round 2, call AF: final state — GPU suite, smoke, default bench, launch list and ncu --set full of HEAD
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2af_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2af_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2af_bench.err; cut -c1-200 gpurun_out/r2af_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2af_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2af_ncu.log 2>&1; echo "ncu rc=$?"
NB200_BENCH_READS=2000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"probe_kernel|call_fast|sw_setup|window_hash|dedupe_kernel|sw_kernel|call_deferred_thread|umi_simple|umi_general|group_sort" -s 30 -c 11 -o gpurun_out/r2af_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2af_ncufull.log 2>&1; echo "ncufull rc=$?"
