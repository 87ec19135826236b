# round 2, call U: profile set of HEAD — tests, default bench (cfg2 + HBM pass), reference arm, launch list, ncu --set full of
# every hot kernel, report stage, one GPU's share of cfg4, file-level wall clock
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2u_tests.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2u_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_ref.json 2> gpurun_out/r2u_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2u_ncu.log 2>&1; echo "ncu rc=$?"
NB200_BENCH_READS=2000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"probe_kernel|call_fast|sw_setup|window_hash|dedupe_kernel|sw_kernel|call_deferred_thread|umi_simple|umi_general|group_sort" -s 30 -c 11 -o gpurun_out/r2u_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2u_ncufull.log 2>&1; echo "ncufull rc=$?"
timeout 600 python bench.py --workload report --steps 10 > gpurun_out/r2u_report.json 2> gpurun_out/r2u_report.err; echo "report rc=$?"
timeout 900 python bench.py --workload cfg4 --reads 10000000 --steps 5 > gpurun_out/r2u_cfg4.json 2> gpurun_out/r2u_cfg4.err; echo "cfg4 rc=$?"
tail -2 gpurun_out/r2u_cfg4.err
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 12000000 > gpurun_out/r2u_file.json 2> gpurun_out/r2u_file.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2u_file.err | tail -2 | cut -c1-400; cat gpurun_out/r2u_file.json
