# round 2, call L: slab pool (pageable + background pinning + cache), parallel TSV parser, A6 launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2l_tests.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_report_launches.csv python bench.py --workload report --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 12000000 > gpurun_out/r2l_file.json 2> gpurun_out/r2l_file.err; echo "file rc=$?"
grep "pipeline" gpurun_out/r2l_file.err | tail -4; cat gpurun_out/r2l_file.json
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2l_bench.err
