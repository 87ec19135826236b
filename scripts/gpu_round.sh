set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s2_ref.json 2> gpurun_out/s2_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s2_ncu.log 2>&1; echo "ncu rc=$?"
NB200_BENCH_READS=2000000 ncu --set full --clock-control none --import-source on -k regex:"probe_kernel|sw_kernel|call_deferred_kernel|umi_kernel" -s 12 -c 4 -o gpurun_out/s2_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s2_ncufull.log 2>&1; echo "ncufull rc=$?"
tail -3 gpurun_out/s2_tests.log; cat gpurun_out/s2_bench.json | cut -c1-600; cat gpurun_out/s2_ref.json | cut -c1-400
