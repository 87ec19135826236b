# One GPU box: full GPU test-suite, default bench (with CPU baseline), reference arm, ncu launch list of the
# bench command and one ncu --set full capture of the hot kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s4_ref.json 2> gpurun_out/s4_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s4_ncu.log 2>&1; echo "ncu rc=$?"
NB200_BENCH_READS=2000000 ncu --set full --clock-control none --import-source on -k regex:"probe_kernel|window_hash_kernel|dedupe_kernel|sw_kernel|call_deferred_kernel|umi_kernel|group_sort_kernel" -s 18 -c 7 -o gpurun_out/s4_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s4_ncufull.log 2>&1; echo "ncufull rc=$?"
tail -3 gpurun_out/s4_tests.log; cut -c1-400 gpurun_out/s4_bench.json; cut -c1-300 gpurun_out/s4_ref.json
