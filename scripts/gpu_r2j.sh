# round 2, call J: state of HEAD for profiles/ — tests, default bench (cfg2 + HBM pass), reference arm, launch list,
# ncu --set full of every hot kernel, file-level wall clock with the post-walker trace
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2j_tests.log | cut -c1-800
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r2j_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"
NB200_BENCH_READS=2000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"probe_kernel|call_fast|call_slow|window_hash|dedupe_kernel|sw_kernel|call_deferred|umi_kernel|group_sort" -s 27 -c 10 -o gpurun_out/r2j_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2j_ncufull.log 2>&1; echo "ncufull rc=$?"
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 12000000 > gpurun_out/r2j_file.json 2> gpurun_out/r2j_file.err; echo "file rc=$?"
grep "pipeline" gpurun_out/r2j_file.err | tail -4; cat gpurun_out/r2j_file.json
