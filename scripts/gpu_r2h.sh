# round 2, call H: call_fast specialised for single-end, streaming pipeline tuned (small slabs, LUT packing)
set -x
mkdir -p gpurun_out
nproc
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2h_tests.log | cut -c1-800
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2h_bench.err
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 8000000 > gpurun_out/r2h_file.json 2> gpurun_out/r2h_file.err; echo "file rc=$?"
grep -v "batch\|agg start" gpurun_out/r2h_file.err | tail -12; cat gpurun_out/r2h_file.json
