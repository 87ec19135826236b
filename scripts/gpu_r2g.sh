# round 2, call G: staged stores in call_fast, streaming file pipeline behind nb200_align_files
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=6 > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2g_tests.log | cut -c1-800
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2g_bench.err
NB200_TRACE=1 timeout 600 python scripts/file_bench.py --reads 4000000 > gpurun_out/r2g_file.json 2> gpurun_out/r2g_file.err; echo "file rc=$?"
grep -v "batch\|agg start" gpurun_out/r2g_file.err | tail -12; cat gpurun_out/r2g_file.json
