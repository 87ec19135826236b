# round 2, call AD: file-level wall clock after the reader-thread fix (16 host threads), 2 sizes
set -x
mkdir -p gpurun_out
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 24000000 > gpurun_out/r2ad_file.json 2> gpurun_out/r2ad_file.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2ad_file.err | tail -2 | cut -c1-420; cat gpurun_out/r2ad_file.json
timeout 600 python -m pytest tests/test_stream_gpu.py tests/test_frontend.py -m "gpu" -q 2>&1 | tail -2
