# round 2, call AC: file-level wall clock with the own inflate (16 host threads), stream / frontend GPU tests
set -x
mkdir -p gpurun_out
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 24000000 > gpurun_out/r2ac_file.json 2> gpurun_out/r2ac_file.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2ac_file.err | tail -2 | cut -c1-420; cat gpurun_out/r2ac_file.json
NB200_ZLIB_INFLATE=1 NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 24000000 --skip-report > gpurun_out/r2ac_file_zlib.json 2> gpurun_out/r2ac_file_zlib.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2ac_file_zlib.err | tail -1 | cut -c1-420
timeout 600 python -m pytest tests/test_stream_gpu.py tests/test_frontend.py tests/test_fast_inflate.py -m "gpu or not gpu" -q 2>&1 | tail -2
