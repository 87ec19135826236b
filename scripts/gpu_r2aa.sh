# round 2, call AA: run-based BAM walk + paired trim case on the GPU, file-level wall clock
set -x
mkdir -p gpurun_out
nproc
timeout 900 python -m pytest tests/test_stream_gpu.py tests/test_trim.py tests/test_frontend.py tests/test_barcode_gpu.py -m gpu -q > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2aa_tests.log | cut -c1-300
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 24000000 --skip-report > gpurun_out/r2aa_file.json 2> gpurun_out/r2aa_file.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2aa_file.err | tail -2 | cut -c1-420; cat gpurun_out/r2aa_file.json
