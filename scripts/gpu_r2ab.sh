# round 2, call AB: file-level wall clock with the wider inflate window (16 host threads)
set -x
mkdir -p gpurun_out
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 24000000 > gpurun_out/r2ab_file.json 2> gpurun_out/r2ab_file.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2ab_file.err | tail -2 | cut -c1-420; cat gpurun_out/r2ab_file.json
timeout 600 python -m pytest tests/test_stream_gpu.py -m gpu -q 2>&1 | tail -2
