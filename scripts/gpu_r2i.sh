# round 2, call I (2 GPUs): streaming pipeline tests incl. multi-GPU, file bench at 1 and 2 GPUs, torchrun bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 900 python -m pytest tests/test_stream_gpu.py tests/test_frontend.py -m gpu -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2i_tests.log | cut -c1-600
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 12000000 > gpurun_out/r2i_file1.json 2> gpurun_out/r2i_file1.err; echo "file1 rc=$?"
grep "pipeline:" gpurun_out/r2i_file1.err | tail -2; cat gpurun_out/r2i_file1.json
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --reads 12000000 --gpus 2 > gpurun_out/r2i_file2.json 2> gpurun_out/r2i_file2.err; echo "file2 rc=$?"
grep "pipeline:" gpurun_out/r2i_file2.err | tail -2; cat gpurun_out/r2i_file2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2i_bench2.json 2> gpurun_out/r2i_bench2.err; echo "bench2 rc=$?"
grep "resident arm" gpurun_out/r2i_bench2.err; cut -c1-300 gpurun_out/r2i_bench2.json
