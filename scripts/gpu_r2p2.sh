# round 2, call P (2 GPUs, small): rehearsal of the 8-GPU script — file-level cfg4 on 2 GPUs vs 1 GPU, bench cfg4 on 2 ranks
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3; nproc
NB200_TRACE=1 timeout 900 python scripts/file_bench.py --workload cfg4 --reads 8000000 --gpus 2 --compare-1gpu > gpurun_out/r2p2_file.json 2> gpurun_out/r2p2_file.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2p2_file.err | tail -3 | cut -c1-400; cat gpurun_out/r2p2_file.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --reads 2000000 --steps 3 > gpurun_out/r2p2_cfg4.json 2> gpurun_out/r2p2_cfg4.err; echo "cfg4 rc=$?"
grep "resident arm\|parity" gpurun_out/r2p2_cfg4.err | cut -c1-300; cut -c1-300 gpurun_out/r2p2_cfg4.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 > gpurun_out/r2p2_bench2.json 2> gpurun_out/r2p2_bench2.err; echo "bench2 rc=$?"
grep "resident arm" gpurun_out/r2p2_bench2.err; cut -c1-300 gpurun_out/r2p2_bench2.json
