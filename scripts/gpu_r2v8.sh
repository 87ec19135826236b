# round 2, call V (8 GPUs): deferred fetch — the count-table D2H runs beside the NCCL gather
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_align_edge_gpu.py -m gpu -q -k "deferred or compact" > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2v_tests.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 > gpurun_out/r2v_bench8.json 2> gpurun_out/r2v_bench8.err; echo "bench8 rc=$?"
grep "resident arm" gpurun_out/r2v_bench8.err | head -8; cut -c1-250 gpurun_out/r2v_bench8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 10 > gpurun_out/r2v_bench4.json 2> gpurun_out/r2v_bench4.err; echo "bench4 rc=$?"
grep "resident arm" gpurun_out/r2v_bench4.err | head -4; cut -c1-250 gpurun_out/r2v_bench4.json
