# round 2, call X (8 GPUs): cfg2 scaling line with the asynchronous deferred fetch (D2H started before the size exchange)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 > gpurun_out/r2x_bench8.json 2> gpurun_out/r2x_bench8.err; echo "bench8 rc=$?"
grep "resident arm" gpurun_out/r2x_bench8.err | head -8; cut -c1-250 gpurun_out/r2x_bench8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 > gpurun_out/r2x_bench2.json 2> gpurun_out/r2x_bench2.err; echo "bench2 rc=$?"
grep "resident arm" gpurun_out/r2x_bench2.err | head -2; cut -c1-250 gpurun_out/r2x_bench2.json
