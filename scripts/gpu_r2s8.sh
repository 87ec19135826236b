# round 2, call S (8 GPUs): cfg2 scaling line, configs[3] at its stated scale (200 M reads over 8 GPUs, combined library),
# file-level 8-GPU run of the cfg4 shape (100 M reads) against the 1-GPU output
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2; df -h /tmp | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 > gpurun_out/r2s_bench8.json 2> gpurun_out/r2s_bench8.err; echo "bench8 rc=$?"
grep "resident arm" gpurun_out/r2s_bench8.err | head -8; cut -c1-250 gpurun_out/r2s_bench8.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload cfg4 --reads 25000000 --steps 5 > gpurun_out/r2s_cfg4_200M.json 2> gpurun_out/r2s_cfg4_200M.err; echo "cfg4 rc=$?"
grep "resident arm\|parity" gpurun_out/r2s_cfg4_200M.err | head -10 | cut -c1-300; cut -c1-250 gpurun_out/r2s_cfg4_200M.json
READS=100000000
AVAIL=$(df --output=avail -BG /tmp | tail -1 | tr -dc 0-9)
if [ "$AVAIL" -lt 30 ]; then READS=40000000; fi
NB200_TRACE=1 timeout 1800 python scripts/file_bench.py --workload cfg4 --reads $READS --gpus 8 --compare-1gpu --skip-report > gpurun_out/r2s_file8.json 2> gpurun_out/r2s_file8.err; echo "file8 rc=$?"
grep "pipeline:\|wrote\|GPUs" gpurun_out/r2s_file8.err | cut -c1-420 | tail -8; cat gpurun_out/r2s_file8.json
