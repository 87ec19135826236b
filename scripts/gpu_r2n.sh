# round 2, call N: max_nf tracking moved into the aggregation (call_fast hot spot), then configs[4] at FULL scale
# (200 k transcripts, k = 31, 401 M k-mers, 12.8 GB table) on one GPU
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2n_tests.log | cut -c1-600
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2n_bench.err
timeout 2400 python bench.py --workload cfg5 --transcripts 200000 --reads 10000000 --steps 5 --parity-reads 20000 --cpu-sample 200000 > gpurun_out/r2n_cfg5_full.json 2> gpurun_out/r2n_cfg5_full.err; echo "cfg5 rc=$?"
tail -6 gpurun_out/r2n_cfg5_full.err; cut -c1-300 gpurun_out/r2n_cfg5_full.json
