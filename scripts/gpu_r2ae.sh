# round 2, call AE: 100 M-read BAM (cfg4 shape) through one GPU with the final reader; peak RSS of the aligner process
set -x
mkdir -p gpurun_out
nproc
NB200_TRACE=1 timeout 1500 python scripts/file_bench.py --workload cfg4 --reads 100000000 --skip-report --rss > gpurun_out/r2ae_file100M.json 2> gpurun_out/r2ae_file100M.err; echo "file rc=$?"
grep "pipeline:" gpurun_out/r2ae_file100M.err | tail -3 | cut -c1-420; cat gpurun_out/r2ae_file100M.json
