set -x
mkdir -p gpurun_out
python bench.py --workload fastq-to-bam > gpurun_out/cb_bench.json 2> gpurun_out/cb_bench.err; echo "rc=$?"
python bench.py --workload fastq-to-bam --whitelist 6794880 --no-cpu-baseline > gpurun_out/cb_bench_v3.json 2> gpurun_out/cb_bench_v3.err; echo "rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/cb_launches.csv python bench.py --workload fastq-to-bam --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/cb_ncu.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"cb_exact_kernel|cb_hamming_kernel" -s 6 -c 2 -o gpurun_out/cb_full python bench.py --workload fastq-to-bam --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/cb_ncufull.log 2>&1; echo "ncufull rc=$?"
cat gpurun_out/cb_bench.json; tail -5 gpurun_out/cb_bench.err; cat gpurun_out/cb_bench_v3.json
