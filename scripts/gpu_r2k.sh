# round 2, call K: compact wire form + A6 rewrite (simple/general UMI kernels, list ranks, one stage-2 sort)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r2k_tests.log | cut -c1-600
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r2k_bench.err
timeout 600 python bench.py --workload report --steps 10 > gpurun_out/r2k_report.json 2> gpurun_out/r2k_report.err; echo "report rc=$?"
cut -c1-700 gpurun_out/r2k_report.json
