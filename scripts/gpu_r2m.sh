# round 2, call M: emit_items lane-per-candidate, umi_general local arrays, run_heads block atomics
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2m_tests.log | cut -c1-600
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2m_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_report_launches.csv python bench.py --workload report --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload report --steps 10 --no-cpu-baseline > gpurun_out/r2m_report.json 2> gpurun_out/r2m_report.err; echo "report rc=$?"
cut -c1-400 gpurun_out/r2m_report.json
