# round 2, call T: sw_setup loop reverted to the simple form, warp-path deferred reads listed instead of scanned
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2t_tests.log | cut -c1-700
timeout 900 python bench.py --hbm-transcripts 0 --steps 10 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r2t_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc=$?"
