# round 2, call O: --trim (MaxInfo) in the file pipeline, random-shape A6 tests, launch list after the call_fast change
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/r2o_tests.log | cut -c1-700
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python scripts/ref_diff.py --aligner nimble_b200/aligner --reads 20000 > gpurun_out/r2o_refdiff_self.json 2> gpurun_out/r2o_refdiff_self.err; echo "refdiff rc=$?"
tail -5 gpurun_out/r2o_refdiff_self.json
