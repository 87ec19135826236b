# round 2, call A: new parity tests + default bench (parity slice, DPX peak, HBM pass) + reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2a_tests.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"
cut -c1-600 gpurun_out/r2a_ref.json
