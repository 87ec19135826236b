# round 2, call Z: peak RSS of the aligner process at two input sizes (bounded slab pool), GPU suite on the final code
set -x
mkdir -p gpurun_out
timeout 900 python scripts/file_bench.py --reads 6000000 --rss --skip-report > gpurun_out/r2z_rss_6M.json 2> gpurun_out/r2z_rss_6M.err; echo "rc=$?"; cat gpurun_out/r2z_rss_6M.json
timeout 1200 python scripts/file_bench.py --reads 48000000 --rss --skip-report > gpurun_out/r2z_rss_48M.json 2> gpurun_out/r2z_rss_48M.err; echo "rc=$?"; cat gpurun_out/r2z_rss_48M.json
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2z_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
