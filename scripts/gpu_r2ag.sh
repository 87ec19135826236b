# round 2, call AG: report path with the faster TSV read
set -x
timeout 300 python -m pytest tests/test_frontend.py tests/test_a6_pandas_golden.py tests/test_a6_reference_vectors.py tests/test_barcode_gpu.py -m gpu -q 2>&1 | tail -2
timeout 300 python scripts/file_bench.py --reads 12000000 2>/dev/null | tail -1
