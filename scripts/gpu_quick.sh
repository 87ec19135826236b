# quick check: parity tests of the align path + one short bench line (kernel times)
timeout 600 python -m pytest tests/test_align_parity.py tests/test_align_edge_gpu.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f M/s  step %.2f ms  e2e %.1f M/s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6), d['kernels_ms_per_step'], d['e2e'])
"
