# bench several builds of the library (launch-bound / block-size variants) back to back
mkdir -p gpurun_out
for v in variants/*.so; do
  n=$(basename $v .so)
  NB200_LIB=$PWD/$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --hbm-transcripts 0 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open('gpurun_out/var_%s.json'%n))
    k=d['kernels_ms_per_step']
    print(n, 'value %.1f e2e %.1f | probe %.2f sw %.2f call %.2f agg %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, k['probe'],k['sw'],k['call'],k['agg']))
except Exception as e: print(n,'failed',e)
PY
done
