# round 2, call Y: first-batch ramp of the e2e path (NB200_RAMP = divisors of the first batches)
set -x
mkdir -p gpurun_out
for R in "8,2" "8,3,1" "16,6,2" "6,2" "12,4,1" "4,2"; do
  NB200_RAMP=$R timeout 300 python bench.py --hbm-transcripts 0 --steps 10 --no-cpu-baseline > gpurun_out/r2y_$R.json 2> gpurun_out/r2y_$R.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2y_$R.json"))
print("ramp $R: e2e %.1f M reads/s (%.2f ms wall, device %.2f ms) resident %.1f M" % (d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["device_ms"]["total_ms"], d["value"]/1e6))
PY
done
