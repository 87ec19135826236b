# the other named workload shapes, at a scale that fits a few minutes
mkdir -p gpurun_out
python bench.py --workload cfg3 --reads 5000000 --cpu-sample 500000 --steps 3 > gpurun_out/cfg3.json 2> gpurun_out/cfg3.err
python bench.py --workload cfg4 --reads 10000000 --cpu-sample 500000 --steps 3 > gpurun_out/cfg4.json 2> gpurun_out/cfg4.err
python bench.py --workload cfg5 --transcripts 50000 --reads 10000000 --cpu-sample 500000 --steps 3 > gpurun_out/cfg5.json 2> gpurun_out/cfg5.err
python - <<'PY'
import json
for f in ("cfg3", "cfg4", "cfg5"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 1),
          {k: round(v, 2) for k, v in d["kernels_ms_per_step"].items() if k != "note"}, round(d["cpu_baseline"]["value"] / 1e6, 3),
          round(d["roofline"]["frac"], 3), round(d["roofline"]["random_access"]["frac_of_hbm_random"], 3))
PY
