import sys, numpy as np
sys.path.insert(0, ".")
import nimble_b200
from nimble_b200 import synth
eng = nimble_b200.Engine(0)
lib, codes = synth.allele_family_library(n_founders=40, alleles_per_founder=50, length=1098, snps_mean=15.0, seed=1)
lg = eng.load_library(lib)
tup = lambda t: (t.cell.tobytes(), t.count.tobytes(), t.feat_off.tobytes(), t.feat_ids.tobytes(), t.dropped_empty, t.n_called)
rng = np.random.default_rng(1)
bad = 0
for it in range(12):
    n = int(rng.integers(1, 7_000_000))
    r1, truth = synth.sample_reads(codes, n, read_len=90, seed=100 + it)
    key = synth.barcodes_10x(n, n_cells=5000, seed=200 + it, truth=truth)
    p = eng.pack(r1, pinned=(it % 2 == 0))
    eng.set_overlap(False); ref = tup(eng.align(lg, p, key=key))
    eng.set_overlap(True)
    for rep in range(4):
        got = tup(eng.align(lg, p, key=key))
        if got != ref: bad += 1; print("MISMATCH iteration", it, "rep", rep, "n", n)
    eng.upload(p, key=key)
    for rep in range(3):
        got = tup(eng.align_resident(lg))
        if got != ref: bad += 1; print("MISMATCH resident iteration", it, "rep", rep, "n", n)
    print("iter", it, "n", n, "ok" if not bad else "BAD", flush=True)
print("bad", bad)
