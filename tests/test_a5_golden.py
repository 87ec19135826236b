"""A5 (fastq-to-bam barcode stage): both oracle restatements against fixtures produced by the
REFERENCE's own functions (tests/golden/make_a5_golden.py -> tests/golden/a5_cases.json), and the
C restatement against the Python one on larger seeded inputs.  CPU only."""
import json
import os

import numpy as np
import pytest

from nimble_b200 import synth
from oracle import a5_py as A
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "a5_cases.json")


def load_cases():
    with open(GOLD) as f:
        return json.load(f)["cases"]


CASES = load_cases()


def quals(text):
    return [ord(c) - 33 for c in text]


def case_arrays(c):
    """What process_pair hands to correct_cell_barcode: (cb, qual, eligible) n x L arrays."""
    L, U = c["cb_length"], c["umi_length"]
    cbs, qs, el = [], [], []
    for (id1, s1, q1, id2, s2, q2) in c["pairs"]:
        ok = A.removesuffix(id1, "/1") == A.removesuffix(id2, "/2") and len(s1) > L + U
        el.append(ok)
        cbs.append(s1[:L].ljust(L, "A") if ok else "A" * L)
        qs.append(quals(q1[:L]) if ok else [0] * L)
    cb = np.frombuffer("".join(cbs).encode("latin-1"), np.uint8).reshape(-1, L)
    return cb, np.array(qs, np.uint8).reshape(-1, L), np.array(el, np.uint8)


def test_fixture_shape():
    assert len(CASES) >= 100
    assert sum(c["n_multi_candidate"] > 0 for c in CASES) >= 20      # the quality rule is exercised
    assert sum(c["tie"] for c in CASES) >= 10                         # ... including order-dependent ties
    assert any(c["stats"]["name_mismatch"] for c in CASES) and any(c["stats"]["too_short"] for c in CASES)
    assert any(c["stats"]["no_remaining_seq"] for c in CASES) and any(c["stats"]["cb_no_correction"] for c in CASES)
    assert {c["cb_length"] for c in CASES} >= {5, 8, 12, 16, 21}


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_python_oracle_matches_reference(ci):
    c = CASES[ci]
    wl = A.Whitelist(c["whitelist"], c["cb_length"])
    pairs = [(a, b, quals(q), d, e, quals(f)) for (a, b, q, d, e, f) in c["pairs"]]
    recs, st = A.process_pairs(pairs, wl, c["cb_length"], c["umi_length"])
    assert st == c["stats"]
    L, U = c["cb_length"], c["umi_length"]
    by_name = {}
    for pi, (id1, s1, q1, id2, s2, q2) in enumerate(c["pairs"]):
        by_name.setdefault(A.removesuffix(id1, "/1"), pi)
    got = []
    k = 0
    for pi, cb in c["records"]:
        r = recs[k]; k += 1
        id1, s1, q1, id2, s2, q2 = c["pairs"][pi]
        assert r["cb"] == cb and r["name"] == A.removesuffix(id1, "/1") and r["umi"] == s1[L:L + U]
        assert r["r1_seq"] == s1[L + U:] and r["r1_qual"] == quals(q1[L + U:])
        assert r["r2_seq"] == s2 and r["r2_qual"] == quals(q2)
    assert k == len(recs)


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_c_oracle_matches_reference(ci):
    c = CASES[ci]
    L = c["cb_length"]
    cb, q, el = case_arrays(c)
    idx, status, st = O.cb_correct(c["whitelist"], cb, q, el, L)
    got = [[k, c["whitelist"][idx[k]]] for k in range(len(el)) if status[k] in (A.CB_PERFECT, A.CB_CORRECTED)]
    assert got == c["records"]
    for k in ("cb_perfect_match", "cb_corrected", "cb_no_correction", "cache_size"):
        assert st[k] == c["stats"][k], k
    assert ((status == A.CB_SKIPPED) == (el == 0)).all()


@pytest.mark.parametrize("L,clustered", [(16, 0.0), (16, 0.5), (10, 0.6), (21, 0.3)])
def test_c_oracle_matches_python_oracle_random(L, clustered):
    wl, cb, q = synth.barcode_workload(6000, n_whitelist=3000, n_cells=300, cb_length=L, err_rate=0.03, n_rate=0.004,
                                       off_whitelist=0.03, seed=L, clustered=clustered)
    q = (q // 8 * 8).astype(np.uint8)                                  # few distinct qualities: ties
    rng = np.random.default_rng(L)
    el = (rng.random(len(cb)) < 0.95).astype(np.uint8)
    wl_s = [bytes(r).decode() for r in wl]
    cbs = [bytes(r).decode() for r in cb]
    pi, ps, cache = A.correct_batch(cbs, q.tolist(), el.tolist(), A.Whitelist(wl_s, L))
    idx, status, st = O.cb_correct(wl_s, cb, q, el, L)
    assert list(idx) == pi and list(status) == ps and st["cache_size"] == cache
    assert (status == A.CB_CORRECTED).sum() > 100 and (status == A.CB_NONE).sum() > 10
