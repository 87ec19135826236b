"""GPU parity on the edge cases and config space (all through the C ABI, bit-exact vs the oracle)."""
import os

import numpy as np
import pytest

from nimble_b200 import synth
from oracle import oracle as O
from helpers import diff_results, oracle_counts, table_tuple, to_concat
from test_oracle_align import REF_A, REF_A2, REF_B, lib_of, rc

pytestmark = pytest.mark.gpu


def both(engine, lib_json, r1, r2=None, key=None, k=20, strand="unstranded", threshold=0.05, disable=False):
    lo = O.Library(lib_json, k=k, strand_filter=strand)
    ro, fo = O.align(lo, to_concat(r1), to_concat(r2) if r2 is not None else None)
    lg = engine.load_library(lib_json, strand_filter=strand, k=k)
    table, rg, fg = engine.align(lg, r1, r2, key=key, threshold=threshold, disable_thresholding=disable, per_read=True)
    bad = diff_results(ro, fo, rg, fg)
    assert not bad, "\n".join(bad)
    if key is not None:
        cell, cnt, off, ids, dropped = oracle_counts(lo, ro, fo, key, threshold, disable)
        assert table_tuple(table.cell, table.count, table.feat_off, table.feat_ids) == table_tuple(cell, cnt, off, ids)
        assert table.dropped_empty == dropped
    return lo, ro, fo, table


def test_empty_input(engine):
    lg = engine.load_library(lib_of([REF_A, REF_B]))
    table, res, feats = engine.align(lg, [], key=np.zeros(0, np.uint64), per_read=True)
    assert len(table) == 0 and len(res) == 0 and table.dropped_empty == 0


def test_ragged_short_and_n_reads(engine):
    rng = np.random.default_rng(5)
    reads = [REF_A[:19], "", "N" * 50, REF_A[:24], REF_A[:30], REF_A[:30] + "N" * 40, REF_A[3:23], rc(REF_B[100:350]),
             REF_B[0:500 - 100], "ACGT" * 100]
    for _ in range(400):
        src = [REF_A, REF_A2, REF_B][int(rng.integers(0, 3))]
        L = int(rng.integers(1, 300))
        a = int(rng.integers(0, len(src) - L + 1))
        s = list(src[a:a + L])
        for _e in range(int(rng.integers(0, 4))):
            s[int(rng.integers(0, L))] = "ACGTN"[int(rng.integers(0, 5))]
        s = "".join(s)
        reads.append(rc(s.replace("N", "A")) if rng.random() < 0.3 else s)
    key = (rng.integers(0, 5, size=len(reads)).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 30, size=len(reads)).astype(np.uint64)
    key[::17] = np.uint64(0xFFFFFFFFFFFFFFFF)       # reads without CB/UB tags are aligned but not counted
    lib = lib_of([REF_A, REF_A2, REF_B], names=["A*01", "A*02", "B*01"])
    both(engine, lib, reads, key=key)


def test_compact_and_full_wire_forms_agree(engine):
    """nb200_align with the compact wire form (seq words + side table of the reads with N, expanded on the device)
    == the full records, per read and in the count table; single-end and paired, resident upload too."""
    rng = np.random.default_rng(15)
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=8, length=500, snps_mean=6, seed=150)
    a1, a2, _ = synth.sample_pairs(codes, 6000, read_len=120, insert_mean=260, insert_sd=30, err_rate=0.01, off_target=0.1, seed=151)
    for a in (a1, a2):                                # every 9th read gets a few non-ACGT bases, some runs of them
        for i in range(0, len(a), 9):
            for j in rng.integers(0, a.shape[1], size=int(rng.integers(1, 4))):
                a[i, j:j + int(rng.integers(1, 6))] = ord("N")
    key = (rng.integers(0, 7, size=len(a1)).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 50, size=len(a1)).astype(np.uint64)
    lg = engine.load_library(lib, k=20)
    for r2 in (None, a2):
        pc1, pf1 = engine.pack(a1, compact=True), engine.pack(a1, compact=False)
        assert pc1.compact and not pf1.compact and len(pc1.n_idx) == len(range(0, len(a1), 9)) and pc1.stride < pf1.stride
        pc2 = engine.pack(r2, compact=True) if r2 is not None else None
        pf2 = engine.pack(r2, compact=False) if r2 is not None else None
        tc, rc_, fc = engine.align(lg, pc1, pc2, key=key, per_read=True)
        tf, rf, ff = engine.align(lg, pf1, pf2, key=key, per_read=True)
        assert not diff_results(rf, ff, rc_, fc)
        assert table_tuple(tc.cell, tc.count, tc.feat_off, tc.feat_ids) == table_tuple(tf.cell, tf.count, tf.feat_off, tf.feat_ids)
        engine.upload(pc1, pc2, key=key)
        tr = engine.align_resident(lg)
        rr, fr = engine.fetch_results(lg)
        assert not diff_results(rf, ff, rr, fr)
        assert table_tuple(tr.cell, tr.count, tr.feat_off, tr.feat_ids) == table_tuple(tf.cell, tf.count, tf.feat_off, tf.feat_ids)
    both(engine, lib, a1, a2, key=key)                # and both equal the oracle (the default pack is the compact form)


def test_indels_and_band_limits(engine):
    reads = []
    for d in range(1, 12):
        reads.append(REF_B[100:140] + REF_B[140 + d:190 + d])                   # deletion of d bases
        reads.append(REF_B[100:140] + "ACGTACGTACGT"[:d] + REF_B[140:190])      # insertion of d bases
    lo, ro, fo, _ = both(engine, lib_of([REF_A, REF_B]), reads)
    assert ro["n_sw"].sum() == len(reads)


@pytest.mark.parametrize("k", [8, 12, 31, 32])
def test_k_values(engine, k):
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=7, length=450, snps_mean=7, seed=70 + k)
    r1, truth = synth.sample_reads(codes, 5000, read_len=100, seed=71)
    key = synth.barcodes_10x(len(r1), n_cells=25, seed=72, truth=truth)
    both(engine, lib, r1, key=key, k=k)


@pytest.mark.parametrize("n_refs,wpl", [(20, 1), (1030, 2), (2100, 4), (4200, 8)])
def test_bitset_width_variants(engine, n_refs, wpl):
    per = 35
    lib, codes = synth.allele_family_library(n_founders=max(1, n_refs // per), alleles_per_founder=per, length=260,
                                             snps_mean=5, seed=80 + wpl, config={"max_hits_to_report": 12})
    r1, truth = synth.sample_reads(codes, 4000, read_len=90, seed=81)
    key = synth.barcodes_10x(len(r1), n_cells=20, seed=82, truth=truth)
    both(engine, lib, r1, key=key)


CONFIGS = [
    {"num_mismatches": 1}, {"num_mismatches": 3}, {"discard_multiple_matches": True}, {"discard_multi_hits": 2},
    {"max_hits_to_report": 3}, {"score_threshold": 60, "score_filter": 70}, {"score_percent": 0.95}, {"score_percent": 0.0},
    {"require_valid_pair": True}, {"intersect_level": 2, "num_mismatches": 2},
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[str(c) for c in CONFIGS])
def test_config_space(engine, cfg):
    lib, codes = synth.allele_family_library(n_founders=5, alleles_per_founder=10, length=700, snps_mean=9, seed=90, config=cfg)
    r1, r2, truth = synth.sample_pairs(codes, 3000, read_len=120, insert_mean=260, seed=91, err_rate=0.01)
    key = synth.barcodes_10x(len(r1), n_cells=15, seed=92, truth=truth)
    both(engine, lib, r1, r2, key=key)


def test_group_on_gene_level(engine):
    lib, codes = synth.allele_family_library(n_founders=6, alleles_per_founder=9, length=500, snps_mean=8, seed=93,
                                             extra_columns=True, config={"group_on": "gene"})
    r1, truth = synth.sample_reads(codes, 6000, read_len=90, seed=94)
    key = synth.barcodes_10x(len(r1), n_cells=30, seed=95, truth=truth)
    lo, ro, fo, table = both(engine, lib, r1, key=key)
    assert len(lo.features) == 6 and (ro["reason"] == 0).mean() > 0.6


@pytest.mark.parametrize("thr,disable", [(0.0, False), (0.2, False), (0.5, False), (0.05, True)])
def test_umi_thresholds(engine, thr, disable):
    lib, codes = synth.allele_family_library(n_founders=5, alleles_per_founder=10, length=500, snps_mean=8, seed=96)
    r1, truth = synth.sample_reads(codes, 12000, read_len=90, seed=97)
    key = synth.barcodes_10x(len(r1), n_cells=10, reads_per_umi=8.0, seed=98)      # UMIs mix alleles -> real thresholding work
    both(engine, lib, r1, key=key, threshold=thr, disable=disable)


def test_multiple_libraries_and_set_config(engine):
    lib1, codes1 = synth.allele_family_library(n_founders=4, alleles_per_founder=6, length=400, seed=101)
    lib2, codes2 = synth.allele_family_library(n_founders=3, alleles_per_founder=5, length=500, seed=102, name_prefix="KIR")
    r1, _ = synth.sample_reads(codes1 + codes2, 5000, read_len=90, seed=103)
    g1, g2 = engine.load_library(lib1), engine.load_library(lib2)
    packed = engine.pack(r1)
    for lg, lj in ((g1, lib1), (g2, lib2), (g1, lib1)):
        lo = O.Library(lj)
        ro, fo = O.align(lo, to_concat(r1))
        _, rg, fg = engine.align(lg, packed, per_read=True)
        assert not diff_results(ro, fo, rg, fg)
    g1.set_config(score_percent=0.9, strand_filter="fiveprime")
    lo = O.Library(lib1, strand_filter="fiveprime")
    lo.cfg.score_percent = 0.9
    ro, fo = O.align(lo, to_concat(r1))
    _, rg, fg = engine.align(g1, packed, per_read=True)
    assert not diff_results(ro, fo, rg, fg)


def test_resident_path_equals_host_path(engine):
    lib, codes = synth.allele_family_library(n_founders=5, alleles_per_founder=8, length=500, seed=111)
    r1, truth = synth.sample_reads(codes, 30000, read_len=90, seed=112)
    key = synth.barcodes_10x(len(r1), n_cells=40, seed=113, truth=truth)
    lg = engine.load_library(lib)
    t1, res1, f1 = engine.align(lg, r1, key=key, per_read=True)
    engine.upload(r1, key=key)
    t2 = engine.align_resident(lg)
    res2, f2 = engine.fetch_results(lg)
    assert np.array_equal(res1, res2) and np.array_equal(f1, f2)
    assert table_tuple(t1.cell, t1.count, t1.feat_off, t1.feat_ids) == table_tuple(t2.cell, t2.count, t2.feat_off, t2.feat_ids)
    tm = engine.timing()
    assert tm["launches"] > 10 and tm["probes"] == 0            # device counters of the probe are off by default
    engine.set_stats(True)
    try:
        t3 = engine.align_resident(lg)
        tm = engine.timing()
    finally:
        engine.set_stats(False)
    # every position without an N is looked up once; a lookup reads one sector, sometimes two
    n_lookups = sum(sum(1 for i in range(90 - 20 + 1) if b"N" not in bytes(r[i:i + 20])) for r in r1)
    assert tm["probes"] == n_lookups and tm["probes"] <= tm["probe_slots"] <= 1.2 * tm["probes"]
    assert table_tuple(t3.cell, t3.count, t3.feat_off, t3.feat_ids) == table_tuple(t2.cell, t2.count, t2.feat_off, t2.feat_ids)


def test_bad_inputs_fail_loudly(engine):
    import nimble_b200
    with pytest.raises(nimble_b200.NimbleB200Error):
        engine.load_library("/nonexistent/lib.json")
    with pytest.raises(nimble_b200.NimbleB200Error):
        engine.load_library(lib_of([REF_A]), strand_filter="sideways")
    with pytest.raises(nimble_b200.NimbleB200Error):
        engine.pack(["A" * 501])
    lg = engine.load_library(lib_of([REF_A, REF_B]))
    with pytest.raises(nimble_b200.NimbleB200Error):
        engine.align(lg, [REF_A[:90]], [REF_A[:90], REF_A[:90]])          # mate count mismatch
    with pytest.raises(nimble_b200.NimbleB200Error):
        engine.align(lg, [REF_A[:90]], key=np.zeros(3, np.uint64))


def _run_in_subprocess(env_extra, body):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "import sys, numpy as np\nsys.path.insert(0, %r); sys.path.insert(0, %r)\n" % (root, os.path.join(root, "tests")) + body
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env_extra), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "sub-ok" in out.stdout, out.stdout + out.stderr
    return out.stdout


WIDE_BODY = r"""
import nimble_b200
from nimble_b200 import synth
from oracle import oracle as O
from helpers import diff_results, oracle_counts, table_tuple, to_concat
eng = nimble_b200.Engine(0)
for paired in (False, True):
    for cfg in ({}, {"group_on": "gene", "num_mismatches": 1}, {"intersect_level": 1}):
        lib, codes = synth.allele_family_library(n_founders=5, alleles_per_founder=40, length=420, snps_mean=6, seed=141,
                                                 extra_columns=True, config=cfg)
        if paired:
            r1, r2, truth = synth.sample_pairs(codes, 2500, read_len=110, insert_mean=240, err_rate=0.01, seed=142)
        else:
            (r1, truth), r2 = synth.sample_reads(codes, 4000, read_len=90, err_rate=0.01, seed=143), None
        key = synth.barcodes_10x(len(r1), n_cells=12, seed=144, truth=truth)
        lo = O.Library(lib); lg = eng.load_library(lib)
        ro, fo = O.align(lo, to_concat(r1), to_concat(r2) if r2 is not None else None)
        table, rg, fg = eng.align(lg, r1, r2, key=key, per_read=True)
        bad = diff_results(ro, fo, rg, fg)
        assert not bad, "\n".join(bad)
        cell, cnt, off, ids, dropped = oracle_counts(lo, ro, fo, key)
        assert table_tuple(table.cell, table.count, table.feat_off, table.feat_ids) == table_tuple(cell, cnt, off, ids)
        assert ro["n_sw"].sum() > 200
print("sub-ok")
"""


@pytest.mark.parametrize("cap", ["0", "1"])
def test_wide_read_path(cap):
    """NB200_NARROW_CAP forces reads through wide_kernel (global scratch, in-kernel SW): same answers."""
    _run_in_subprocess({"NB200_NARROW_CAP": cap}, WIDE_BODY)


def test_library_beyond_8192_references(engine):
    lib, codes = synth.allele_family_library(n_founders=300, alleles_per_founder=30, length=150, snps_mean=3, seed=151,
                                             config={"max_hits_to_report": 40})
    assert len(codes) == 9000
    r1, truth = synth.sample_reads(codes, 6000, read_len=80, seed=152)
    key = synth.barcodes_10x(len(r1), n_cells=20, seed=153, truth=truth)
    both(engine, lib, r1, key=key)


def test_class_wider_than_five_words(engine):
    """A segment shared by 400 references: classes spill to the overflow list (13 words)."""
    rng = np.random.default_rng(161)
    shared = "".join(rng.choice(list("ACGT"), size=150))
    seqs = ["".join(rng.choice(list("ACGT"), size=60)) + shared + "".join(rng.choice(list("ACGT"), size=60)) for _ in range(400)]
    lib = lib_of(seqs, max_hits_to_report=64)
    reads = [shared[10:100], shared[30:140], seqs[7][20:110], seqs[399][100:200], rc(shared[5:95])]
    reads += [seqs[int(i)][int(a):int(a) + 90] for i, a in zip(rng.integers(0, 400, 300), rng.integers(0, 180, 300))]
    mut = list(shared[20:120]); mut[50] = "A" if mut[50] != "A" else "C"
    reads.append("".join(mut))                      # SW against 400 candidates
    lo, ro, fo, _ = both(engine, lib, reads)
    assert ro["n_cand"][0][0] == 400 and ro["reason"][0] == 6 and ro["n_cand"][-1][0] == 400 and ro["n_sw"][-1] == 1


def test_sw_worklist_overflow_retry():
    """Force a tiny Smith-Waterman work list: the engine must grow it and redo the pass."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import nimble_b200
from nimble_b200 import synth
from oracle import oracle as O
from helpers import diff_results, to_concat
lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=8, length=400, seed=121)
r1, _ = synth.sample_reads(codes, 5000, read_len=90, err_rate=0.02, seed=122)
eng = nimble_b200.Engine(0)
lg = eng.load_library(lib)
_, rg, fg = eng.align(lg, r1, per_read=True)
lo = O.Library(lib); ro, fo = O.align(lo, to_concat(r1))
assert ro["n_sw"].sum() > 500
assert not diff_results(ro, fo, rg, fg)
print("retry-ok", eng.timing()["sw_pairs"])
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NB200_ITEMS_CAP="64")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "retry-ok" in out.stdout, out.stdout + out.stderr


def test_counts_device_matches_host_table(engine):
    """nb200_counts_device: the device copy the multi-GPU gather reads is the table the call returned."""
    import torch
    lib, codes = synth.allele_family_library(n_founders=3, alleles_per_founder=6, length=300, snps_mean=5, seed=21)
    r1, truth = synth.sample_reads(codes, 3000, read_len=90, seed=22)
    key = synth.barcodes_10x(len(r1), n_cells=15, seed=22, truth=truth)
    lg = engine.load_library(lib, k=20)
    table = engine.align(lg, r1, key=key)
    dv = engine.counts_device()

    class Arr:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}

    got = {k: torch.as_tensor(Arr(*v), device="cuda").cpu().numpy().view(np.uint32) for k, v in dv.items() if v[1]}
    assert len(table) > 0
    assert np.array_equal(got["cell"], table.cell) and np.array_equal(got["count"], table.count)
    assert np.array_equal(got["feat_off"], table.feat_off) and np.array_equal(got["feat_ids"], table.feat_ids)


def _fuzz_case(seed):
    """Random library shape, read shape and EVERY config knob (nimble/types.py:12-25), all seeded."""
    rng = np.random.default_rng(seed)
    k = int(rng.choice([6, 11, 16, 20, 20, 20, 25, 31, 32]))
    cfg = {
        "score_threshold": int(rng.choice([0, 10, 20, 20, 45, 80])), "score_filter": int(rng.choice([0, 25, 25, 60])),
        "score_percent": float(rng.choice([0.0, 0.25, 0.5, 0.5, 0.8, 1.0])), "num_mismatches": int(rng.choice([0, 0, 1, 2, 5])),
        "discard_multiple_matches": bool(rng.random() < 0.2), "intersect_level": int(rng.choice([0, 1, 2])),
        "discard_multi_hits": int(rng.choice([0, 0, 0, 1, 3])), "require_valid_pair": bool(rng.random() < 0.3),
        "max_hits_to_report": int(rng.choice([1, 2, 5, 10, 10, 17])),
    }
    grouped = rng.random() < 0.3
    if grouped:
        cfg["group_on"] = "gene"
    lib, codes = synth.allele_family_library(n_founders=int(rng.integers(1, 8)), alleles_per_founder=int(rng.integers(1, 40)),
                                             length=int(rng.integers(max(60, k + 20), 900)), snps_mean=float(rng.choice([0.0, 2.0, 8.0, 25.0])),
                                             seed=seed, config=cfg, extra_columns=grouped)
    min_len = min(len(c) for c in codes)
    paired = rng.random() < 0.5
    n = int(rng.integers(200, 2500))
    if paired:
        rl = int(rng.integers(20, max(21, min(251, min_len - 25))))
        r1, r2, truth = synth.sample_pairs(codes, n, read_len=rl, insert_mean=min(min_len, int(rl * rng.uniform(1.0, 2.5))),
                                           insert_sd=20, err_rate=float(rng.choice([0.0, 0.005, 0.03])), off_target=0.15, seed=seed + 1)
    else:
        rl = int(rng.integers(15, max(16, min(301, min_len - 20))))
        r1, truth = synth.sample_reads(codes, n, read_len=rl, err_rate=float(rng.choice([0.0, 0.005, 0.03])), off_target=0.15,
                                       rc_frac=float(rng.choice([0.0, 0.1, 0.5])), seed=seed + 1)
        r2 = None
    # sprinkle N bases and ragged lengths
    reads1 = [bytes(r).decode() for r in r1]
    reads2 = [bytes(r).decode() for r in r2] if r2 is not None else None
    for i in rng.choice(n, size=n // 20, replace=False):
        s = list(reads1[i]); s[int(rng.integers(0, len(s)))] = "N"; reads1[i] = "".join(s)
    for i in rng.choice(n, size=n // 25, replace=False):
        reads1[i] = reads1[i][:int(rng.integers(0, len(reads1[i]) + 1))]
        if reads2 is not None and rng.random() < 0.5:
            reads2[i] = reads2[i][int(rng.integers(0, len(reads2[i]))):]
    key = None
    if rng.random() < 0.8:
        key = synth.barcodes_10x(n, n_cells=int(rng.integers(1, 40)), reads_per_umi=float(rng.choice([1.0, 3.0, 10.0])), seed=seed + 2,
                                 truth=truth if rng.random() < 0.5 else None)
        key[rng.random(n) < 0.03] = np.uint64(0xFFFFFFFFFFFFFFFF)
    strand = str(rng.choice(["unstranded", "fiveprime", "threeprime", "none"]))
    thr = float(rng.choice([0.0, 0.05, 0.05, 0.2, 0.5]))
    return lib, reads1, reads2, key, k, strand, thr, bool(rng.random() < 0.15)


@pytest.mark.parametrize("seed", range(1000, 1000 + int(os.environ.get("NB200_FUZZ_CASES", "40"))))
def test_randomised_differential(engine, seed):
    """Seeded fuzz over library shape, read shape, k, strand filter and every config knob: per-read records,
    feature calls and the count table must equal the oracle's bit for bit."""
    lib, r1, r2, key, k, strand, thr, disable = _fuzz_case(seed)
    both(engine, lib, r1, r2, key=key, k=k, strand=strand, threshold=thr, disable=disable)


def test_very_large_umi_group_takes_the_global_sort(engine):
    """Rows of a (cell, umi) group are ordered by a per-group sort; a group larger than that sort handles (data
    without real UMIs) must fall back to the global token sort and give the same table as the oracle."""
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=9, length=400, snps_mean=7, seed=301)
    r1, truth = synth.sample_reads(codes, 6000, read_len=90, seed=302)
    key = synth.barcodes_10x(len(r1), n_cells=6, reads_per_umi=2.0, seed=303, truth=truth)
    key[:2500] = key[0]                                     # one group of 2500 reads with mixed feature lists
    key[2500:2560] = key[2500]                              # and one of 60 (just above the per-group limit)
    both(engine, lib, r1, key=key, threshold=0.05)
    both(engine, lib, r1, key=key, threshold=0.3)


def test_batch_pipelining_on_and_off_agree(engine):
    """Several batches in flight (two streams, double-buffered per-batch state) vs kernels back to back."""
    lib, codes = synth.allele_family_library(n_founders=6, alleles_per_founder=12, length=600, snps_mean=8, seed=401)
    n = (1 << 21) + (1 << 20) + 12345                       # three batches of the resident path
    r1, truth = synth.sample_reads(codes, n, read_len=90, seed=402)
    key = synth.barcodes_10x(n, n_cells=200, seed=403, truth=truth)
    lg = engine.load_library(lib)
    packed = engine.pack(r1, pinned=False)
    tup = lambda t: (t.cell.tolist(), t.count.tolist(), t.feat_off.tolist(), t.feat_ids.tolist(), t.dropped_empty, t.n_called)
    try:
        engine.set_overlap(True)
        a, ra, fa = engine.align(lg, packed, key=key, per_read=True)
        engine.set_overlap(False)
        b, rb, fb = engine.align(lg, packed, key=key, per_read=True)
    finally:
        engine.set_overlap(True)
    assert tup(a) == tup(b) and np.array_equal(fa, fb)
    for f in ("score", "edits", "status", "reason", "n_feat", "n_cand", "n_hits", "pair_score"):
        assert np.array_equal(ra[f], rb[f]), f
    m = 50_000                                              # and both agree with the oracle on a slice spanning no batch edge...
    lo = O.Library(lib)
    ro, fo = O.align(lo, to_concat(r1[(1 << 21) - m // 2:(1 << 21) + m // 2]))       # ... and one across the first batch boundary
    sl = slice((1 << 21) - m // 2, (1 << 21) + m // 2)
    bad = diff_results(ro, fo, ra[sl], fa[sl])
    assert not bad, "\n".join(bad)


def test_deferred_fetch_gives_the_same_table(engine):
    """nb200_set_defer_fetch: the call returns with the table on the device only; nb200_fetch_counts brings the same rows."""
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=8, length=400, snps_mean=6, seed=160)
    r1, truth = synth.sample_reads(codes, 20000, read_len=90, seed=161)
    key = synth.barcodes_10x(len(r1), n_cells=40, seed=162, truth=truth)
    lg = engine.load_library(lib, k=20)
    want = engine.align(lg, r1, key=key)
    engine.set_defer_fetch(True)
    try:
        n_rows = engine.align(lg, r1, key=key, fetch_counts=False)
        assert n_rows == len(want)
        dv = engine.counts_device()
        assert dv["cell"][1] == n_rows and dv["feat_ids"][1] == len(want.feat_ids)
        got = engine.fetch_counts()
        got2 = engine.fetch_counts()                      # a second fetch hands out the same host table
        engine.upload(r1, key=key)
        assert engine.align_resident(lg, fetch_counts=False) == n_rows
        engine.fetch_counts_start()                       # copies in flight on the second copy stream
        got3 = engine.fetch_counts()
        assert engine.align_resident(lg, fetch_counts=False) == n_rows
        engine.fetch_counts_start()                       # started and never waited for: the next call cleans up
        assert engine.align_resident(lg, fetch_counts=False) == n_rows
    finally:
        engine.set_defer_fetch(False)
    for t in (got, got2, got3):
        assert table_tuple(t.cell, t.count, t.feat_off, t.feat_ids) == table_tuple(want.cell, want.count, want.feat_off, want.feat_ids)
    again = engine.align(lg, r1, key=key)                 # and the normal mode is back
    assert table_tuple(again.cell, again.count, again.feat_off, again.feat_ids) == table_tuple(want.cell, want.count, want.feat_off, want.feat_ids)
