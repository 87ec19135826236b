"""GPU parity: the CUDA path through the C ABI vs the C oracle, bit-exact, same seeded inputs."""
import numpy as np
import pytest

from nimble_b200 import synth
from oracle import oracle as O
from helpers import diff_results, oracle_counts, table_tuple, to_concat

pytestmark = pytest.mark.gpu


def run_both(engine, lib_json, r1, r2=None, key=None, k=20, strand="unstranded", threshold=0.05):
    lo = O.Library(lib_json, k=k, strand_filter=strand)
    ro, fo = O.align(lo, to_concat(r1), to_concat(r2) if r2 is not None else None)
    lg = engine.load_library(lib_json, strand_filter=strand, k=k)
    assert lg.feature_names == lo.features
    table, rg, fg = engine.align(lg, r1, r2, key=key, threshold=threshold, per_read=True)
    bad = diff_results(ro, fo, rg, fg)
    assert not bad, "\n".join(bad)
    if key is not None:
        cell, cnt, off, ids, dropped = oracle_counts(lo, ro, fo, key, threshold)
        assert table_tuple(table.cell, table.count, table.feat_off, table.feat_ids) == table_tuple(cell, cnt, off, ids)
        assert table.dropped_empty == dropped
    return ro, fo, table


def test_single_end_mhc_default(engine):
    lib, codes = synth.allele_family_library(n_founders=6, alleles_per_founder=12, length=600, snps_mean=8, seed=11)
    r1, truth = synth.sample_reads(codes, 20000, read_len=90, seed=12)
    key = synth.barcodes_10x(len(r1), n_cells=50, seed=12, truth=truth)
    ro, fo, table = run_both(engine, lib, r1, key=key)
    assert (ro["reason"] == 0).sum() > 5000 and len(table) > 50
    assert ro["n_sw"].sum() > 1000          # the Smith-Waterman kernel really ran


@pytest.mark.parametrize("strand", ["unstranded", "fiveprime", "threeprime", "none"])
@pytest.mark.parametrize("level", [0, 1, 2])
def test_paired_end_kir_like(engine, strand, level):
    lib, codes = synth.allele_family_library(n_founders=5, alleles_per_founder=10, length=1300, snps_mean=10, seed=21,
                                             name_prefix="KIR", config={"intersect_level": level})
    r1, r2, truth = synth.sample_pairs(codes, 6000, read_len=150, seed=22)
    run_both(engine, lib, r1, r2, strand=strand)


def test_bulk_counts_match_oracle_histogram(engine):
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=8, length=500, snps_mean=6, seed=31)
    r1, _ = synth.sample_reads(codes, 8000, read_len=75, seed=32)
    lo = O.Library(lib)
    ro, fo = O.align(lo, to_concat(r1))
    lg = engine.load_library(lib)
    table = engine.align(lg, r1)
    hist = {}
    for i in np.nonzero(ro["n_feat"])[0]:
        t = tuple(int(x) for x in fo[i, :ro["n_feat"][i]])
        hist[t] = hist.get(t, 0) + 1
    got = {tuple(int(x) for x in table.feat_ids[table.feat_off[i]:table.feat_off[i + 1]]): int(table.count[i])
           for i in range(len(table))}
    assert got == hist
    assert all(int(c) == 0 for c in table.cell)


def test_full_size_properties(engine):
    """BASELINE.json configs[1] at full size (10 M reads x 90 bp + CB/UB vs 40 x 50 alleles): properties that
    need no oracle, plus a bounded slice against it."""
    n = 10_000_000
    lib, codes = synth.allele_family_library(n_founders=40, alleles_per_founder=50, length=1098, snps_mean=15.0, seed=1)
    r1, truth = synth.sample_reads(codes, n, read_len=90, err_rate=0.005, off_target=0.2, rc_frac=0.1, seed=2)
    key = synth.barcodes_10x(n, n_cells=10000, seed=2, truth=truth)
    lg = engine.load_library(lib, k=20)
    packed = engine.pack(r1, pinned=False)
    whole = engine.align(lg, packed, key=key)
    tup = lambda t: (t.cell.tolist(), t.count.tolist(), t.feat_off.tolist(), t.feat_ids.tolist())
    assert len(whole) > 1_000_000 and whole.n_called > 4_000_000
    assert int(whole.count.sum()) + whole.dropped_empty <= whole.n_umis
    # rows are in the reference's output order: ascending cell, then ascending feature string
    assert (np.diff(whole.cell.astype(np.int64)) >= 0).all()
    # idempotence: a second pass over the same resident batch gives the same table
    again = engine.align(lg, packed, key=key)
    assert tup(again) == tup(whole)
    # shard invariance (what the multi-GPU run relies on): split by cell -> the two tables are the whole table
    from nimble_b200 import shard
    owner = shard.shard_of_key(key, 2)
    parts = []
    for rk in (0, 1):
        sel = np.nonzero(owner == rk)[0]
        parts.append(engine.align(lg, r1[sel], key=key[sel]))
    def row_sig(t):
        """Per count row: (cell, count, n_feat, order-sensitive hash of the feature ids)."""
        off = t.feat_off.astype(np.int64)
        nf = off[1:] - off[:-1]
        pos = np.arange(len(t.feat_ids), dtype=np.uint64) - np.repeat(off[:-1], nf).astype(np.uint64)
        w = (t.feat_ids.astype(np.uint64) + np.uint64(1)) * ((pos + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        h = np.add.reduceat(w, off[:-1]) if len(w) else np.zeros(0, np.uint64)
        return np.stack([t.cell.astype(np.uint64), t.count.astype(np.uint64), nf.astype(np.uint64), h], axis=1)

    sig = np.concatenate([row_sig(parts[0]), row_sig(parts[1])])
    sig = sig[np.argsort(sig[:, 0], kind="stable")]        # cells are disjoint between shards: a merge by cell
    assert np.array_equal(sig, row_sig(whole))
    # order invariance: reads of a (cell, umi) may arrive in any order
    perm = np.random.default_rng(5).permutation(n)
    shuffled = engine.align(lg, r1[perm], key=key[perm])
    assert tup(shuffled) == tup(whole)
    # a bounded slice against the oracle, per read and per count row
    m = 200_000
    lo = O.Library(lib, k=20)
    ro, fo = O.align(lo, to_concat(r1[:m]))
    table, rg, fg = engine.align(lg, r1[:m], key=key[:m], per_read=True)
    bad = diff_results(ro, fo, rg, fg)
    assert not bad, "\n".join(bad)
    cell, cnt, off, ids, dropped = oracle_counts(lo, ro, fo, key[:m], 0.05)
    assert np.array_equal(table.cell, cell) and np.array_equal(table.count, cnt) and np.array_equal(table.feat_ids, ids)


def test_full_size_properties_paired_bulk(engine):
    """BASELINE.json configs[2] shape (bulk paired-end 2x150 bp vs KIR-like library, strict config, strand filter,
    intersect feature-calling) at 2 M pairs: idempotence, order invariance of the per-feature-set histogram, and a
    bounded slice against the oracle."""
    n = 2_000_000
    lib, codes = synth.allele_family_library(n_founders=17, alleles_per_founder=90, length=1350, snps_mean=12.0, seed=3,
                                             name_prefix="KIR", config={"num_mismatches": 0, "intersect_level": 2})
    r1, r2, _ = synth.sample_pairs(codes, n, read_len=150, insert_mean=300, insert_sd=50, err_rate=0.005, off_target=0.1, seed=3)
    lg = engine.load_library(lib, strand_filter="fiveprime", k=20)
    whole = engine.align(lg, r1, r2)
    tup = lambda t: (t.cell.tolist(), t.count.tolist(), t.feat_off.tolist(), t.feat_ids.tolist())
    assert len(whole) > 1000 and int(whole.count.sum()) == whole.n_called and whole.n_called > 300_000
    assert (whole.cell == 0).all()
    assert tup(engine.align(lg, r1, r2)) == tup(whole)
    perm = np.random.default_rng(9).permutation(n)
    assert tup(engine.align(lg, r1[perm], r2[perm])) == tup(whole)
    # the two halves add up to the whole (bulk counts are a histogram over feature sets)
    def hist(t):
        return {tuple(t.feat_ids[t.feat_off[i]:t.feat_off[i + 1]].tolist()): int(t.count[i]) for i in range(len(t))}
    ha, hb, hw = hist(engine.align(lg, r1[:n // 2], r2[:n // 2])), hist(engine.align(lg, r1[n // 2:], r2[n // 2:])), hist(whole)
    merged = dict(ha)
    for k_, v in hb.items():
        merged[k_] = merged.get(k_, 0) + v
    assert merged == hw
    m = 60_000
    lo = O.Library(lib, k=20, strand_filter="fiveprime")
    ro, fo = O.align(lo, to_concat(r1[:m]), to_concat(r2[:m]))
    table, rg, fg = engine.align(lg, r1[:m], r2[:m], per_read=True)
    bad = diff_results(ro, fo, rg, fg)
    assert not bad, "\n".join(bad)
