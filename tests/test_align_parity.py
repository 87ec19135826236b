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
