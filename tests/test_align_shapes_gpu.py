"""GPU parity on the two BASELINE.json shapes that only bench.py exercised in round 1:
configs[3] (combined MHC + KIR + immune-gene library, union feature-calling) and configs[4]
(transcriptome-scale library, k = 31, lenient score_percent, dense hash table).  Per-read records and
the count table are diffed against the C oracle, bit for bit."""
import os

import numpy as np
import pytest

from nimble_b200 import synth
from oracle import oracle as O
from helpers import diff_results, oracle_counts, table_tuple, to_concat

pytestmark = pytest.mark.gpu


def both(engine, lib, r1, key, k, strand="unstranded", env=None):
    lo = O.Library(lib, k=k, strand_filter=strand)
    ro, fo = O.align(lo, to_concat(r1))
    old = {}
    for name, v in (env or {}).items():
        old[name] = os.environ.get(name)
        os.environ[name] = v
    try:
        lg = engine.load_library(lib, strand_filter=strand, k=k)
    finally:
        for name, v in old.items():
            if v is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = v
    assert lg.feature_names == lo.features
    table, rg, fg = engine.align(lg, r1, key=key, per_read=True)
    bad = diff_results(ro, fo, rg, fg)
    assert not bad, "\n".join(bad)
    cell, cnt, off, ids, dropped = oracle_counts(lo, ro, fo, key, 0.05)
    assert table_tuple(table.cell, table.count, table.feat_off, table.feat_ids) == table_tuple(cell, cnt, off, ids)
    assert table.dropped_empty == dropped
    return lg, ro, table


def test_cfg4_shape_combined_union_library(engine):
    """MHC ∪ KIR ∪ transcripts, 6 530 references, intersect_level 0, 60 k reads + CB/UB."""
    lib, codes = synth.combined_library(n_transcripts=3000, seed=4)
    assert len(lib[1]["columns"][1]) >= 6000 and lib[0]["intersect_level"] == 0
    r1, truth = synth.sample_reads(codes, 60_000, read_len=90, err_rate=0.005, off_target=0.2, rc_frac=0.1, seed=4)
    key = synth.barcodes_10x(len(r1), n_cells=400, seed=4, truth=truth)
    lg, ro, table = both(engine, lib, r1, key, k=20)
    assert lg.info["n_refs"] == len(lib[1]["columns"][1])
    assert (ro["reason"] == 0).sum() > 20_000 and ro["n_sw"].sum() > 5_000 and len(table) > 400


@pytest.mark.parametrize("lf", [None, "0.6", "0.9"])
def test_cfg5_shape_transcript_library_k31(engine, lf):
    """Transcript families, k = 31, score_percent 0.25; the table is also forced to the dense layouts the
    200 k-transcript library gets (load factor 0.6) and beyond (0.9: long displacement chains)."""
    lib, codes = synth.random_transcript_library(n_seqs=4000, mean_len=2000, family_frac=0.3, seed=5,
                                                 config={"score_percent": 0.25})
    r1, truth = synth.sample_reads(codes, 50_000, read_len=100, err_rate=0.005, off_target=0.2, rc_frac=0.1, seed=6)
    # every tenth read is a chimera (last 58 bases random): aligns over ~42 bp, callable only under the lenient bar
    rng = np.random.default_rng(7)
    chim = np.arange(0, len(r1), 10)
    r1[chim, 42:] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(len(chim), 58))]
    key = synth.barcodes_10x(len(r1), n_cells=300, seed=6, truth=truth)
    lg, ro, table = both(engine, lib, r1, key, k=31, env={"NB200_TABLE_LF": lf} if lf else None)
    assert lg.info["n_kmers"] > 5_000_000
    assert (ro["reason"] == 0).sum() > 25_000 and len(table) > 300
    # lenient score_percent is live: some called reads sit below the default 0.5 bar
    frac = ro["score"].max(axis=1) / 100.0
    assert ((ro["reason"] == 0) & (frac < 0.5)).sum() > 0


def test_cfg5_shape_wide_path_on_transcripts(engine):
    """Same shape with the shared-memory class lists disabled: every read takes the wide (global scratch) kernel."""
    lib, codes = synth.random_transcript_library(n_seqs=600, mean_len=1500, family_frac=0.5, seed=15,
                                                 config={"score_percent": 0.25})
    r1, truth = synth.sample_reads(codes, 6_000, read_len=100, err_rate=0.01, off_target=0.2, rc_frac=0.3, seed=16)
    key = synth.barcodes_10x(len(r1), n_cells=40, seed=16, truth=truth)
    both(engine, lib, r1, key, k=31, env={"NB200_NARROW_CAP": "0"})
