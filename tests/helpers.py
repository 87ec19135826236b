"""Shared helpers for the parity tests: run oracle and CUDA path on the same inputs and diff."""
import numpy as np

from nimble_b200 import synth
from oracle import oracle as O

FIELDS = ["score", "n_hits", "n_cand", "edits", "status", "reason", "config", "n_feat", "n_sw", "pair_score"]


def ascii_matrix_to_concat(m):
    m = np.ascontiguousarray(m, np.uint8)
    off = np.arange(0, m.size + 1, m.shape[1], dtype=np.int64) if m.shape[1] else np.zeros(m.shape[0] + 1, np.int64)
    return m.reshape(-1), off


def to_concat(reads):
    if isinstance(reads, np.ndarray):
        return ascii_matrix_to_concat(reads)
    off = np.zeros(len(reads) + 1, np.int64)
    if len(reads):
        np.cumsum([len(r) for r in reads], out=off[1:])
    return np.frombuffer("".join(reads).encode("ascii"), np.uint8), off


def diff_results(res_o, feats_o, res_g, feats_g, limit=5):
    """Returns a list of human-readable mismatches (empty == bit-exact)."""
    bad = []
    for f in FIELDS:
        a, b = res_o[f], res_g[f]
        if not np.array_equal(a, b):
            idx = np.nonzero((a != b).reshape(len(a), -1).any(axis=1))[0]
            for i in idx[:limit]:
                bad.append("read %d field %s oracle=%s gpu=%s | oracle row=%s | gpu row=%s" % (i, f, a[i], b[i], res_o[i], res_g[i]))
    if not np.array_equal(feats_o, feats_g):
        idx = np.nonzero((feats_o != feats_g).any(axis=1))[0]
        for i in idx[:limit]:
            bad.append("read %d feats oracle=%s gpu=%s" % (i, feats_o[i], feats_g[i]))
    return bad


def oracle_counts(lib_o, res_o, feats_o, key, threshold=0.05, disable=False):
    """A6 through the C oracle on the oracle's own per-read calls."""
    n = len(res_o)
    nf = res_o["n_feat"].astype(np.int64)
    off = np.zeros(n + 1, np.int32)
    np.cumsum(nf, out=off[1:])
    mask = np.arange(feats_o.shape[1])[None, :] < nf[:, None]
    ids = feats_o[mask].astype(np.uint32)
    if key is None:
        key = np.zeros(n, np.uint64)   # bulk: one (cell 0) group per read handled by caller
    return O.a6_ids(key, off, ids, None, lib_o.tok_end, lib_o.tok_comma, threshold, disable)


def table_tuple(cell, count, off, ids):
    return [(int(cell[i]), tuple(int(x) for x in ids[off[i]:off[i + 1]]), int(count[i])) for i in range(len(cell))]
