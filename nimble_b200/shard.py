"""Cell-barcode sharding across GPUs: a read goes to rank hash(CB) % world, so
every (cell, umi) group — the only cross-read coupling of the path (nimble/__main__.py:251,
nimble/utils.py:191) — stays rank-local and no data-path collective is needed.  The per-rank
count tables are gathered once at the end."""
from __future__ import annotations

import numpy as np

_M = np.uint64(0x9E3779B97F4A7C15)


def shard_of_key(key, world):
    """key: uint64 array (cell << 32 | umi) -> rank per read.  Reads without a barcode go to rank 0."""
    key = np.asarray(key, np.uint64)
    cell = key >> np.uint64(32)
    h = (cell * _M) >> np.uint64(40)
    out = (h % np.uint64(max(1, world))).astype(np.int32)
    out[key == np.uint64(0xFFFFFFFFFFFFFFFF)] = 0
    return out


def split_by_rank(key, world):
    """Index arrays, one per rank, preserving input order inside a rank."""
    r = shard_of_key(key, world)
    return [np.nonzero(r == i)[0] for i in range(world)]


def table_to_tensor_rows(table, width):
    """CountTable -> int64 matrix [n, 3 + width]: cell, count, n_feat, ids (-1 padded)."""
    n = len(table)
    m = np.full((n, 3 + width), -1, np.int64)
    if n:
        nf = (table.feat_off[1:].astype(np.int64) - table.feat_off[:-1].astype(np.int64))
        m[:, 0] = table.cell
        m[:, 1] = table.count
        m[:, 2] = nf
        cols = np.arange(width)[None, :]
        mask = cols < nf[:, None]
        m[:, 3:][mask] = table.feat_ids.astype(np.int64)
    return m


def merge_rank_tables(mats, tok_end=None, tok_comma=None):
    """Concatenate per-rank matrices (cells are disjoint across ranks by construction) and order
    rows by (cell, feature-string order).  With token ranks absent, ids order is used."""
    mats = [m for m in mats if len(m)]
    if not mats:
        return np.zeros((0, 3), np.int64)
    m = np.concatenate(mats, axis=0)
    width = m.shape[1] - 3
    keys = []
    for p in range(width - 1, -1, -1):
        ids = m[:, 3 + p]
        if tok_end is not None:
            last = (m[:, 2] - 1) == p
            t = np.where(ids < 0, 0, np.where(last, np.asarray(tok_end)[np.maximum(ids, 0)], np.asarray(tok_comma)[np.maximum(ids, 0)]) + 1)
        else:
            t = ids + 1
        keys.append(t)
    keys.append(m[:, 0])
    order = np.lexsort(keys)
    return m[order]
