"""ORACLE python wrapper (test infrastructure — see oracle/nimble_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module.  nimble_b200/ never does.
"""
from __future__ import annotations

import ctypes as ct
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc.so")

RESULT_DTYPE = np.dtype([
    ("score", "<u2", (4,)), ("n_hits", "<u2", (4,)), ("n_cand", "<u2", (4,)),
    ("edits", "u1", (4,)), ("status", "u1", (4,)),
    ("reason", "u1"), ("config", "u1"), ("n_feat", "u1"), ("n_sw", "u1"), ("pair_score", "<u4"),
])

STRAND = {"unstranded": 0, "fiveprime": 1, "threeprime": 2, "none": 3}


class OrcConfig(ct.Structure):
    _fields_ = [("k", ct.c_int32), ("score_threshold", ct.c_int32), ("score_filter", ct.c_int32),
                ("score_percent", ct.c_double), ("num_mismatches", ct.c_int32),
                ("discard_multiple_matches", ct.c_int32), ("intersect_level", ct.c_int32),
                ("discard_multi_hits", ct.c_int32), ("require_valid_pair", ct.c_int32),
                ("max_hits_to_report", ct.c_int32), ("strand_filter", ct.c_int32), ("pad_", ct.c_int32)]


def build(force=False):
    """Compile oracle/nimble_oracle.c -> oracle/_build/liborc.so (gcc, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("nimble_oracle.c", "cb_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/liborc.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ct.CDLL(_SO)
        L.orc_index_build.restype = ct.c_void_p
        L.orc_index_build.argtypes = [ct.c_int32, ct.POINTER(ct.c_char_p), ct.c_void_p, ct.c_int32, ct.c_int32]
        L.orc_index_free.argtypes = [ct.c_void_p]
        for f in ("orc_index_n_kmers", "orc_index_n_classes", "orc_index_class_members"):
            getattr(L, f).restype = ct.c_int64
            getattr(L, f).argtypes = [ct.c_void_p]
        L.orc_align.restype = ct.c_int32
        L.orc_align.argtypes = [ct.c_void_p, ct.POINTER(OrcConfig), ct.c_int64, ct.c_void_p, ct.c_void_p,
                                ct.c_void_p, ct.c_void_p, ct.c_int32, ct.c_void_p, ct.c_void_p]
        L.orc_a6.restype = ct.c_int64
        L.orc_a6.argtypes = [ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p,
                             ct.c_void_p, ct.c_double, ct.c_int32,
                             ct.POINTER(ct.c_void_p), ct.POINTER(ct.c_void_p), ct.POINTER(ct.c_void_p),
                             ct.POINTER(ct.c_void_p), ct.POINTER(ct.c_int64)]
        L.orc_a6_mt.restype = ct.c_int64
        L.orc_a6_mt.argtypes = [ct.c_int32] + list(L.orc_a6.argtypes)
        L.orc_free.argtypes = [ct.c_void_p]
        L.orc_cb_correct.restype = ct.c_int32
        L.orc_cb_correct.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int32, ct.c_void_p, ct.c_void_p,
                                     ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p]
        assert L.orc_sizeof_result() == RESULT_DTYPE.itemsize
        assert L.orc_sizeof_config() == ct.sizeof(OrcConfig)
        _lib = L
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


# ---- library JSON (nimble/__main__.py:64-65, nimble/types.py:10-32) -------------------------------
def token_ranks(names):
    """names: feature names in id order (ascending by str).  Returns (tok_end, tok_comma) ranks:
    order of `name + NUL` / `name + ','` among all 2F tokens == order of comma-joined strings."""
    toks = []
    for i, n in enumerate(names):
        b = n.encode("utf-8")
        toks.append((b + b"\x00", 0, i))
        toks.append((b + b",", 1, i))
    toks.sort(key=lambda t: t[0])
    te = np.zeros(len(names), np.uint32)
    tc = np.zeros(len(names), np.uint32)
    for rank, (_, kind, i) in enumerate(toks):
        (te if kind == 0 else tc)[i] = rank
    return te, tc


class Library:
    """Parsed `[config, data]` library JSON, with features resolved through `group_on`."""

    def __init__(self, obj, k=20, strand_filter="unstranded"):
        if isinstance(obj, (str, os.PathLike)):
            with open(obj) as f:
                obj = json.load(f)
        cfg, data = obj[0], obj[1]
        headers = data["headers"]
        cols = data["columns"]
        self.names = list(cols[headers.index("sequence_name")])
        self.seqs = list(cols[headers.index("sequence")])
        group_on = cfg.get("group_on", "") or ""
        if group_on and group_on in headers:
            feat_names = list(cols[headers.index(group_on)])
        else:
            feat_names = self.names
        self.features = sorted(set(feat_names))             # id = rank by name (code point == utf-8 byte order)
        fid = {n: i for i, n in enumerate(self.features)}
        self.ref_feature = np.array([fid[n] for n in feat_names], np.int32)
        self.tok_end, self.tok_comma = token_ranks(self.features)
        self.cfg = OrcConfig(
            k=k, score_threshold=int(cfg.get("score_threshold", 20)), score_filter=int(cfg.get("score_filter", 25)),
            score_percent=float(cfg.get("score_percent", 0.5)), num_mismatches=int(cfg.get("num_mismatches", 0)),
            discard_multiple_matches=int(bool(cfg.get("discard_multiple_matches", False))),
            intersect_level=int(cfg.get("intersect_level", 0)), discard_multi_hits=int(cfg.get("discard_multi_hits", 0)),
            require_valid_pair=int(bool(cfg.get("require_valid_pair", False))),
            max_hits_to_report=int(cfg.get("max_hits_to_report", 10)), strand_filter=STRAND[strand_filter], pad_=0)
        self.k = k
        self._index = None

    @property
    def index(self):
        if self._index is None:
            self._index = Index(self.seqs, self.ref_feature, len(self.features), self.k)
        return self._index


class Index:
    def __init__(self, seqs, ref_feature, n_features, k):
        L = lib()
        arr = (ct.c_char_p * len(seqs))(*[s.encode("ascii", "replace") for s in seqs])
        rf = np.ascontiguousarray(ref_feature, np.int32)
        self.ptr = L.orc_index_build(len(seqs), arr, rf.ctypes.data, int(n_features), int(k))
        if not self.ptr:
            raise ValueError("orc_index_build failed (k out of range?)")
        self.n_kmers = L.orc_index_n_kmers(self.ptr)
        self.n_classes = L.orc_index_n_classes(self.ptr)
        self.n_class_members = L.orc_index_class_members(self.ptr)

    def __del__(self):
        try:
            if self.ptr:
                lib().orc_index_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _concat(reads):
    if isinstance(reads, tuple):   # already (bytes/ndarray buffer, offsets)
        return reads
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    buf = "".join(reads).encode("ascii")
    return np.frombuffer(buf, np.uint8), off


def align(library: Library, r1, r2=None, n_threads=0, cfg=None):
    """r1/r2: list[str] or (uint8 buffer, int64 offsets).  Returns (results[RESULT_DTYPE], feats[n, max_hits])."""
    L = lib()
    b1, o1 = _concat(r1)
    n = len(o1) - 1
    b1 = np.ascontiguousarray(b1, np.uint8)
    if r2 is not None:
        b2, o2 = _concat(r2)
        b2 = np.ascontiguousarray(b2, np.uint8)
        p2, po2 = b2.ctypes.data, o2.ctypes.data
    else:
        p2 = po2 = None
    cfg = cfg or library.cfg
    out = np.zeros(n, RESULT_DTYPE)
    feats = np.full((n, cfg.max_hits_to_report), -1, np.int32)
    rc = L.orc_align(library.index.ptr, ct.byref(cfg), n, b1.ctypes.data, o1.ctypes.data, p2, po2,
                     int(n_threads), out.ctypes.data, feats.ctypes.data)
    if rc != 0:
        raise ValueError("orc_align rc=%d (read longer than 500 bases or bad config)" % rc)
    return out, feats


def a6_ids(key, off, ids, score, tok_end, tok_comma, threshold=0.05, disable_thresholding=False, n_threads=1):
    """id-based A6 (C).  Returns (cell[u4], count[u4], off[i4], ids[u4], dropped_empty).
    n_threads != 1: cells dealt to that many host threads (0 = all), same result (orc_a6_mt)."""
    L = lib()
    key = np.ascontiguousarray(key, np.uint64)
    off = np.ascontiguousarray(off, np.int32)
    ids = np.ascontiguousarray(ids, np.uint32)
    sp = None
    if score is not None:
        score = np.ascontiguousarray(score, np.float64)
        sp = score.ctypes.data
    te = np.ascontiguousarray(tok_end, np.uint32)
    tc = np.ascontiguousarray(tok_comma, np.uint32)
    oc, on, oo, oi = ct.c_void_p(), ct.c_void_p(), ct.c_void_p(), ct.c_void_p()
    dropped = ct.c_int64(0)
    args = (len(key), key.ctypes.data, off.ctypes.data, ids.ctypes.data if len(ids) else None, sp,
            te.ctypes.data if len(te) else None, tc.ctypes.data if len(tc) else None,
            float(threshold), int(bool(disable_thresholding)),
            ct.byref(oc), ct.byref(on), ct.byref(oo), ct.byref(oi), ct.byref(dropped))
    n = L.orc_a6(*args) if n_threads == 1 else L.orc_a6_mt(int(n_threads), *args)
    def take(p, count, dt):
        a = np.ctypeslib.as_array(ct.cast(p, ct.POINTER(ct.c_uint8)), shape=(max(count, 1) * np.dtype(dt).itemsize,))
        r = a[:count * np.dtype(dt).itemsize].view(dt).copy()
        L.orc_free(p)
        return r
    o_off = take(oo, n + 1, np.int32)
    cell = take(oc, n, np.uint32)
    cnt = take(on, n, np.uint32)
    o_ids = take(oi, int(o_off[-1]) if n >= 0 else 0, np.uint32)
    return cell, cnt, o_off, o_ids, int(dropped.value)


def a6_strings(rows, threshold=0.05, disable_thresholding=False):
    """Same contract as oracle.a6_py.report_counts but through the C id path: rows =
    (cb, umi, features_string, score).  Returns ([(feature_string, count, cb)], dropped_empty)."""
    import math
    clean = []
    for cb, umi, feats, score in rows:
        bad = any(x is None or (isinstance(x, float) and math.isnan(x)) for x in (cb, umi, feats, score))
        if bad or cb == "" or umi == "" or feats == "":
            continue
        clean.append((cb, umi, feats.split(","), float(score)))
    names = sorted({f for r in clean for f in r[2]})
    fid = {n: i for i, n in enumerate(names)}
    cbs = sorted({r[0] for r in clean})
    cid = {n: i for i, n in enumerate(cbs)}
    uid = {n: i for i, n in enumerate(sorted({r[1] for r in clean}))}
    te, tc = token_ranks(names)
    key = np.array([(cid[r[0]] << 32) | uid[r[1]] for r in clean], np.uint64)
    off = np.zeros(len(clean) + 1, np.int32)
    ids = []
    for i, r in enumerate(clean):
        l = sorted(fid[f] for f in r[2])
        ids.extend(l)
        off[i + 1] = len(ids)
    score = np.array([r[3] for r in clean], np.float64)
    cell, cnt, o_off, o_ids, dropped = a6_ids(key, off, np.array(ids, np.uint32), score, te, tc, threshold, disable_thresholding)
    out = []
    for i in range(len(cell)):
        out.append((",".join(names[j] for j in o_ids[o_off[i]:o_off[i + 1]]), int(cnt[i]), cbs[cell[i]]))
    return out, dropped


# ---- A5: whitelist cell-barcode correction (nimble/fastq_barcode_processor.py:73-128) ---------------
def cb_correct(whitelist, cb, qual, eligible=None, cb_length=16):
    """C restatement (oracle/cb_oracle.c).  whitelist: list of str (all lines of the file);
    cb, qual: uint8 arrays n x cb_length (ASCII bases / phred values).  Returns (idx int32[n] = index
    into `whitelist` of the corrected barcode or -1, status uint8[n], stats dict)."""
    same = [(i, w) for i, w in enumerate(whitelist) if len(w) == cb_length]
    wl = np.frombuffer("".join(w for _, w in same).encode("latin-1"), np.uint8).copy() if same else np.zeros(0, np.uint8)
    wl_idx = np.array([i for i, _ in same], np.int32)
    cb = np.ascontiguousarray(cb, np.uint8).reshape(-1, cb_length)
    qual = np.ascontiguousarray(qual, np.uint8).reshape(-1, cb_length)
    n = cb.shape[0]
    el = None if eligible is None else np.ascontiguousarray(eligible, np.uint8)
    idx = np.empty(n, np.int32)
    status = np.empty(n, np.uint8)
    st = np.zeros(4, np.int64)
    rc = lib().orc_cb_correct(wl.ctypes.data, wl_idx.ctypes.data, len(same), cb_length, cb.ctypes.data, qual.ctypes.data,
                              None if el is None else el.ctypes.data, n, idx.ctypes.data, status.ctypes.data, st.ctypes.data)
    if rc != 0:
        raise ValueError("orc_cb_correct failed")
    return idx, status, {"cb_perfect_match": int(st[0]), "cb_corrected": int(st[1]), "cb_no_correction": int(st[2]),
                         "cache_size": int(st[3])}
