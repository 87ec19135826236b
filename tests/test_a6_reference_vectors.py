"""The 24 known-answer tests of the reference (test/test.py, both shadowed classes), lifted as
data and run against the oracle restatements (not gpu) and the CUDA A6 path (gpu).

Each vector cites the reference test it comes from (file:line)."""
import numpy as np
import pytest

from oracle import a6_py
from oracle import oracle as O


def per_umi_thresholding(rows, threshold):
    """Restated nimble/utils.py:119-207 on top of oracle.a6_py.threshold_group.
    rows: (cb, umi, features, score) in frame order -> [(features, filtered_features)] (non-empty only)."""
    groups = {}
    for cb, umi, f, s in rows:
        groups.setdefault((cb, umi), []).append((f, float(s)))
    out = []
    for key in sorted(groups):
        keep = a6_py.threshold_group(groups[key], threshold)
        for f, _ in groups[key]:
            k = sorted(set(f.split(",")) & keep)
            if k:
                out.append((f, ",".join(k)))
    return out


THRESHOLDING = [
    # (reference test.py lines, rows, threshold, expected [(features, filtered)])
    ("10-30 basic", [("cell1", "UMI1", "A,B", 10), ("cell1", "UMI1", "A,C", 20)], 0.2, [("A,B", "A"), ("A,C", "A,C")]),
    ("32-45 all below", [("cell1", "UMI1", "A,B,C", 3)], 0.4, []),
    ("47-60 single", [("cell1", "UMI1", "A", 10)], 0.9, [("A", "A")]),
    ("195-207 high thr", [("cell1", "UMI1", "A,B", 100), ("cell1", "UMI1", "A,B,C,D", 100)], 0.3, [("A,B", "A,B"), ("A,B,C,D", "A,B")]),
    ("209-222 non-uniform", [("cell1", "UMI1", "A", 80), ("cell1", "UMI1", "B,C", 20)], 0.25, [("A", "A")]),
    ("324-337 tie at threshold kept", [("cell1", "UMI1", "A,B", 10)], 0.5, [("A,B", "A,B")]),
    ("339-354 zero scores", [("cell1", "UMI1", "A,B", 0), ("cell1", "UMI1", "C,D", 20)], 0.1, [("C,D", "C,D")]),
    ("414-429 duplicate names", [("cell1", "UMI1", "A,A,B", 15)], 0.2, [("A,A,B", "A,B")]),
    ("481-495 1e12", [("cell1", "UMI1", "A,B,C", 1e12), ("cell1", "UMI1", "C,D,E", 1e12)], 0.2, [("A,B,C", "C"), ("C,D,E", "C")]),
    ("497-511 decimals", [("cell1", "UMI1", "A,B", 0.6), ("cell1", "UMI1", "A,C", 0.4)], 0.5, [("A,B", "A"), ("A,C", "A")]),
]

# test.py:243-259 asserts only the union of surviving names
COMPLEX = ([("cell1", "UMI1", "A,B", 10), ("cell1", "UMI1", "A,C", 15), ("cell1", "UMI1", "B,C,D", 5),
            ("cell1", "UMI1", "D,E", 20)], 0.2, {"A", "E", "D"})

INTERSECTIONS = [
    ("62-78", [["A", "B"], ["A", "C"], ["A", "D"]], ["A"]),
    ("80-91", [["A", "B"], ["C", "D"]], []),
    ("93-98 empty input", [], []),
    ("100-105", [["A", "B", "C"]], ["A", "B", "C"]),
    ("224-235", [["A", "B", "C"]], ["A", "B", "C"]),
    ("261-272", [["A", "B", "C"], ["A", "C"], ["B", "C", "D"], ["C", "D", "E"]], ["C"]),
]

PIPELINES = [
    # (lines, rows (cb, umi, features, score), threshold, expected [(feature, count, cell)])
    ("107-138", [("cell1", "UMI1", "A,B", 10), ("cell1", "UMI1", "A,C", 20), ("cell2", "UMI2", "D,E", 30),
                 ("cell2", "UMI2", "D,F", 40), ("cell3", "UMI3", "G", 50)], 0.2,
     [("A", 1, "cell1"), ("D", 1, "cell2"), ("G", 1, "cell3")]),
    ("140-160", [("cell1", "UMI1", "A,B,C", 3)], 0.4, []),
    ("162-193", [("cell1", "UMI1", "A,B", 10), ("cell1", "UMI1", "A,B", 10)], 0.1, [("A,B", 1, "cell1")]),
    ("274-322", [("cell1", "UMI1", "A,B", 10), ("cell1", "UMI1", "A,C", 20), ("cell1", "UMI2", "B,D", 15),
                 ("cell2", "UMI3", "E,F", 5), ("cell2", "UMI3", "F,G", 35), ("cell3", "UMI4", "H,I", 25),
                 ("cell3", "UMI5", "I,J", 15), ("cell3", "UMI5", "H,J", 10)], 0.2,
     [("A", 1, "cell1"), ("B,D", 1, "cell1"), ("F", 1, "cell2"), ("H,I", 1, "cell3"), ("J", 1, "cell3")]),
    ("365-412", [("cell1", "UMI1", "A,B", 10), ("cell2", "UMI2", "C,D", 20), ("cell2", "UMI2", "D,E", 30),
                 ("cell3", "UMI3", "F,G", 40)], 0.0,
     [("A,B", 1, "cell1"), ("D", 1, "cell2"), ("F,G", 1, "cell3")]),
    ("431-479", [("cell1", "UMI1", "A", 10), ("cell1", "UMI1", "B", 5), ("cell1", "UMI2", "A,B", 8), ("cell1", "UMI2", "B,C", 12),
                 ("cell1", "UMI2", "C", 3), ("cell2", "UMI3", "D", 20), ("cell2", "UMI3", "E", 15), ("cell2", "UMI4", "F", 25),
                 ("cell2", "UMI4", "F,G", 5), ("cell3", "UMI5", "H,I", 10), ("cell3", "UMI5", "I,J", 15), ("cell3", "UMI5", "H,J", 5)],
     0.15, [("F", 1, "cell2")]),
]


@pytest.mark.parametrize("name,rows,thr,expected", THRESHOLDING, ids=[t[0] for t in THRESHOLDING])
def test_thresholding_vectors(name, rows, thr, expected):
    assert per_umi_thresholding(rows, thr) == expected


def test_thresholding_complex_scores():
    rows, thr, names = COMPLEX
    got = per_umi_thresholding(rows, thr)
    assert set(",".join(f for _, f in got).split(",")) == names


@pytest.mark.parametrize("name,lists,expected", INTERSECTIONS, ids=[t[0] for t in INTERSECTIONS])
def test_intersection_vectors(name, lists, expected):
    got = sorted(set.intersection(*map(set, lists))) if lists else []
    assert got == expected
    # the same through report_counts with thresholding disabled (one UMI, rows = the lists)
    rows = [("c", "u", ",".join(l), 1) for l in lists]
    out, dropped = a6_py.report_counts(rows, 0.05, True)
    if expected:
        assert out == [(",".join(expected), 1, "c")]
    else:
        assert out == [] and dropped == (1 if lists else 0)
    out_c, dropped_c = O.a6_strings(rows, 0.05, True)
    assert out_c == out and dropped_c == dropped


@pytest.mark.parametrize("name,rows,thr,expected", PIPELINES, ids=[t[0] for t in PIPELINES])
def test_pipeline_vectors(name, rows, thr, expected):
    out, _ = a6_py.report_counts(rows, thr, False)
    assert out == expected
    out_c, _ = O.a6_strings(rows, thr, False)
    assert out_c == expected


def all_report_inputs():
    for name, rows, thr, _ in THRESHOLDING:
        yield name, rows, thr
    yield "complex", COMPLEX[0], COMPLEX[1]
    for name, rows, thr, _ in PIPELINES:
        yield name, rows, thr


def test_c_oracle_equals_python_oracle_on_every_reference_input():
    for name, rows, thr in all_report_inputs():
        assert O.a6_strings(rows, thr, False) == a6_py.report_counts(rows, thr, False), name


def gpu_report(engine, rows, thr, disable=False):
    """rows of strings -> CUDA nb200_umi_counts -> [(feature, count, cell)]"""
    rows = [(cb, umi, ",".join(sorted(f.split(","))), float(s)) for cb, umi, f, s in rows if cb and umi and f]
    names = sorted({x for r in rows for x in r[2].split(",")})
    fid = {n: i for i, n in enumerate(names)}
    cbs = sorted({r[0] for r in rows})
    umis = sorted({r[1] for r in rows})
    lib = engine.load_feature_names(names)
    key = np.array([(cbs.index(r[0]) << 32) | umis.index(r[1]) for r in rows], np.uint64)
    off = np.zeros(len(rows) + 1, np.uint32)
    ids = []
    for i, r in enumerate(rows):
        ids.extend(sorted(fid[x] for x in r[2].split(",")))
        off[i + 1] = len(ids)
    t = engine.umi_counts(lib, key, off, np.array(ids, np.uint32), np.array([r[3] for r in rows]), thr, disable)
    return [(f, c, cbs[cell]) for f, c, cell in t.rows(names)], t.dropped_empty


@pytest.mark.gpu
def test_gpu_a6_on_every_reference_input(engine):
    for name, rows, thr in all_report_inputs():
        assert gpu_report(engine, rows, thr) == a6_py.report_counts(rows, thr, False), name
    for name, lists, expected in INTERSECTIONS:
        rows = [("c", "u", ",".join(l), 1) for l in lists]
        if rows:
            assert gpu_report(engine, rows, 0.05, True) == a6_py.report_counts(rows, 0.05, True), name


@pytest.mark.gpu
@pytest.mark.parametrize("seed,shape", [(1, "long_rows"), (2, "big_groups"), (3, "many_small"), (4, "scores")])
def test_gpu_a6_random_shapes_vs_c_oracle(engine, seed, shape):
    """nb200_umi_counts on random CSR rows vs the C oracle (itself pinned on the reference's vectors): rows of up to 5000
    names (no padding to the longest row, far past the 4096-name limit of round 1), UMIs of up to 300 rows (global token
    sort fallback, global pools of umi_general_kernel), a million tiny UMIs, fractional scores."""
    rng = np.random.default_rng(100 + seed)
    nfeat = 6000
    names = ["g%05d" % i if i % 3 else "G%d-x" % i for i in range(nfeat)]
    order = sorted(range(nfeat), key=lambda i: names[i].encode())
    names = [names[i] for i in order]
    lib = engine.load_feature_names(names)
    n_rows = {"long_rows": 3000, "big_groups": 40000, "many_small": 400000, "scores": 60000}[shape]
    if shape == "long_rows":
        lens = np.where(rng.random(n_rows) < 0.01, rng.integers(1000, 5001, n_rows), rng.integers(1, 6, n_rows))
        cells, umis = rng.integers(0, 20, n_rows), rng.integers(0, 200, n_rows)
    elif shape == "big_groups":
        lens = rng.integers(1, 5, n_rows)
        cells, umis = rng.integers(0, 4, n_rows), rng.integers(0, 40, n_rows)          # ~250 rows per UMI
    elif shape == "many_small":
        lens = rng.integers(1, 4, n_rows)
        cells, umis = rng.integers(0, 3000, n_rows), rng.integers(0, 1 << 20, n_rows)
    else:
        lens = rng.integers(1, 7, n_rows)
        cells, umis = rng.integers(0, 50, n_rows), rng.integers(0, 300, n_rows)
    off = np.zeros(n_rows + 1, np.uint32)
    np.cumsum(lens, out=off[1:])
    ids = np.empty(int(off[-1]), np.uint32)
    for i in range(n_rows):                   # ascending names per row, a few duplicates, drawn from a small pool per cell
        pool = (int(cells[i]) * 37) % (nfeat - 5200)
        r = np.sort(rng.integers(pool, pool + max(8, int(lens[i]) + 3), int(lens[i])))
        ids[off[i]:off[i + 1]] = r
    key = (cells.astype(np.uint64) << np.uint64(32)) | umis.astype(np.uint64)
    score = None if shape != "scores" else rng.choice([0.5, 1.0, 2.0, 0.1, 3.25], n_rows)
    tok_end, tok_comma = O.token_ranks(names)
    for thr, disable in ((0.05, False), (0.3, False), (0.05, True)):
        t = engine.umi_counts(lib, key, off, ids, score, thr, disable)
        cell, cnt, o_off, o_ids, dropped = O.a6_ids(key, off.astype(np.int32), ids, score, tok_end, tok_comma, thr, disable)
        assert np.array_equal(cell, t.cell) and np.array_equal(cnt, t.count), (shape, thr, disable)
        assert np.array_equal(o_off.astype(np.int64), t.feat_off.astype(np.int64)) and np.array_equal(o_ids, t.feat_ids)
        assert dropped == t.dropped_empty
