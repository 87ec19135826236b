"""Generates tests/golden/a6_pandas_cases.json by running the REFERENCE's own pandas code
(/root/reference/nimble/__main__.py: report(), nimble/utils.py) on seeded random per-read
TSVs.  Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_a6_golden.py

Two compatibility shims are needed to import/run the unmodified reference here; neither
changes its arithmetic:
  * pysam / Bio / seaborn / matplotlib / jinja2 are absent -> stubbed with MagicMock modules
    (only imported at module top level, never touched by report()).
  * pandas 3.0.2 no longer passes grouping columns to groupby.apply (reference pins
    pandas==1.5.3, requirements.txt:16) -> DataFrameGroupBy.apply is wrapped so the applied
    function receives the group's rows including the key columns, as pandas 1.5 did.
With these shims all 24 tests of /root/reference/test/test.py pass (both shadowed classes).
"""
import csv, io, json, os, random, sys, tempfile, contextlib
from unittest import mock

REF = "/root/reference"
sys.path.insert(0, REF)
for m in ["pysam", "Bio", "Bio.SeqIO", "Bio.Entrez", "seaborn", "matplotlib", "matplotlib.pyplot",
          "jinja2", "Bio.Seq", "Bio.SeqRecord", "matplotlib.colors", "matplotlib.patches"]:
    sys.modules.setdefault(m, mock.MagicMock())

import pandas as pd
from pandas.core.groupby.generic import DataFrameGroupBy

_orig_apply = DataFrameGroupBy.apply


def _apply_with_keys(self, func, *args, **kwargs):
    obj = self.obj
    return _orig_apply(self, lambda g: func(obj.loc[g.index], *args, **kwargs))


DataFrameGroupBy.apply = _apply_with_keys

import importlib
ref_main = importlib.import_module("nimble.__main__")

FEATURE_POOLS = [
    ["A", "B", "C", "D", "E"],
    ["HLA-A", "HLA-A*01:01", "HLA-A*02:01", "HLA-B", "HLA-B*07", "HLA-A 2"],
    ["Mamu-A1*001", "Mamu-A1*001:01", "Mamu-A1", "Mamu-B*017", "Mamu-B", "Mamu-B!x", "Mamu-B-1"],
    ["KIR2DL1", "KIR2DL1*001", "KIR2DL10", "KIR2DL1.2", "KIR3DL1", "a", "Z"],
]
THRESHOLDS = [0.05, 0.05, 0.05, 0.1, 0.2, 0.25, 0.5, 1.0 / 3.0, 0.0, 0.15]


def make_case(rng, idx):
    """Each (cell, umi) has a true feature carried by most of its rows, plus ambiguity and noise."""
    pool = rng.choice(FEATURE_POOLS)
    n_cells = rng.randint(1, 4)
    n_umis = rng.randint(1, 5)
    n_rows = rng.randint(1, 40)
    float_scores = rng.random() < 0.25
    truth = {}
    rows = []
    for _ in range(n_rows):
        ci, ui = rng.randint(0, n_cells - 1), rng.randint(0, n_umis - 1)
        t = truth.setdefault((ci, ui), rng.choice(pool))
        cb = "CELL%d" % ci if rng.random() > 0.03 else ""
        umi = "UMI%d" % ui if rng.random() > 0.03 else ""
        k = rng.choice([0, 0, 0, 1, 1, 2, 3])
        feats = rng.sample([p for p in pool if p != t], min(k, len(pool) - 1))
        if rng.random() < 0.85:
            feats.append(t)
        if not feats:
            feats = [rng.choice(pool)]
        rng.shuffle(feats)
        f = ",".join(feats) if rng.random() > 0.03 else ""
        if float_scores:
            score = rng.choice([0.5, 0.25, 1.5, 2.0, 0.1, 0.3, 1.0, 3.0])
        else:
            score = rng.choice([1, 1, 1, 1, 2, 3, 5, 19, 20])
        rows.append([f, score, cb, umi])
    return {"id": idx, "threshold": rng.choice(THRESHOLDS),
            "disable_thresholding": rng.random() < 0.1, "rows": rows}


def boundary_case(rng, idx):
    """One UMI whose feature X sits EXACTLY on the threshold through fractional shares
    (1/2+1/3+1/6, 1/4+1/4+1/2, ...), so Kahan-vs-plain summation and row order decide whether
    X survives; every row carries X, so the final count row shows the decision."""
    others = ["B", "C", "D", "E", "F", "G", "H", "I", "J"]
    recipe = rng.choice([[2, 3, 6], [3, 3, 3], [4, 4, 2], [6, 6, 6, 2], [5, 5, 5, 5, 5], [2, 6, 6, 6],
                         [4, 4, 4, 4], [2, 4, 8, 8], [3, 6, 6, 3, 3][:rng.choice([3, 5])]])
    mult = rng.choice([1, 1, 2, 3])
    x = rng.choice(["X", "A", "Z"])
    rows = []
    for n in recipe:
        feats = [x] + rng.sample(others, n - 1)
        rng.shuffle(feats)
        for _ in range(mult):
            rows.append([",".join(feats), 1, "CELL0", "UMI0"])
    rng.shuffle(rows)
    thr = 1.0 / len(recipe) if sum(1.0 / n for n in recipe) > 0.99 and abs(sum(1.0 / n for n in recipe) - 1) < 1e-9 else 0.2
    return {"id": idx, "threshold": thr, "disable_thresholding": False, "rows": rows}


KAHAN_ORDERS = [[3, 4, 4, 6], [3, 3, 6, 6]]  # 4 rows -> total 4 -> ratio = sum/4 exactly


def kahan_case(rng, idx):
    """Shares 1/d in an order where plain left-to-right summation gives 0.9999999999999999 but
    pandas' Kahan group_sum gives 1.0; threshold = 1/len so X survives only with Kahan.  Row
    order inside the UMI is the sorted feature-string order, forced with leading names."""
    dens = rng.choice(KAHAN_ORDERS)
    lead = ["A", "B", "C", "D", "E", "F"]
    rows = []
    for i, d in enumerate(dens):
        extras = [lead[i]] + ["%s%d" % (lead[i], j) for j in range(d - 2)]
        feats = extras + ["X"]
        rng.shuffle(feats)
        rows.append([",".join(feats), 1, "CELL0", "UMI0"])
    rng.shuffle(rows)
    return {"id": idx, "threshold": 1.0 / len(dens), "disable_thresholding": False, "rows": rows}


def run_reference(case, tmp):
    inp = os.path.join(tmp, "in.tsv")
    out = os.path.join(tmp, "out.tsv")
    with open(inp, "w") as f:
        f.write("nimble_features\tnimble_score\tr1_CB\tr1_UB\n")
        for feats, score, cb, umi in case["rows"]:
            f.write("%s\t%s\t%s\t%s\n" % (feats, repr(score) if isinstance(score, float) else score, cb, umi))
    if os.path.exists(out):
        os.remove(out)
    buf = io.StringIO()
    err = None
    try:
        with contextlib.redirect_stdout(buf):
            ref_main.report(inp, out, None, case["threshold"], case["disable_thresholding"])
    except Exception as e:  # reference crashes when thresholding removes every row (KeyError)
        err = "%s: %s" % (type(e).__name__, e)
    text = open(out).read() if (err is None and os.path.exists(out)) else None
    return text, buf.getvalue(), err


def main():
    rng = random.Random(20261018)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for i in range(260):
            c = kahan_case(rng, i) if i % 20 == 10 else (make_case(rng, i) if i % 4 else boundary_case(rng, i))
            text, log, err = run_reference(c, tmp)
            c["expected_tsv"] = text
            c["reference_stdout"] = log
            c["reference_error"] = err
            cases.append(c)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "a6_pandas_cases.json")
    with open(dst, "w") as f:
        json.dump({"generator": "tests/golden/make_a6_golden.py", "pandas": pd.__version__,
                   "reference": "nimble/__main__.py report() @ /root/reference", "cases": cases}, f)
    print("wrote", dst, len(cases), "cases")


if __name__ == "__main__":
    main()


def format_goldens():
    """Format oracles that need no aligner: the default config JSON (nimble/types.py:12-25 dumped as
    nimble/__main__.py:64-65 does) and append_path_string / library-name answers (nimble/utils.py:9-32)."""
    from nimble.types import Config, Data
    from nimble.utils import append_path_string, get_library_name_from_filename
    paths = [("out.tsv", ".mhc"), ("/a/b/out.tsv.gz", ".kir_lib"), ("out", ".x"), ("dir.v1/out.tsv", ".l"),
             ("./rel/out.counts.tsv.gz", ".a b"), ("out.tsv", "")]
    names = ["/x/y/my_mhc_library.fasta", "lib.csv", "a_b_c.d.json", "noext"]
    g = {"library_json_empty": json.dumps([Config().__dict__, Data().__dict__], indent=2),
         "append_path_string": [[p, a, append_path_string(p, a)] for p, a in paths],
         "library_name": [[n, get_library_name_from_filename(n)] for n in names]}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "format_goldens.json")
    with open(dst, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    format_goldens()
