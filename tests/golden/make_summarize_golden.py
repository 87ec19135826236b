"""Generates tests/golden/summarize_cases.json by running the REFERENCE's summarize_fields
(/root/reference/nimble/__main__.py:295-297, fed by convert_df_to_proper_umi :234-252 through
check_df_from_input :213-232) on seeded per-read TSVs.  Build container only.

    python tests/golden/make_summarize_golden.py

Shims: the import stubs of make_a6_golden.py, plus DataFrame.applymap (removed in pandas 3;
the reference pins pandas 1.5.3) aliased to DataFrame.map, which is the same element-wise call.
"""
import io, json, os, random, sys, tempfile, contextlib
from unittest import mock

REF = "/root/reference"
sys.path.insert(0, REF)
for m in ["pysam", "Bio", "Bio.SeqIO", "Bio.Entrez", "seaborn", "matplotlib", "matplotlib.pyplot",
          "jinja2", "Bio.Seq", "Bio.SeqRecord", "matplotlib.colors", "matplotlib.patches"]:
    sys.modules.setdefault(m, mock.MagicMock())
import pandas as pd

if not hasattr(pd.DataFrame, "applymap"):
    pd.DataFrame.applymap = pd.DataFrame.map
import importlib

ref_main = importlib.import_module("nimble.__main__")

COLS = ["r1_GN", "r1_POS", "r2_POS", "r1_forward_score", "r1_CB", "qname"]


def make_case(rng, idx):
    n = rng.randint(1, 30)
    n_umi = rng.randint(1, 5)
    rows = []
    int_pos = rng.random() < 0.5           # an all-integer column stays int; one blank makes it float
    for i in range(n):
        umi = "UMI%d" % rng.randint(0, n_umi - 1) if rng.random() > 0.08 else ""
        gn = rng.choice(["HLA-A", "HLA-B", "KIR2DL1", "", "gene with space", "A;B"])
        pos1 = str(rng.choice([1, 5, 5, 100, 2500]))
        pos2 = str(rng.choice([7, 7, 9])) if (int_pos or rng.random() > 0.2) else ""
        fs = rng.choice(["20", "35.5", "90", "90.0", ""])
        cb = rng.choice(["CELL1", "CELL2", ""])
        rows.append([rng.choice(["A", "A,B", "B", ""]), str(rng.choice([1, 1, 2])), cb, umi, gn, pos1, pos2, fs, "q%d" % rng.randint(0, 6)])
    k = rng.randint(1, 4)
    cols = rng.sample(COLS, k)
    return {"id": idx, "rows": rows, "columns": cols}


HEADER = ["nimble_features", "nimble_score", "r1_CB", "r1_UB", "r1_GN", "r1_POS", "r2_POS", "r1_forward_score", "qname"]


def run_reference(case, tmp):
    inp = os.path.join(tmp, "in.tsv")
    with open(inp, "w") as f:
        f.write("\t".join(HEADER) + "\n")
        for r in case["rows"]:
            f.write("\t".join(r) + "\n")
    df = ref_main.check_df_from_input(inp, os.path.join(tmp, "unused.tsv"))
    if df is None:
        return None
    _, df_init = ref_main.convert_df_to_proper_umi(df)
    cols = ["cb" if c == "r1_CB" else c for c in case["columns"]]      # report() renames r1_CB -> cb before summarising
    out = os.path.join(tmp, "summary.tsv")
    try:
        ref_main.summarize_fields(df_init, cols, out)
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}
    return {"text": open(out).read(), "columns_after_rename": cols}


def main():
    rng = random.Random(77)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for i in range(60):
            c = make_case(rng, i)
            with contextlib.redirect_stdout(io.StringIO()):
                c["expected"] = run_reference(c, tmp)
            cases.append(c)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "summarize_cases.json")
    with open(path, "w") as f:
        json.dump({"header": HEADER, "cases": cases}, f, separators=(",", ":"))
    print("wrote %d cases, %d errors, %d bytes" % (len(cases), sum(1 for c in cases if c["expected"] and "error" in c["expected"]), os.path.getsize(path)))
    print(cases[0]["columns"], cases[0]["expected"])


if __name__ == "__main__":
    main()
