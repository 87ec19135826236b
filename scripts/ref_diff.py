#!/usr/bin/env python
"""Differential harness against the REAL nimble-aligner binary (the arithmetic of SURVEY.md §8 rows X1-X5 lives in the
un-vendored BimberLab/nimble-aligner release, `nimble/__main__.py:127-131`; it cannot be fetched offline, so parity of
those rows is pinned only to oracle/nimble_oracle.c — DESIGN.md §2).  The day a binary is at hand:

    python scripts/ref_diff.py --aligner /path/to/nimble/aligner [--reads 100000] [--out-dir /tmp/nb200_ref_diff]

1. writes two fixtures: `cfg1` (two-reference library, a few hundred reads: BASELINE.json configs[0] shape) and a
   `cfg2` slice (2000-allele MHC-like library, --reads 10x-tagged 90 bp reads), each as library JSON + BAM with CB/UB;
2. runs PATH with exactly the argv of nimble/__main__.py:177-192 (`--input BAM -c N --strand_filter unstranded
   -r LIB.json -o OUT.tsv.gz`) and `nimble_b200/aligner` with the same argv;
3. diffs per read (joined on r1_QNAME): the `nimble_features` call, and the columns nimble's `report` consumes
   (`nimble/__main__.py:237-241`); then runs `report` (nb200_report_file, pinned to the reference's pandas code by
   tests/golden) on both per-read files and diffs the count tables.
Prints one JSON summary; exit code 0 = identical, 1 = differences (listed), 3 = aligner not available.
Needs a GPU for step 2's second half (no CPU path)."""
import argparse
import csv
import gzip
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def fixtures(out_dir, n_reads):
    import numpy as np
    from file_bench import write_bam
    from nimble_b200 import synth
    fx = []
    for name, kw, n in (("cfg1", dict(n_founders=1, alleles_per_founder=2, length=600, snps_mean=8.0, seed=11), 400),
                        ("cfg2", dict(n_founders=40, alleles_per_founder=50, length=1098, snps_mean=15.0, seed=1), n_reads)):
        lib, codes = synth.allele_family_library(**kw)
        r1, truth = synth.sample_reads(codes, n, read_len=90, seed=2)
        key = synth.barcodes_10x(n, n_cells=max(4, n // 1000), seed=2, truth=truth)
        lib_path = os.path.join(out_dir, name + ".json")
        with open(lib_path, "w") as f:
            json.dump(lib, f)
        bam = os.path.join(out_dir, name + ".bam")
        write_bam(bam, np.ascontiguousarray(r1), key)
        fx.append((name, lib_path, bam))
    return fx


def read_tsv(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt", newline="") as f:
        rows = list(csv.reader(f, delimiter="\t", quoting=csv.QUOTE_NONE))
    if not rows:
        return [], {}
    hdr = rows[0]
    qi = hdr.index("r1_QNAME") if "r1_QNAME" in hdr else None
    return hdr, {(r[qi] if qi is not None else str(i)): dict(zip(hdr, r)) for i, r in enumerate(rows[1:])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--aligner", required=True, help="the reference's aligner executable (nimble/aligner after `nimble download`)")
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--cores", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out-dir", default="/tmp/nb200_ref_diff")
    a = ap.parse_args()
    if not (os.path.isfile(a.aligner) and os.access(a.aligner, os.X_OK)):
        print(json.dumps({"status": "aligner not available", "path": a.aligner,
                          "note": "X1-X5 stay pinned to oracle/nimble_oracle.c only (DESIGN.md §2)"}))
        return 3
    os.makedirs(a.out_dir, exist_ok=True)
    ours = os.path.join(ROOT, "nimble_b200", "aligner")
    from nimble_b200 import frontend
    summary, differs = {"status": "ran", "fixtures": {}}, False
    for name, lib_path, bam in fixtures(a.out_dir, a.reads):
        outs = {}
        for tag, exe in (("ref", a.aligner), ("b200", ours)):
            out = os.path.join(a.out_dir, "%s.%s.tsv.gz" % (name, tag))
            argv = [exe, "--input", bam, "-c", str(a.cores), "--strand_filter", "unstranded", "-r", lib_path, "-o", out]   # __main__.py:177-192
            rc = subprocess.call(argv)
            if rc != 0:
                print(json.dumps({"status": "aligner failed", "which": tag, "rc": rc, "argv": argv}))
                return 1
            outs[tag] = out
        h_ref, ref = read_tsv(outs["ref"])
        h_our, our = read_tsv(outs["b200"])
        cols = ["nimble_features", "nimble_score", "r1_CB", "r1_UB"]            # what report() consumes (:237-241)
        only_ref = sorted(set(ref) - set(our))
        only_our = sorted(set(our) - set(ref))
        diff = [(q, {c: (ref[q].get(c), our[q].get(c)) for c in cols if ref[q].get(c) != our[q].get(c)}) for q in sorted(set(ref) & set(our))]
        diff = [(q, d) for q, d in diff if d]
        counts = {}
        for tag in ("ref", "b200"):
            cpath = os.path.join(a.out_dir, "%s.%s.counts.tsv" % (name, tag))
            frontend.report(outs[tag], cpath, None, 0.05, False)
            with open(cpath) as f:
                counts[tag] = f.read().splitlines()
        cdiff = sorted(set(counts["ref"]) ^ set(counts["b200"]))
        summary["fixtures"][name] = {
            "reads_called_ref": len(ref), "reads_called_b200": len(our), "called_only_by_ref": len(only_ref), "called_only_by_b200": len(only_our),
            "called_by_both_with_different_columns": len(diff), "count_rows_ref": len(counts["ref"]), "count_rows_b200": len(counts["b200"]),
            "count_rows_differing": len(cdiff), "columns_missing_here": [c for c in h_ref if c not in h_our],
            "examples": {"only_ref": only_ref[:5], "only_b200": only_our[:5], "different": diff[:5], "count_rows": cdiff[:10]}}
        differs |= bool(only_ref or only_our or diff or cdiff)
    summary["identical"] = not differs
    print(json.dumps(summary, indent=1))
    return 1 if differs else 0


if __name__ == "__main__":
    sys.exit(main())
