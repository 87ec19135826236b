"""ORACLE (test infrastructure, never shipped, never on the product path).

Pure-Python restatement of nimble's post-alignment UMI stage ("A6" in SURVEY.md §8a):
per-read TSV rows -> per-cell counts.  Small inputs only; the C oracle
(`oracle/nimble_oracle.c: orc_a6`) is the fast id-based restatement and is checked
against this file.

Pinned against: the 24 known-answer tests of /root/reference/test/test.py
(tests/golden/a6_reference_tests.json) and fixtures produced by running the reference's
own pandas code (tests/golden/make_a6_golden.py -> tests/golden/a6_pandas_*.json).

Follows (reference file:line):
  convert_df_to_proper_umi   nimble/__main__.py:234-252
  per_umi_thresholding       nimble/utils.py:119-207
  umi_intersection           nimble/utils.py:209-219
  intersect_lists            nimble/utils.py:221-224
  report (count + ordering)  nimble/__main__.py:254-293
"""
from __future__ import annotations

import math


MAX_ROUNDS = 64


def kahan_sum(values):
    """pandas' groupby-sum kernel (group_sum) is Kahan-compensated, in order of appearance.
    Verified numerically against pandas 3.0.2 in tests/golden/make_a6_golden.py."""
    s = 0.0
    c = 0.0
    for v in values:
        y = v - c
        t = s + y
        c = t - s - y
        if c != c:  # inf handling as in pandas
            c = 0.0
        s = t
    return s


def _is_null(x):
    return x is None or (isinstance(x, float) and math.isnan(x))


def merge_rows(rows):
    """rows: iterable of (cb, umi, features_string, score).
    nimble/__main__.py:244-251 — drop null/empty, sort names inside the feature string,
    groupby(cb, umi, features).sum(); pandas groupby sorts its keys ascending (str order)."""
    acc = {}
    for cb, umi, feats, score in rows:
        if _is_null(cb) or _is_null(umi) or _is_null(feats) or _is_null(score):
            continue
        if cb == "" or umi == "" or feats == "":
            continue
        feats = ",".join(sorted(feats.split(",")))
        acc.setdefault((cb, umi, feats), []).append(float(score))
    out = []
    for key in sorted(acc):
        out.append((key[0], key[1], key[2], kahan_sum(acc[key])))
    return out


def threshold_group(group, threshold):
    """group: list of (features_string, score) for one (cb, umi), in merged (sorted) order.
    Returns the kept feature-name set.  nimble/utils.py:120-171."""
    def scores(drop):
        per = {}
        order = []
        total = 0.0
        any_row = False
        for feats, s in group:
            names = [f for f in feats.split(",") if f not in drop]
            if not names:
                continue
            any_row = True
            share = s / len(names)
            total += s
            for f in names:
                if f not in per:
                    per[f] = []
                    order.append(f)
                per[f].append(share)
        return {f: kahan_sum(per[f]) for f in order}, total, any_row

    fs, total, _ = scores(set())
    # The reference loops `while True`; ratios are monotone non-decreasing after a drop, so
    # in exact arithmetic it ends after <= 2 rounds.  MAX_ROUNDS only bounds float-rounding
    # pathologies where the reference itself would not terminate (documented in DESIGN.md).
    for _round in range(MAX_ROUNDS):
        if not fs:
            return set()
        to_drop = set()
        for f, v in fs.items():
            # feature_ratios = feature_scores / total_score ; ratios < threshold (strict)
            if total == 0.0:
                ratio = float("nan") if v == 0.0 else math.copysign(float("inf"), v)
            else:
                ratio = v / total
            if ratio < threshold:
                to_drop.add(f)
        if not to_drop:
            return set(fs)
        # NB the reference re-derives from the ORIGINAL rows removing only this round's
        # to_drop (utils.py:158-159) — names dropped in an earlier round come back into the
        # recomputation.  Restated faithfully.
        fs, total, any_row = scores(to_drop)
        if not any_row:
            return set()
    return set(fs)


def report_counts(rows, threshold=0.05, disable_thresholding=False):
    """Full A6: returns (list of (feature_string, count, cell_barcode) in output order,
    number of UMIs dropped for empty intersection).  nimble/__main__.py:254-293."""
    merged = merge_rows(rows)
    groups = {}
    for cb, umi, feats, s in merged:
        groups.setdefault((cb, umi), []).append((feats, s))
    per_umi = []
    dropped_empty = 0
    for (cb, umi) in sorted(groups):
        g = groups[(cb, umi)]
        lists = []
        if disable_thresholding:
            for feats, _ in g:
                lists.append(feats.split(","))
        else:
            keep = threshold_group(g, threshold)
            for feats, _ in g:
                k = sorted(set(feats.split(",")) & keep)
                if k:  # utils.py:205 rows with empty filtered_features removed
                    lists.append(k)
        if not lists:
            continue  # the (cb,umi) vanished entirely before umi_intersection
        inter = sorted(set.intersection(*map(set, lists)))
        if not inter:
            dropped_empty += 1
            continue
        per_umi.append((cb, ",".join(inter)))
    counts = {}
    for cb, f in per_umi:
        counts[(cb, f)] = counts.get((cb, f), 0) + 1
    out = [(f, counts[(cb, f)], cb) for (cb, f) in sorted(counts)]
    return out, dropped_empty
