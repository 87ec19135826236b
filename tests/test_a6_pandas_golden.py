"""A6 against fixtures produced by the reference's own pandas report() (tests/golden/make_a6_golden.py).
CPU: both oracle restatements.  GPU: nb200_umi_counts through the C ABI."""
import json
import os

import pytest

from oracle import a6_py
from oracle import oracle as O
from test_a6_reference_vectors import gpu_report

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "a6_pandas_cases.json")) as f:
    GOLD = json.load(f)
CASES = [c for c in GOLD["cases"] if not c["reference_error"]]


def rows_of(c):
    return [(cb, umi, f, s) for f, s, cb, umi in c["rows"]]


def tsv(out):
    return "".join("%s\t%d\t%s\n" % r for r in out)


def check(c, out, dropped):
    assert tsv(out) == (c["expected_tsv"] or ""), "case %d" % c["id"]
    log = c["reference_stdout"]
    assert ("Dropped %d UMIs" % dropped) in log or "No data" in log, "case %d dropped-count" % c["id"]


def test_fixture_is_complete():
    assert len(CASES) >= 250 and GOLD["reference"].startswith("nimble/__main__.py report()")


def test_python_oracle_matches_pandas_reference():
    for c in CASES:
        check(c, *a6_py.report_counts(rows_of(c), c["threshold"], c["disable_thresholding"]))


def test_c_oracle_matches_pandas_reference():
    for c in CASES:
        check(c, *O.a6_strings(rows_of(c), c["threshold"], c["disable_thresholding"]))


def test_kahan_summation_is_required():
    """The fixtures contain cases a plain left-to-right sum gets wrong (pandas group_sum is Kahan)."""
    saved = a6_py.kahan_sum

    def plain(v):
        s = 0.0
        for x in v:
            s += x
        return s
    a6_py.kahan_sum = plain
    try:
        bad = sum(tsv(a6_py.report_counts(rows_of(c), c["threshold"], c["disable_thresholding"])[0]) != (c["expected_tsv"] or "")
                  for c in CASES)
    finally:
        a6_py.kahan_sum = saved
    assert bad > 0


@pytest.mark.gpu
def test_gpu_matches_pandas_reference(engine):
    for c in CASES:
        rows = [r for r in rows_of(c)]
        out, dropped = gpu_report(engine, rows, c["threshold"], c["disable_thresholding"])
        check(c, out, dropped)
