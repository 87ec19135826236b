"""A5 on the GPU (nb200_correct_barcodes / nb200_fastq_to_bam through the C ABI) against the
reference-generated fixtures and the C oracle.  Bit-exact: whitelist index and status per read,
every counter the reference prints."""
import gzip
import json
import os

import numpy as np
import pytest

from nimble_b200 import frontend, synth
from oracle import a5_py as A
from oracle import oracle as O
from test_a5_golden import CASES, case_arrays, quals

pytestmark = pytest.mark.gpu

KEYS = ("cb_perfect_match", "cb_corrected", "cb_no_correction", "cache_size")


def same_len(whitelist, L):
    return [w for w in whitelist if len(w) == L]


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_correct_barcodes_matches_reference_fixture(engine, ci):
    c = CASES[ci]
    L = c["cb_length"]
    wl_s = same_len(c["whitelist"], L)        # entries of another length are inert (file loader skips them too)
    wl = engine.load_whitelist(wl_s, L)
    cb, q, el = case_arrays(c)
    idx, status, st = engine.correct_barcodes(wl, cb, q, el)
    got = [[k, wl.entry(idx[k])] for k in range(len(el)) if status[k] in (A.CB_PERFECT, A.CB_CORRECTED)]
    assert got == c["records"]
    for k in KEYS:
        assert st[k] == c["stats"][k], k
    assert ((status == A.CB_SKIPPED) == (el == 0)).all()


def write_fastq(path, recs, gz=False):
    op = gzip.open if gz else open
    with op(path, "wt") as f:
        for rid, seq, qual in recs:
            f.write("@%s some description\n%s\n+\n%s\n" % (rid, seq, qual))


@pytest.mark.parametrize("ci", range(0, len(CASES), 3))
def test_fastq_to_bam_matches_reference_fixture(engine, tmp_path, ci):
    c = CASES[ci]
    L, U = c["cb_length"], c["umi_length"]
    gz = ci % 2 == 0
    r1, r2 = str(tmp_path / ("r1.fastq" + (".gz" if gz else ""))), str(tmp_path / "r2.fastq")
    write_fastq(r1, [(p[0], p[1], p[2]) for p in c["pairs"]], gz)
    write_fastq(r2, [(p[3], p[4], p[5]) for p in c["pairs"]])
    wlp = str(tmp_path / ("wl.txt" + (".gz" if not gz else "")))
    with (gzip.open if not gz else open)(wlp, "wt") as f:
        f.write("\n".join(c["whitelist"][:3]) + "\n\n  " + "\n".join(c["whitelist"][3:]) + "  \n")
    out = str(tmp_path / "out.bam")
    st = engine.fastq_to_bam(r1, r2, wlp, out, L, U)
    for k, v in c["stats"].items():
        assert st[k] == v, k
    recs = frontend.read_bam(out, with_qual=True)
    assert len(recs) == 2 * len(c["records"])
    for j, (pi, cb) in enumerate(c["records"]):
        id1, s1, q1, id2, s2, q2 = c["pairs"][pi]
        a, b = recs[2 * j], recs[2 * j + 1]
        name = A.removesuffix(id1, "/1")
        assert a[0] == name and b[0] == name and a[1] == 77 and b[1] == 141 and a[4] == -1 and b[4] == -1
        assert a[2] == s1[L + U:] and a[5] == q1[L + U:] and b[2] == s2 and b[5] == q2
        assert a[3] == {"CB": cb, "UB": s1[L:L + U]} and b[3] == a[3]


@pytest.mark.parametrize("L,clustered,n", [(16, 0.0, 400_000), (16, 0.5, 300_000), (9, 0.6, 100_000), (21, 0.3, 100_000)])
def test_correct_barcodes_matches_oracle(engine, L, clustered, n):
    wl_a, cb, q = synth.barcode_workload(n, n_whitelist=50_000, n_cells=2000, cb_length=L, err_rate=0.02, n_rate=0.002,
                                         off_whitelist=0.02, seed=3 * L, clustered=clustered)
    q = (q // 8 * 8).astype(np.uint8)
    rng = np.random.default_rng(L)
    el = (rng.random(n) < 0.97).astype(np.uint8)
    cb[rng.integers(0, n, 50), rng.integers(0, L, 50)] = ord("x")        # bytes outside ACGTN
    wl_s = [bytes(r).decode() for r in wl_a]
    wl = engine.load_whitelist(wl_s, L)
    idx, status, st = engine.correct_barcodes(wl, cb, q, el)
    o_idx, o_status, o_st = O.cb_correct(wl_s, cb, q, el, L)
    assert np.array_equal(status, o_status)
    assert np.array_equal(idx, o_idx)
    for k in KEYS:
        assert st[k] == o_st[k], k
    assert st["n_multi"] > 0 if clustered else True
    # resident arm gives the same answer
    engine.cb_upload(wl, cb, q, el)
    idx2, status2, st2 = engine.correct_barcodes_resident(wl)
    assert np.array_equal(idx2, idx) and np.array_equal(status2, status) and st2["cache_size"] == st["cache_size"]


def test_first_read_decides_for_its_raw_barcode(engine):
    """The reference's correction_cache: qualities of the first read with a raw barcode choose between candidates."""
    wl_s = ["AAAAAAAAAAAAAAAC", "AAAAAAAAAAAAAACA", "TTTTTTTTTTTTTTTT"]
    wl = engine.load_whitelist(wl_s, 16)
    raw = "AAAAAAAAAAAAAAAA"                        # one substitution from entries 0 and 1
    cb = np.frombuffer((raw * 4).encode(), np.uint8).reshape(4, 16)
    q = np.full((4, 16), 30, np.uint8)
    q[0, 14] = 5                                    # first read: position 14 is the doubtful base -> entry 1
    q[1, 15] = 2; q[2, 15] = 2; q[3, 15] = 2        # later reads would choose entry 0 on their own
    idx, status, st = engine.correct_barcodes(wl, cb, q)
    assert idx.tolist() == [1, 1, 1, 1] and (status == A.CB_CORRECTED).all() and st["cache_size"] == 1 and st["n_multi"] == 4
    el = np.array([0, 1, 1, 1], np.uint8)           # an ineligible pair never reaches the cache
    idx, status, st = engine.correct_barcodes(wl, cb, q, el)
    assert idx.tolist() == [-1, 0, 0, 0] and status.tolist() == [0, 2, 2, 2]
    q[:] = 30                                       # equal qualities: ascending string order -> ...AAAC < ...AACA
    idx, _, _ = engine.correct_barcodes(wl, cb, q)
    assert idx.tolist() == [0, 0, 0, 0]


def test_edge_cases(engine):
    wl = engine.load_whitelist(["ACGTACGTACGTACGT", "ACGTACGTACGTACGT", "NCGTACGTACGTACGT"], 16)
    assert wl.n_entries == 3 and wl.n_unique == 2
    idx, status, st = engine.correct_barcodes(wl, np.zeros((0, 16), np.uint8), np.zeros((0, 16), np.uint8))
    assert len(idx) == 0 and st["cache_size"] == 0
    rows = ["ACGTACGTACGTACGT", "NCGTACGTACGTACGT", "GCGTACGTACGTACGT", "acgtACGTACGTACGT", "ACGTACGTACGTACGN", "TTTTTTTTTTTTTTTT"]
    cb = np.frombuffer("".join(rows).encode(), np.uint8).reshape(-1, 16)
    q = np.full(cb.shape, 20, np.uint8)
    idx, status, st = engine.correct_barcodes(wl, cb, q)
    #   exact (first of the duplicate lines), exact N entry, two candidates -> 'A..' < 'N..', lower case = no match,
    #   N in the read corrected, no candidate
    assert idx.tolist() == [0, 2, 0, -1, 0, -1]
    assert status.tolist() == [1, 1, 2, 3, 2, 3]
    assert st["cache_size"] == 6
    with pytest.raises(Exception):
        engine.load_whitelist(["ACGTACGTACGTACGX"], 16)
    with pytest.raises(Exception):
        engine.load_whitelist(["ACGT"], 22)


def test_full_size_properties(engine):
    """10 M reads against a 737 k whitelist (the 10x v2 size): properties that need no oracle."""
    n = 10_000_000
    wl_a, cb, q = synth.barcode_workload(n, seed=11)
    wl = engine.load_whitelist([bytes(r).decode() for r in wl_a], 16)
    idx, status, st = engine.correct_barcodes(wl, cb, q)
    assert st["cb_perfect_match"] + st["cb_corrected"] + st["cb_no_correction"] == n
    ok = status != A.CB_NONE
    assert 0.97 < ok.mean() < 0.999 and (status == A.CB_CORRECTED).mean() > 0.02
    # a corrected barcode is one substitution away; a perfect one is identical
    d = (wl_a[idx[ok]] != cb[ok]).sum(axis=1)
    assert np.array_equal(d == 0, status[ok] == A.CB_PERFECT) and d.max() == 1
    # idempotence: corrected barcodes are on the whitelist
    idx2, status2, _ = engine.correct_barcodes(wl, wl_a[idx[ok]], q[ok])
    assert (status2 == A.CB_PERFECT).all() and np.array_equal(idx2, idx[ok])
    # a bounded slice against the oracle, in file order from the start (cache semantics included)
    m = 1_000_000
    o_idx, o_status, _ = O.cb_correct([bytes(r).decode() for r in wl_a], cb[:m], q[:m], None, 16)
    i3, s3, _ = engine.correct_barcodes(wl, cb[:m], q[:m])
    assert np.array_equal(i3, o_idx) and np.array_equal(s3, o_status)


def test_fastq_to_bam_file_edge_cases(engine, tmp_path):
    wlp = str(tmp_path / "wl.txt")
    with open(wlp, "w") as f:
        f.write("ACGTACGTACGTACGT\nTTTTTTTTTTTTTTTT\n")
    out = str(tmp_path / "o.bam")
    # empty inputs -> header-only BAM, all counters zero
    r1, r2 = str(tmp_path / "e1.fastq"), str(tmp_path / "e2.fastq")
    open(r1, "w").close(); open(r2, "w").close()
    st = engine.fastq_to_bam(r1, r2, wlp, out)
    assert st["total_pairs"] == 0 and st["written_pairs"] == 0 and frontend.read_bam(out) == []
    # CRLF line ends, no final newline, R2 longer than R1 (zip stops at the shorter file), lower-case bases kept
    cb, umi = "ACGTACGTACGTACGT", "GGGGCCCCAAAA"
    with open(r1, "wb") as f:
        f.write(("@a/1 x\r\n%s%sacgtn\r\n+\r\n%s\r\n@b\r\n%s%sTT\r\n+\r\n%s" % (cb, umi, "I" * 33, cb[:-1] + "A", umi, "5" * 30)).encode())
    with open(r2, "w") as f:
        f.write("@a/2\nCCCC\n+\nIIII\n@b\nGG\n+\n##\n@c\nAA\n+\nII\n")
    st = engine.fastq_to_bam(r1, r2, wlp, out)
    assert st["total_pairs"] == 2 and st["written_pairs"] == 2 and st["cb_perfect_match"] == 1 and st["cb_corrected"] == 1
    recs = frontend.read_bam(out, with_qual=True)
    assert [r[0] for r in recs] == ["a", "a", "b", "b"] and [r[1] for r in recs] == [77, 141, 77, 141]
    assert recs[0][2] == "ACGTN" and recs[0][5] == "IIIII" and recs[1][2] == "CCCC"      # BAM stores upper-case codes
    assert recs[2][2] == "TT" and recs[2][3] == {"CB": cb, "UB": umi} and recs[3][2] == "GG" and recs[3][5] == "##"
    # errors are reported, not swallowed: missing file, sequence/quality length mismatch, over-long read name
    with pytest.raises(Exception):
        engine.fastq_to_bam(str(tmp_path / "nope.fastq"), r2, wlp, out)
    with open(r1, "w") as f:
        f.write("@a\n%s%sAC\n+\nIII\n" % (cb, umi))
    with pytest.raises(Exception):
        engine.fastq_to_bam(r1, r2, wlp, out)
    long_name = "n" * 300
    with open(r1, "w") as f:
        f.write("@%s\n%s%sAC\n+\n%s\n" % (long_name, cb, umi, "I" * 30))
    with open(r2, "w") as f:
        f.write("@%s\nCC\n+\nII\n" % long_name)
    with pytest.raises(Exception):
        engine.fastq_to_bam(r1, r2, wlp, out)
    # the previous output is still intact (OUT.tmp + rename)
    assert len(frontend.read_bam(out)) == 4


def test_align_10x_fastq_equals_fastq_to_bam_then_align(engine, tmp_path):
    """The one-pass route (no intermediate BAM) writes byte for byte the per-read TSV that fastq-to-bam followed
    by align on its BAM writes."""
    rng = np.random.default_rng(77)
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=6, length=500, snps_mean=6, seed=78)
    libp = str(tmp_path / "lib.json")
    with open(libp, "w") as f:
        json.dump(lib, f)
    n = 3000
    r2, truth = synth.sample_reads(codes, n, read_len=90, seed=79)
    tail, _ = synth.sample_reads(codes, n, read_len=40, seed=80)
    wl_a, cb, q = synth.barcode_workload(n, n_whitelist=500, n_cells=12, err_rate=0.02, n_rate=0.003, off_whitelist=0.05, seed=81,
                                         clustered=0.4)
    umi = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(n, 12))]
    wlp = str(tmp_path / "wl.txt")
    with open(wlp, "w") as f:
        f.write("\n".join(bytes(r).decode() for r in wl_a) + "\n")
    f1, f2 = str(tmp_path / "R1.fastq.gz"), str(tmp_path / "R2.fastq")
    with gzip.open(f1, "wt") as g1, open(f2, "w") as g2:
        for i in range(n):
            t = bytes(tail[i]).decode()[:int(rng.integers(0, 41))]          # some pairs have no remainder: dropped
            s1 = bytes(cb[i]).decode() + bytes(umi[i]).decode() + t
            g1.write("@p%d/1\n%s\n+\n%s\n" % (i, s1, "".join(chr(33 + int(x)) for x in q[i]) + "F" * (len(s1) - 16)))
            s2 = bytes(r2[i]).decode()
            g2.write("@p%d/2\n%s\n+\n%s\n" % (i, s2, "F" * len(s2)))
    lg = engine.load_library(libp)
    bam = str(tmp_path / "x.bam")
    st_a = engine.fastq_to_bam(f1, f2, wlp, bam)
    engine.align_files([bam], [lg], [str(tmp_path / "two_step.tsv")])
    st_b = engine.align_10x_fastq(f1, f2, wlp, [lg], [str(tmp_path / "one_pass.tsv")])
    a, b = open(tmp_path / "two_step.tsv").read(), open(tmp_path / "one_pass.tsv").read()
    assert a == b and a.count("\n") > 500
    for k in ("total_pairs", "written_pairs", "cb_perfect_match", "cb_corrected", "cb_no_correction", "no_remaining_seq", "cache_size"):
        assert st_a[k] == st_b[k], k
    assert st_a["no_remaining_seq"] > 0 and st_a["cb_corrected"] > 0
    # the front-end wrapper and `report` on its output
    assert frontend.align_10x(libp, str(tmp_path / "fe.tsv"), f1, f2, wlp, engine=engine) == 0
    assert open(tmp_path / "fe.tsv").read() == a
    frontend.report(str(tmp_path / "fe.tsv"), str(tmp_path / "counts.tsv"), engine=engine)
    assert sum(1 for _ in open(tmp_path / "counts.tsv")) > 10


@pytest.mark.parametrize("seed", range(3000, 3000 + int(os.environ.get("NB200_FUZZ_CASES", "40"))))
def test_randomised_barcode_differential(engine, seed):
    """Seeded fuzz: barcode length 1..21, tiny dense whitelists (many candidates per read), few distinct
    qualities (ties), N and junk bytes, eligibility masks, repeated raw barcodes (the cache rule)."""
    rng = np.random.default_rng(seed)
    L = int(rng.integers(1, 22))
    alpha = np.frombuffer(b"ACGT" if rng.random() < 0.7 else b"ACGTN", np.uint8)
    n_wl = int(min(rng.integers(1, 400), len(alpha) ** min(L, 6)))
    wl = np.unique(alpha[rng.integers(0, len(alpha), size=(n_wl, L))], axis=0)
    wl = wl[rng.permutation(len(wl))]
    n = int(rng.integers(1, 3000))
    base = wl[rng.integers(0, len(wl), size=n)].copy()
    pool = np.frombuffer(b"ACGTN", np.uint8)
    for _ in range(int(rng.integers(0, 3))):                         # 0-2 substitutions per read
        hit = rng.random(n) < 0.5
        pos = rng.integers(0, L, size=n)
        base[hit, pos[hit]] = pool[rng.integers(0, 5, size=int(hit.sum()))]
    junk = rng.random(n) < 0.02
    base[junk, rng.integers(0, L, size=int(junk.sum()))] = np.frombuffer(b"axZ.", np.uint8)[rng.integers(0, 4, size=int(junk.sum()))]
    rep = rng.random(n) < 0.3                                        # repeat an earlier raw barcode with other qualities
    src = (rng.random(n) * np.arange(n)).astype(np.int64)
    base[rep] = base[src[rep]]
    q = rng.choice(np.array([2, 11, 25, 37], np.uint8) if rng.random() < 0.6 else np.arange(2, 41, dtype=np.uint8), size=(n, L))
    el = None if rng.random() < 0.3 else (rng.random(n) < 0.9).astype(np.uint8)
    wl_s = [bytes(r).decode() for r in wl]
    w = engine.load_whitelist(wl_s, L)
    idx, status, st = engine.correct_barcodes(w, base, q, el)
    o_idx, o_status, o_st = O.cb_correct(wl_s, base, q, el, L)
    assert np.array_equal(status, o_status) and np.array_equal(idx, o_idx)
    for k in KEYS:
        assert st[k] == o_st[k], k
