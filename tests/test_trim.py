"""`--trim <TARGET_LENGTH>:<STRICTNESS>` (nimble/__main__.py:191-192,400): the MaxInfo criterion of csrc/trim.hpp against its
restatement oracle/trim_py.py (CPU), and the file pipeline with --trim against the same pipeline on pre-trimmed reads (GPU).
Parity of this piece is UNPINNED (the real aligner's arithmetic is not available): DESIGN.md §2.9."""
import ctypes as ct
import gzip
import os

import numpy as np
import pytest

from nimble_b200 import _lib, frontend, synth
from oracle import trim_py


def test_maxinfo_matches_restatement():
    L = _lib.load()
    rng = np.random.default_rng(9)
    cases = [np.full(90, 30), np.full(90, 2), np.r_[np.full(60, 35), np.full(30, 2)], np.full(150, 20), np.full(1, 40), np.full(500, 37)]
    for _ in range(400):
        n = int(rng.integers(1, 301))
        q = rng.integers(0, 42, n)
        if rng.random() < 0.5:                          # a decaying tail, like real reads
            q = np.clip(38 - (np.arange(n) * rng.random() * 0.6).astype(int) + rng.integers(-4, 5, n), 0, 41)
        cases.append(q)
    for q in cases:
        for target, strict in ((50, 0.9), (36, 0.5), (100, 0.2), (0, 1.0), (50, 0.0)):
            want = trim_py.trim_maxinfo([int(x) for x in q], target, strict)
            raw = np.ascontiguousarray(q, np.uint8)
            assert L.nb200_trim_maxinfo(raw.ctypes.data, len(raw), 0, target, strict) == want
            asc = (raw + 33).astype(np.uint8)
            assert L.nb200_trim_maxinfo(asc.ctypes.data, len(asc), 33, target, strict) == want
    # high-quality 90-base reads are left alone with the reference's defaults (types.py:24-25); a bad tail goes
    assert trim_py.trim_maxinfo([37] * 90, 50, 0.9) == 90
    assert trim_py.trim_maxinfo([37] * 60 + [2] * 30, 50, 0.9) == 60


def test_trim_argument_forms():
    assert frontend.parse_trim("", 2) == []
    assert frontend.parse_trim("50:0.9", 1) == [(50, 0.9)]
    assert frontend.parse_trim("50:0.9,36:0.5", 2) == [(50, 0.9), (36, 0.5)]
    for bad, n in (("50", 1), ("50:1.5", 1), ("x:0.5", 1), ("50:0.9", 2), ("50:0.9,", 2), (":0.9", 1)):
        with pytest.raises(ValueError):
            frontend.parse_trim(bad, n)


def _fastq(path, names, seqs, quals):
    with gzip.open(path, "wt") as f:
        for n, s, q in zip(names, seqs, quals):
            f.write("@%s\n%s\n+\n%s\n" % (n, s, "".join(chr(33 + int(x)) for x in q)))


@pytest.mark.gpu
def test_file_pipeline_trim_equals_pretrimmed_reads(engine, tmp_path):
    """FASTQ (bulk shape) and BAM (per-read TSV, forward and reverse-strand records) with decaying qualities: `--trim 50:0.9`
    gives exactly the output of the same pipeline, untrimmed, on reads cut to trim_py's lengths; two libraries with
    different settings in one pass."""
    import json
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_frontend import write_bam
    rng = np.random.default_rng(21)
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=6, length=500, snps_mean=6, seed=210)
    r1, _ = synth.sample_reads(codes, 3000, read_len=120, err_rate=0.01, seed=211)
    seqs = [bytes(r).decode() for r in r1]
    quals = []
    for i in range(len(seqs)):
        cut = int(rng.integers(30, 121))
        q = np.r_[rng.integers(30, 40, cut), rng.integers(2, 12, 120 - cut)]
        quals.append(q)
    libs = []
    for j in range(2):
        p = str(tmp_path / ("lib%d.json" % j))
        with open(p, "w") as f:
            json.dump(lib, f)
        libs.append(p)
    settings = [(50, 0.9), (80, 0.4)]
    trim_arg = ",".join("%d:%g" % s for s in settings)
    names = ["q%05d" % i for i in range(len(seqs))]
    # ---- FASTQ ----
    fq = str(tmp_path / "in.fastq.gz")
    _fastq(fq, names, seqs, quals)
    out = str(tmp_path / "o.tsv")
    assert frontend.align(",".join(libs), out, [fq], 4, "unstranded", trim_arg, None, engine=engine) == 0
    for j, (t, s) in enumerate(settings):
        keep = [trim_py.trim_maxinfo([int(x) for x in q], t, s) for q in quals]
        assert min(keep) < 120 and max(keep) > 30
        fq2 = str(tmp_path / ("pre%d.fastq.gz" % j))
        _fastq(fq2, names, [sq[:k] for sq, k in zip(seqs, keep)], [q[:k] for q, k in zip(quals, keep)])
        ref_out = str(tmp_path / ("ref%d.tsv" % j))
        assert frontend.align(libs[j], ref_out, [fq2], 4, "unstranded", "", None, engine=engine) == 0
        got = open(str(tmp_path / ("o.lib%d.tsv" % j))).read()
        assert got == open(ref_out).read() and got.count("\n") > 3
    # ---- paired FASTQ: both mates trimmed, each by its own qualities ----
    a1, a2, _ = synth.sample_pairs(codes, 1500, read_len=100, insert_mean=220, insert_sd=25, err_rate=0.01, off_target=0.1, seed=212)
    s1, s2 = [bytes(r).decode() for r in a1], [bytes(r).decode() for r in a2]
    q1 = [np.r_[rng.integers(30, 40, c), rng.integers(2, 12, 100 - c)] for c in rng.integers(25, 101, len(s1))]
    q2 = [np.r_[rng.integers(30, 40, c), rng.integers(2, 12, 100 - c)] for c in rng.integers(25, 101, len(s2))]
    pn = ["p%05d" % i for i in range(len(s1))]
    f1, f2 = str(tmp_path / "p_R1.fastq.gz"), str(tmp_path / "p_R2.fastq.gz")
    _fastq(f1, pn, s1, q1); _fastq(f2, pn, s2, q2)
    outp = str(tmp_path / "p.tsv")
    assert frontend.align(libs[0], outp, [f1, f2], 4, "unstranded", "60:0.8", None, engine=engine) == 0
    k1 = [trim_py.trim_maxinfo([int(x) for x in q], 60, 0.8) for q in q1]
    k2 = [trim_py.trim_maxinfo([int(x) for x in q], 60, 0.8) for q in q2]
    g1, g2 = str(tmp_path / "pp_R1.fastq.gz"), str(tmp_path / "pp_R2.fastq.gz")
    _fastq(g1, pn, [x[:k] for x, k in zip(s1, k1)], [q[:k] for q, k in zip(q1, k1)])
    _fastq(g2, pn, [x[:k] for x, k in zip(s2, k2)], [q[:k] for q, k in zip(q2, k2)])
    refp = str(tmp_path / "pref.tsv")
    assert frontend.align(libs[0], refp, [g1, g2], 4, "unstranded", "", None, engine=engine) == 0
    assert open(outp).read() == open(refp).read() and open(outp).read().count("\n") > 3
    # ---- BAM: half of the records stored reverse-complemented (flag 16) with reversed qualities ----
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    recs, recs_pre = [], [[], []]
    keeps = [[trim_py.trim_maxinfo([int(x) for x in q], t, s) for q in quals] for (t, s) in settings]
    for i, (sq, q) in enumerate(zip(seqs, quals)):
        tags = {"CB": "ACGTACGTACGTAC%s" % "ACGT"[i % 4] + "A", "UB": "".join("ACGT"[(i >> (2 * b)) & 3] for b in range(10))}
        rev = i % 2 == 1
        stored = "".join(comp[c] for c in reversed(sq)) if rev else sq
        recs.append((names[i], 16 if rev else 0, stored, tags, list(reversed(q)) if rev else list(q)))
        for j in range(2):
            k = keeps[j][i]
            cut = sq[:k]
            recs_pre[j].append((names[i], 16 if rev else 0, "".join(comp[c] for c in reversed(cut)) if rev else cut, tags))
    bam = str(tmp_path / "in.bam")
    write_bam(bam, recs)
    outb = str(tmp_path / "b.tsv")
    assert frontend.align(",".join(libs), outb, [bam], 4, "unstranded", trim_arg, None, engine=engine) == 0
    for j in range(2):
        bam2 = str(tmp_path / ("pre%d.bam" % j))
        write_bam(bam2, recs_pre[j])
        ref_out = str(tmp_path / ("refb%d.tsv" % j))
        assert frontend.align(libs[j], ref_out, [bam2], 4, "unstranded", "", None, engine=engine) == 0
        got = open(str(tmp_path / ("b.lib%d.tsv" % j))).read()
        assert got == open(ref_out).read() and got.count("\n") > 1000
