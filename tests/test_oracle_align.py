"""Oracle aligner sanity: the frozen SPEC on hand-checkable inputs.  No GPU."""
import numpy as np
import pytest

from nimble_b200 import synth
from oracle import oracle as O


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def lib_of(seqs, names=None, **cfg):
    names = names or ["ref%d" % i for i in range(len(seqs))]
    c = dict(synth.DEFAULT_CONFIG)
    c.update(cfg)
    return [c, {"headers": ["reference_genome", "sequence_name", "nt_length", "sequence"],
                "columns": [["t"] * len(seqs), names, [str(len(s)) for s in seqs], seqs]}]


RNG = np.random.default_rng(99)
REF_A = "".join(RNG.choice(list("ACGT"), size=400))
REF_B = "".join(RNG.choice(list("ACGT"), size=400))
REF_A2 = REF_A[:200] + ("A" if REF_A[200] != "A" else "C") + REF_A[201:]     # one SNP vs REF_A


def align1(lib, reads, r2=None, k=20, strand="unstranded"):
    lo = O.Library(lib, k=k, strand_filter=strand)
    res, feats = O.align(lo, reads, r2)
    return lo, res, feats


def test_exact_read_unique_reference():
    lo, res, feats = align1(lib_of([REF_A, REF_B]), [REF_A[50:140]])
    r = res[0]
    assert r["status"][0] == 1 and r["score"][0] == 90 and r["edits"][0] == 0 and r["n_hits"][0] == 71
    assert r["status"][1] == 2 and r["reason"] == 0 and r["n_sw"] == 0
    assert [lo.features[i] for i in feats[0, :r["n_feat"]]] == ["ref0"]


def test_reverse_complement_read_uses_config_R():
    lo, res, feats = align1(lib_of([REF_A, REF_B]), [rc(REF_B[10:100])])
    r = res[0]
    assert r["status"][1] == 1 and r["score"][1] == 90 and r["config"] == 1
    assert [lo.features[i] for i in feats[0, :r["n_feat"]]] == ["ref1"]
    # strand filters: fiveprime only accepts the forward orientation
    _, res5, _ = align1(lib_of([REF_A, REF_B]), [rc(REF_B[10:100])], strand="fiveprime")
    assert res5[0]["reason"] == 1 and res5[0]["n_feat"] == 0
    _, res3, _ = align1(lib_of([REF_A, REF_B]), [rc(REF_B[10:100])], strand="threeprime")
    assert res3[0]["reason"] == 0


def test_ambiguous_read_reports_both_alleles_and_snp_read_only_one():
    lib = lib_of([REF_A, REF_A2, REF_B], names=["A*01", "A*02", "B*01"])
    lo, res, feats = align1(lib, [REF_A[20:110], REF_A[150:240], REF_A2[150:240]])
    assert [lo.features[i] for i in feats[0, :res[0]["n_feat"]]] == ["A*01", "A*02"]
    assert [lo.features[i] for i in feats[1, :res[1]["n_feat"]]] == ["A*01"]
    assert [lo.features[i] for i in feats[2, :res[2]["n_feat"]]] == ["A*02"]


def test_smith_waterman_scores_one_mismatch():
    read = list(REF_A[100:190])
    read[45] = "A" if read[45] != "A" else "C"
    lo, res, feats = align1(lib_of([REF_A, REF_B]), ["".join(read)])
    r = res[0]
    # 90 aligned bases, one mismatch: 89 matches - 2 = 87 ; V = 64*87 - 1 edit
    assert r["n_sw"] == 1 and r["score"][0] == 87 and r["edits"][0] == 1 and r["n_hits"][0] == 71 - 20
    assert r["status"][0] == 1 and r["reason"] == 0


def test_smith_waterman_refines_candidates_beyond_kmer_evidence():
    """A sequencing error sits inside every k-mer that tells A*01 from A*02: k-mer evidence keeps
    both alleles, the banded alignment keeps only the better one."""
    read = list(REF_A2[150:240])            # carries A*02's base at read position 50
    for p in (49, 51):                      # errors on both sides: every 31-mer covering position 50 is gone
        read[p] = "A" if read[p] != "A" else "C"
    read = "".join(read)
    lib = lib_of([REF_A, REF_A2, REF_B], names=["A*01", "A*02", "B*01"])
    lo, res, feats = align1(lib, [read], k=31)
    r = res[0]
    assert r["n_sw"] == 1 and r["n_cand"][0] == 1 and r["edits"][0] == 2 and r["score"][0] == 88 - 4
    assert [lo.features[i] for i in feats[0, :r["n_feat"]]] == ["A*02"]
    # with num_mismatches = 1 the runner-up (one more mismatch) stays in the class
    lib1 = lib_of([REF_A, REF_A2, REF_B], names=["A*01", "A*02", "B*01"], num_mismatches=1)
    lo1, res1, feats1 = align1(lib1, [read], k=31)
    assert res1[0]["n_cand"][0] == 2 and res1[0]["n_feat"] == 2


def test_indel_inside_band():
    read = REF_B[100:140] + REF_B[142:192]          # 2-base deletion in the read
    lo, res, _ = align1(lib_of([REF_A, REF_B]), [read])
    r = res[0]
    # 90 matches, one 2-base gap: 90 - 6 = 84, 2 edits
    assert r["score"][0] == 84 and r["edits"][0] == 2 and r["status"][0] == 1


def test_short_reads_n_bases_and_thresholds():
    lo, res, _ = align1(lib_of([REF_A, REF_B]), [REF_A[:19], "", "N" * 50, REF_A[:24], REF_A[:30], REF_A[:30] + "N" * 40])
    assert res[0]["status"][0] == 2 and res[1]["status"][0] == 2 and res[2]["status"][0] == 2
    assert res[3]["score"][0] == 24 and res[3]["reason"] == 4          # passes score_threshold 20, fails score_filter 25
    assert res[4]["score"][0] == 30 and res[4]["reason"] == 0
    assert res[5]["score"][0] == 30 and res[5]["status"][0] == 5       # 30/70 < score_percent 0.5


def test_group_on_collapses_alleles_to_genes():
    lib, codes = synth.allele_family_library(n_founders=3, alleles_per_founder=5, length=300, seed=8, extra_columns=True,
                                             config={"group_on": "gene"})
    reads, truth = synth.sample_reads(codes, 300, read_len=80, off_target=0.0, err_rate=0.0, n_frac=0.0, rc_frac=0.0, seed=9)
    lo = O.Library(lib)
    res, feats = O.align(lo, ["".join(map(chr, r)) for r in reads])
    genes = lib[1]["columns"][4]
    for i in range(len(res)):
        assert res[i]["reason"] == 0 and res[i]["n_feat"] == 1
        assert lo.features[feats[i, 0]] == genes[truth[i]]


def test_max_hits_and_multi_hits_filters():
    seqs = [REF_A] * 12
    lo, res, _ = align1(lib_of(seqs), [REF_A[10:100]])
    assert res[0]["reason"] == 6 and res[0]["n_cand"][0] == 12        # 12 > max_hits_to_report 10
    lo, res, _ = align1(lib_of(seqs[:4], discard_multi_hits=3), [REF_A[10:100]])
    assert res[0]["reason"] == 5
    lo, res, _ = align1(lib_of(seqs[:4], discard_multiple_matches=True), [REF_A[10:100]])
    assert res[0]["status"][0] == 6 and res[0]["reason"] == 1


@pytest.mark.parametrize("level,expect_reason,expect", [(0, 0, ["A*01", "A*02"]), (1, 0, ["A*01"]), (2, 0, ["A*01"])])
def test_pair_intersect_levels(level, expect_reason, expect):
    lib = lib_of([REF_A, REF_A2, REF_B], names=["A*01", "A*02", "B*01"], intersect_level=level)
    m1 = REF_A[20:110]                     # ambiguous A*01/A*02
    m2 = rc(REF_A[150:240])                # covers the SNP: A*01 only
    lo, res, feats = align1(lib, [m1], [m2])
    assert res[0]["reason"] == expect_reason and res[0]["config"] == 0 and res[0]["pair_score"] == 180
    assert [lo.features[i] for i in feats[0, :res[0]["n_feat"]]] == expect


def test_pair_force_intersect_and_valid_pair():
    lib2 = lib_of([REF_A, REF_B], intersect_level=2)
    lo, res, _ = align1(lib2, [REF_A[20:110]], [rc(REF_B[150:240])])
    assert res[0]["reason"] == 3
    lib1 = lib_of([REF_A, REF_B], intersect_level=1)
    lo, res, feats = align1(lib1, [REF_A[20:110]], [rc(REF_B[150:240])])
    assert res[0]["reason"] == 0 and [lo.features[i] for i in feats[0, :1]] == ["ref0"]      # tie -> mate 1
    libv = lib_of([REF_A, REF_B], require_valid_pair=True)
    lo, res, _ = align1(libv, [REF_A[20:110]], ["ACGT" * 20])
    assert res[0]["reason"] == 2
    lo, res, _ = align1(lib_of([REF_A, REF_B]), [REF_A[20:110]], ["ACGT" * 20])
    assert res[0]["reason"] == 0 and res[0]["pair_score"] == 90


def test_too_long_read_is_an_error():
    lo = O.Library(lib_of([REF_A, REF_B]))
    with pytest.raises(ValueError):
        O.align(lo, ["A" * 501])


def test_threads_do_not_change_results():
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=6, length=400, seed=3)
    reads, _ = synth.sample_reads(codes, 3000, read_len=90, seed=4)
    lo = O.Library(lib)
    off = np.arange(0, reads.size + 1, reads.shape[1], dtype=np.int64)
    a = O.align(lo, (reads.reshape(-1), off), n_threads=1)
    b = O.align(lo, (reads.reshape(-1), off), n_threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_a6_threads_equal_serial():
    """orc_a6_mt (cells dealt to host threads: the benchmark's CPU arm) == orc_a6, table and dropped count."""
    lib, codes = synth.allele_family_library(n_founders=5, alleles_per_founder=10, length=500, snps_mean=8, seed=41)
    lo = O.Library(lib)
    r1, truth = synth.sample_reads(codes, 60000, read_len=80, seed=42)
    key = synth.barcodes_10x(len(r1), n_cells=300, seed=42, truth=truth)
    key[::97] = np.uint64(0xFFFFFFFFFFFFFFFF)
    off = np.arange(0, r1.size + 1, r1.shape[1], dtype=np.int64)
    ro, fo = O.align(lo, (r1.reshape(-1), off))
    nf = ro["n_feat"].astype(np.int64)
    foff = np.zeros(len(ro) + 1, np.int32)
    np.cumsum(nf, out=foff[1:])
    ids = fo[np.arange(fo.shape[1])[None, :] < nf[:, None]].astype(np.uint32)
    score = np.random.default_rng(3).random(len(ro)) * 3
    for sc in (None, score):
        a = O.a6_ids(key, foff, ids, sc, lo.tok_end, lo.tok_comma, 0.05, False)
        for t in (0, 3, 7):
            b = O.a6_ids(key, foff, ids, sc, lo.tok_end, lo.tok_comma, 0.05, False, n_threads=t)
            assert all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4])) and a[4] == b[4]
    assert len(a[0]) > 300
