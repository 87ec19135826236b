"""Seeded synthetic libraries and reads of the shapes BASELINE.json names.

There is no network and the reference ships no fixtures, so every benchmark / parity input is
generated here.  Pure numpy; deterministic for a given seed.
"""
from __future__ import annotations

import numpy as np

_ASCII = np.frombuffer(b"ACGT", np.uint8)

DEFAULT_CONFIG = {  # nimble/types.py:12-25 (json spelling as dumped by nimble/__main__.py:64-65)
    "score_threshold": 20, "score_filter": 25, "score_percent": 0.5, "num_mismatches": 0,
    "discard_multiple_matches": False, "intersect_level": 0, "group_on": "", "discard_multi_hits": 0,
    "require_valid_pair": False, "data_type": "RNA", "filters": [], "max_hits_to_report": 10,
    "trim_target_length": 50, "trim_strictness": 0.9,
}


def allele_family_library(n_founders=40, alleles_per_founder=50, length=1098, snps_mean=15.0,
                          seed=1, name_prefix="Mamu", config=None, extra_columns=False,
                          reference_genome="synthetic mhc"):
    """MHC-like library: `n_founders` random genes, each with `alleles_per_founder` alleles that
    are the founder plus Poisson(snps_mean) substitutions.  Returns (library_json_obj, codes)
    where codes is a list of uint8 arrays (0..3) per allele."""
    rng = np.random.default_rng(seed)
    names, seqs, codes, genes = [], [], [], []
    for g in range(n_founders):
        founder = rng.integers(0, 4, size=length, dtype=np.uint8)
        for a in range(alleles_per_founder):
            s = founder.copy()
            if a > 0:
                n_snp = int(rng.poisson(snps_mean))
                if n_snp:
                    pos = rng.choice(length, size=min(n_snp, length), replace=False)
                    s[pos] = (s[pos] + rng.integers(1, 4, size=len(pos), dtype=np.uint8)) & 3
            codes.append(s)
            names.append("%s-%c%d*%03d:%02d" % (name_prefix, "ABEI"[g % 4], g // 4 + 1, a + 1, 1))
            genes.append("%s-%c%d" % (name_prefix, "ABEI"[g % 4], g // 4 + 1))
            seqs.append(_ASCII[s].tobytes().decode("ascii"))
    cfg = dict(DEFAULT_CONFIG)
    if config:
        cfg.update(config)
    headers = ["reference_genome", "sequence_name", "nt_length", "sequence"]
    columns = [[reference_genome] * len(names), names, [str(len(s)) for s in seqs], seqs]
    if extra_columns:
        headers.append("gene")
        columns.append(genes)
    return [cfg, {"headers": headers, "columns": columns}], codes


def random_transcript_library(n_seqs=1000, mean_len=2000, family_frac=0.3, seed=5, config=None):
    """Transcriptome-like library (BASELINE config 5 shape, scaled by n_seqs): log-normal lengths,
    `family_frac` of the sequences share exon blocks with a sibling."""
    rng = np.random.default_rng(seed)
    names, seqs, codes = [], [], []
    for i in range(n_seqs):
        L = int(np.clip(rng.lognormal(np.log(mean_len), 0.5), 200, 20000))
        if i > 0 and rng.random() < family_frac:
            sib = codes[int(rng.integers(0, i))]
            s = rng.integers(0, 4, size=L, dtype=np.uint8)
            blk = min(len(sib), L) // 2
            a = int(rng.integers(0, len(sib) - blk + 1))
            b = int(rng.integers(0, L - blk + 1))
            s[b:b + blk] = sib[a:a + blk]
        else:
            s = rng.integers(0, 4, size=L, dtype=np.uint8)
        codes.append(s)
        names.append("TX%06d" % i)
        seqs.append(_ASCII[s].tobytes().decode("ascii"))
    cfg = dict(DEFAULT_CONFIG)
    if config:
        cfg.update(config)
    data = {"headers": ["reference_genome", "sequence_name", "nt_length", "sequence"],
            "columns": [["synthetic tx"] * n_seqs, names, [str(len(s)) for s in seqs], seqs]}
    return [cfg, data], codes


def combined_library(n_transcripts=3000, scale=1.0, seed=4, config=None):
    """BASELINE config 4 shape: MHC-like + KIR-like allele families + immune-gene transcripts in ONE
    library, union feature-calling (intersect_level 0).  `scale` shrinks the two allele families
    (founders x alleles) for tests.  Returns (library_json_obj, codes)."""
    fa = max(2, int(round(40 * scale)))
    fb = max(2, int(round(17 * scale)))
    lib_a, codes_a = allele_family_library(n_founders=fa, alleles_per_founder=50, length=1098, snps_mean=15.0, seed=1)
    lib_b, codes_b = allele_family_library(n_founders=fb, alleles_per_founder=90, length=1350, snps_mean=12.0, seed=3,
                                           name_prefix="KIR")
    lib_c, codes_c = random_transcript_library(n_seqs=n_transcripts, mean_len=2000, family_frac=0.3, seed=seed)
    cfg = dict(lib_a[0], intersect_level=0)
    if config:
        cfg.update(config)
    data = {"headers": lib_a[1]["headers"],
            "columns": [lib_a[1]["columns"][j] + lib_b[1]["columns"][j] + lib_c[1]["columns"][j] for j in range(4)]}
    return [cfg, data], codes_a + codes_b + codes_c


def _revcomp_codes(a):
    return (3 - a[:, ::-1]).astype(np.uint8)


def sample_reads(codes, n_reads, read_len=90, err_rate=0.005, off_target=0.2, rc_frac=0.1,
                 n_frac=0.001, seed=2, ref_idx=None):
    """Reads as a [n, read_len] uint8 ASCII matrix sampled uniformly from the alleles in `codes`.
    Substitution errors at `err_rate`, `off_target` random reads, `rc_frac` reverse-complemented,
    `n_frac` of reads get one 'N'.  Returns (ascii_matrix, truth_ref[int32; -1 off-target])."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for c in codes], np.int64)
    starts_of = np.zeros(len(codes) + 1, np.int64)
    np.cumsum(lens, out=starts_of[1:])
    flat = np.concatenate(codes)
    ok = np.nonzero(lens >= read_len)[0]
    ref = rng.choice(ok, size=n_reads) if ref_idx is None else np.asarray(ref_idx)
    span = lens[ref] - read_len + 1
    pos = (rng.random(n_reads) * span).astype(np.int64)
    out = np.empty((n_reads, read_len), np.uint8)
    step = 1 << 20
    ar = np.arange(read_len, dtype=np.int64)
    for a in range(0, n_reads, step):
        b = min(n_reads, a + step)
        base = starts_of[ref[a:b]] + pos[a:b]
        out[a:b] = flat[base[:, None] + ar[None, :]]
    n_err = rng.binomial(n_reads * read_len, err_rate)
    if n_err:
        ei = rng.integers(0, n_reads, size=n_err)
        ej = rng.integers(0, read_len, size=n_err)
        out[ei, ej] = (out[ei, ej] + rng.integers(1, 4, size=n_err, dtype=np.uint8)) & 3
    truth = ref.astype(np.int32)
    off = rng.random(n_reads) < off_target
    n_off = int(off.sum())
    if n_off:
        out[off] = rng.integers(0, 4, size=(n_off, read_len), dtype=np.uint8)
        truth[off] = -1
    rc = rng.random(n_reads) < rc_frac
    if rc.any():
        out[rc] = _revcomp_codes(out[rc])
    asc = _ASCII[out]
    n_n = int(n_reads * n_frac)
    if n_n:
        asc[rng.integers(0, n_reads, size=n_n), rng.integers(0, read_len, size=n_n)] = ord("N")
    return asc, truth


def sample_pairs(codes, n_pairs, read_len=150, insert_mean=300, insert_sd=50, err_rate=0.005,
                 off_target=0.1, seed=3):
    """Paired-end reads: mate 1 forward at the fragment start, mate 2 reverse-complemented at the
    fragment end (half of the fragments flipped).  Returns (r1_ascii, r2_ascii, truth_ref)."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for c in codes], np.int64)
    starts_of = np.zeros(len(codes) + 1, np.int64)
    np.cumsum(lens, out=starts_of[1:])
    flat = np.concatenate(codes)
    ok = np.nonzero(lens >= read_len + 20)[0]
    ref = rng.choice(ok, size=n_pairs)
    ins = np.clip(rng.normal(insert_mean, insert_sd, n_pairs).astype(np.int64), read_len, None)
    ins = np.minimum(ins, lens[ref])
    pos = (rng.random(n_pairs) * (lens[ref] - ins + 1)).astype(np.int64)
    ar = np.arange(read_len, dtype=np.int64)
    base = starts_of[ref] + pos
    m1 = flat[base[:, None] + ar[None, :]]
    m2 = _revcomp_codes(flat[(base + ins - read_len)[:, None] + ar[None, :]])
    for m in (m1, m2):
        n_err = rng.binomial(m.size, err_rate)
        ei = rng.integers(0, n_pairs, size=n_err)
        ej = rng.integers(0, read_len, size=n_err)
        m[ei, ej] = (m[ei, ej] + rng.integers(1, 4, size=n_err, dtype=np.uint8)) & 3
    truth = ref.astype(np.int32)
    off = rng.random(n_pairs) < off_target
    n_off = int(off.sum())
    if n_off:
        m1[off] = rng.integers(0, 4, size=(n_off, read_len), dtype=np.uint8)
        m2[off] = rng.integers(0, 4, size=(n_off, read_len), dtype=np.uint8)
        truth[off] = -1
    flip = rng.random(n_pairs) < 0.5
    m1[flip], m2[flip] = m2[flip].copy(), m1[flip].copy()
    return _ASCII[m1], _ASCII[m2], truth


def barcodes_10x(n_reads, n_cells=10000, reads_per_umi=3.0, seed=2, truth=None):
    """(CB, UB) as 2-bit packed integers: cb = 16-mer -> uint32, ub = 12-mer -> 24 bits.
    Reads of one UMI are contiguous runs in a random permutation; cells log-normal sized.
    Returns key[u8] = cb << 32 | ub."""
    rng = np.random.default_rng(seed + 1000)
    n_umi = max(1, int(n_reads / reads_per_umi))
    w = rng.lognormal(0.0, 1.0, n_cells)
    cell_cb = rng.integers(0, 1 << 32, size=n_cells, dtype=np.uint64)
    umi_cell = rng.choice(n_cells, size=n_umi, p=w / w.sum())
    umi_ub = rng.integers(0, 1 << 24, size=n_umi, dtype=np.uint64)
    read_umi = rng.integers(0, n_umi, size=n_reads)
    if truth is not None:
        # a UMI is one molecule: give all of its reads to the UMI that matches their allele where possible
        order = np.argsort(truth, kind="stable")
        read_umi_sorted = np.sort(read_umi)
        read_umi = np.empty_like(read_umi)
        read_umi[order] = read_umi_sorted
    return (cell_cb[umi_cell[read_umi]] << np.uint64(32)) | umi_ub[read_umi]


def unpack_barcode(v, n):
    """2-bit packed integer -> ACGT string of n bases (most significant base first)."""
    return "".join("ACGT"[(int(v) >> (2 * (n - 1 - i))) & 3] for i in range(n))


def barcode_workload(n_reads, n_whitelist=737_280, n_cells=10000, cb_length=16, err_rate=0.003, n_rate=0.0005,
                     off_whitelist=0.01, seed=9, clustered=0.0):
    """10x-like raw cell barcodes for the fastq-to-bam stage: a whitelist of random ACGT cb_length-mers,
    `n_cells` of them in use (log-normal sizes), per-base substitution errors (low quality at the
    error), occasional N, and a fraction of barcodes that are not on the whitelist at all.
    `clustered` = fraction of whitelist entries that are 1-2 substitutions from another entry
    (several correction candidates -> the quality rule and the first-read cache rule matter).
    Returns (whitelist uint8[n_wl, L] ASCII, cb uint8[n, L] ASCII, qual uint8[n, L] phred)."""
    rng = np.random.default_rng(seed)
    L = cb_length
    wl = rng.integers(0, 4, size=(int(n_whitelist * 1.02) + 8, L), dtype=np.uint8)
    n_cl = int(clustered * len(wl))
    if n_cl:
        src = rng.integers(0, len(wl) - n_cl, size=n_cl)
        nb = wl[src].copy()
        for _ in range(2):
            pos = rng.integers(0, L, size=n_cl)
            nb[np.arange(n_cl), pos] = (nb[np.arange(n_cl), pos] + rng.integers(1, 4, size=n_cl)) % 4
        wl[len(wl) - n_cl:] = nb
    wl = np.unique(wl, axis=0)
    wl = wl[rng.permutation(len(wl))[:n_whitelist]]
    w = rng.lognormal(0.0, 1.0, n_cells)
    cells = rng.choice(len(wl), size=n_cells, replace=False)
    cb = wl[cells[rng.choice(n_cells, size=n_reads, p=w / w.sum())]].copy()
    qual = rng.integers(25, 41, size=(n_reads, L), dtype=np.uint8)
    err = rng.random((n_reads, L)) < err_rate
    cb[err] = (cb[err] + rng.integers(1, 4, size=int(err.sum()), dtype=np.uint8)) % 4
    qual[err] = rng.integers(2, 20, size=int(err.sum()), dtype=np.uint8)
    off = rng.random(n_reads) < off_whitelist
    cb[off] = rng.integers(0, 4, size=(int(off.sum()), L), dtype=np.uint8)
    asc = _ASCII[cb]
    nmask = rng.random((n_reads, L)) < n_rate
    asc[nmask] = ord("N")
    qual[nmask] = 2
    return _ASCII[wl], asc, qual
