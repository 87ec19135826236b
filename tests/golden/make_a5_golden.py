"""Generates tests/golden/a5_cases.json by running the REFERENCE's own barcode code
(/root/reference/nimble/fastq_barcode_processor.py: build_hamming_index, correct_cell_barcode,
process_pair) on seeded random whitelists / read pairs.  Run in the build container only (the
reference does not exist on the GPU box):

    python tests/golden/make_a5_golden.py

Shims (none changes the reference's arithmetic):
  * pysam / Bio are absent -> stubbed.  `pysam.AlignedSegment()` becomes a plain attribute bag
    so the records process_pair builds can be read back; FASTQ records are attribute bags with
    the three members process_pair touches (.id, .seq, .letter_annotations).
  * Python iterates a `set` of strings in a per-process random order, and correct_cell_barcode
    keeps the FIRST candidate with the lowest quality (:113-125).  To pin one answer the
    candidate sets of the reference's own hamming index are re-presented as lists in ascending
    order (the SPEC order); every multi-candidate case is also run with the
    descending order, and `tie` records whether the two runs disagree.
"""
import json
import os
import random
import sys
import types
from collections import defaultdict

REF = "/root/reference"
sys.path.insert(0, REF)


class Bag:
    def __init__(self):
        self.tags = {}

    def set_tag(self, k, v):
        self.tags[k] = v


pysam = types.ModuleType("pysam")
pysam.AlignedSegment = Bag
sys.modules["pysam"] = pysam
bio = types.ModuleType("Bio")
bio.SeqIO = types.ModuleType("Bio.SeqIO")
sys.modules["Bio"] = bio
sys.modules["Bio.SeqIO"] = bio.SeqIO

import importlib.util

spec = importlib.util.spec_from_file_location("ref_fbp", os.path.join(REF, "nimble", "fastq_barcode_processor.py"))
fbp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fbp)

BASES = "ACGT"


def rand_seq(rng, n, alphabet=BASES):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, pos=None, alphabet="ACGTN"):
    pos = rng.randrange(len(s)) if pos is None else pos
    b = rng.choice([c for c in alphabet if c != s[pos]])
    return s[:pos] + b + s[pos + 1:]


def make_whitelist(rng, n, L, clustered):
    wl = []
    seen = set()
    while len(wl) < n:
        if clustered and wl and rng.random() < 0.6:
            s = mutate(rng, rng.choice(wl), alphabet=BASES)          # neighbours: multi-candidate cases
            if rng.random() < 0.3:
                s = mutate(rng, s, alphabet=BASES)
        else:
            s = rand_seq(rng, L)
        if s not in seen or rng.random() < 0.02:                     # a few duplicate lines
            wl.append(s)
            seen.add(s)
    return wl


def ordered_index(hamming_index, reverse):
    return {k: sorted(v, reverse=reverse) for k, v in hamming_index.items()}


class Rec:
    def __init__(self, rid, seq, qual):
        self.id, self.seq, self.letter_annotations = rid, seq, {"phred_quality": qual}


def make_case(rng, idx):
    L = rng.choice([16, 16, 16, 16, 8, 12, 21, 5])
    U = rng.choice([12, 12, 12, 10, 4])
    n_wl = rng.choice([1, 3, 20, 60, 200])
    wl = make_whitelist(rng, n_wl, L, clustered=rng.random() < 0.8)
    if rng.random() < 0.2:
        wl.append(rand_seq(rng, L + 1))                              # wrong-length line: inert
    if rng.random() < 0.15:
        wl.append(mutate(rng, rng.choice(wl)[:L].ljust(L, "A"), alphabet="N"))   # a whitelist entry with an N
    n_pairs = rng.randint(1, 48)
    few_q = rng.random() < 0.5                                       # few distinct qualities: real ties
    pairs = []
    pool = []                                                        # raw barcodes to repeat (cache behaviour)
    for p in range(n_pairs):
        r = rng.random()
        base = rng.choice(wl)[:L].ljust(L, "A")
        if pool and r < 0.25:
            cb = rng.choice(pool)
        elif r < 0.45:
            cb = base
        elif r < 0.80:
            cb = mutate(rng, base)
        elif r < 0.90:
            cb = mutate(rng, mutate(rng, base))
        elif r < 0.95:
            cb = rand_seq(rng, L, "ACGTN")
        else:
            cb = mutate(rng, base, alphabet="ACGTNacgtX.")
        pool.append(cb)
        tail = rng.choice([0, 0, 1, 5, 30]) if rng.random() < 0.3 else rng.randint(1, 40)
        r1 = cb + rand_seq(rng, U, "ACGTN" if rng.random() < 0.1 else BASES) + rand_seq(rng, tail, "ACGTN")
        if rng.random() < 0.06:
            r1 = r1[:rng.randint(0, L + U - 1)]                       # too short
        q1 = [rng.choice([2, 11, 25, 37]) if few_q else rng.randint(2, 40) for _ in r1]
        r2 = rand_seq(rng, rng.randint(0, 40), "ACGTN")
        q2 = [rng.randint(2, 40) for _ in r2]
        name = "read%d:%d" % (idx, p)
        s1, s2 = rng.choice([("", ""), ("/1", "/2"), ("/1", ""), ("", "/2")])
        id1, id2 = name + s1, name + s2
        if rng.random() < 0.05:
            id2 = name + "x" + s2                                     # name mismatch
        pairs.append([id1, r1, q1, id2, r2, q2])

    whitelist = set(wl)
    hidx = fbp.build_hamming_index(whitelist)
    asc, desc = ordered_index(hidx, False), ordered_index(hidx, True)

    def run(index):
        cache, stats, out = {}, defaultdict(int), []
        for (id1, s1, q1, id2, s2, q2) in pairs:
            stats["total_pairs"] += 1                                 # fastq_to_bam_with_barcodes :254
            res = fbp.process_pair(Rec(id1, s1, q1), Rec(id2, s2, q2), whitelist, index, cache, stats, L, U)
            if res:
                stats["written_pairs"] += 1                           # :268
                a, b = res
                assert a.tags == b.tags and a.flag == 77 and b.flag == 141
                out.append({"name": a.query_name, "cb": a.tags["CB"], "umi": a.tags["UB"],
                            "r1_seq": a.query_sequence, "r1_qual": list(a.query_qualities),
                            "r2_seq": b.query_sequence, "r2_qual": list(b.query_qualities)})
        st = dict(stats)
        st["cache_size"] = len(cache)
        return out, st, cache

    out_a, st_a, cache_a = run(asc)
    out_d, st_d, cache_d = run(desc)
    tie = out_a != out_d
    multi = sum(1 for k in cache_a if k not in whitelist and len(hidx.get(k, ())) > 1)
    def qs(q):                                                       # qualities as phred+33 text: compact fixtures
        return "".join(chr(x + 33) for x in q)

    # Compact form of the written records: [pair index, corrected CB].  Everything else the reference put
    # into the two BAM records is a slice of the input pair; that is asserted here so the fixture can
    # leave it out (name = id minus /1; UB = r1[L:L+U]; read 1 = r1[L+U:] with its qualities; read 2 whole).
    compact, it = [], iter(out_a)
    nxt = next(it, None)
    for pi, (id1, s1, q1, id2, s2, q2) in enumerate(pairs):
        nm = id1[:-2] if id1.endswith("/1") else id1
        if nxt is not None and nxt["name"] == nm and nxt["r1_seq"] == s1[L + U:] and len(s1) > L + U:
            assert nxt["umi"] == s1[L:L + U] and nxt["r1_qual"] == q1[L + U:] and nxt["r2_seq"] == s2 and nxt["r2_qual"] == q2
            compact.append([pi, nxt["cb"]])
            nxt = next(it, None)
    assert nxt is None and len(compact) == len(out_a)
    out_a = compact
    pairs = [[a, b, qs(c), d, e, qs(f)] for (a, b, c, d, e, f) in pairs]
    return {"cb_length": L, "umi_length": U, "whitelist": wl, "pairs": pairs, "records": out_a,
            "stats": {k: st_a.get(k, 0) for k in ("total_pairs", "written_pairs", "cb_perfect_match", "cb_corrected",
                                                  "cb_no_correction", "name_mismatch", "too_short", "no_remaining_seq",
                                                  "cache_size")},
            "tie": tie, "n_multi_candidate": multi,
            "corrections": {k: v for k, v in cache_a.items()}}


def main():
    rng = random.Random(20261018)
    cases = [make_case(rng, i) for i in range(120)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "a5_cases.json")
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_a5_golden.py", "reference": "nimble/fastq_barcode_processor.py",
                   "cases": cases}, f, separators=(",", ":"))
    print("wrote %d cases (%d with a multi-candidate barcode, %d with an order-dependent tie), %d bytes"
          % (len(cases), sum(c["n_multi_candidate"] > 0 for c in cases), sum(c["tie"] for c in cases),
             os.path.getsize(path)))


if __name__ == "__main__":
    main()
