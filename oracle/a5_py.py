"""ORACLE (test infrastructure, never shipped, never on the product path).

Pure-Python restatement of nimble's `fastq-to-bam` barcode stage ("A5" in SURVEY.md §8a):
10x R1 -> (cell barcode, UMI, remainder), whitelist correction of the cell barcode, the
per-pair skip rules and the statistics.  `oracle/cb_oracle.c` is the fast restatement and is
checked against this file.

Pinned against: tests/golden/a5_cases.json, produced by running the REFERENCE's own
functions (build_hamming_index, correct_cell_barcode, process_pair) in
tests/golden/make_a5_golden.py.

Follows (reference file:line, nimble/fastq_barcode_processor.py):
  build_hamming_index        :17-36    variant -> whitelist entries one substitution away;
                                        the substituted base is one of A C G T N
  correct_cell_barcode       :73-128   cache -> exact -> no candidate -> single -> lowest quality
  parse_10x_barcode_from_r1  :131-141
  process_pair               :144-209  name check, too_short, no_remaining_seq, statistics

Two places where the reference's outcome depends on run-time accidents are frozen here
(DESIGN.md §2.8):
  * `correction_cache` is filled by whichever read reaches it first (thread timing with
    num_cores > 1).  SPEC: reads are processed in file order, so the FIRST eligible read
    carrying a raw barcode decides its correction (its qualities break ties) for all
    later reads with the same raw barcode.
  * With several candidates the loop keeps the first one with the strictly lowest
    quality (:113-125), "first" being Python's per-process randomised set order.
    SPEC: candidates are visited in ascending string order.
"""
from __future__ import annotations

ALPHABET = "ACGTN"

# status codes shared with the C ABI (include/nimble_b200.h)
CB_SKIPPED, CB_PERFECT, CB_CORRECTED, CB_NONE = 0, 1, 2, 3


class Whitelist:
    """load_cb_whitelist (:38-71): the set of valid barcodes; index = first occurrence in the file."""

    def __init__(self, entries, cb_length=16):
        self.cb_length = cb_length
        self.entries = list(entries)
        self.index = {}
        for i, e in enumerate(self.entries):
            if e not in self.index:
                self.index[e] = i

    def candidates(self, raw):
        """hamming_index.get(raw) (:17-36, :99): whitelist entries that differ from `raw` in exactly
        one position, where raw's base at that position is one of A C G T N.  Ascending order."""
        out = []
        for i, ch in enumerate(raw):
            if ch not in ALPHABET:
                continue
            for b in ALPHABET:
                if b != ch:
                    v = raw[:i] + b + raw[i + 1:]
                    if v in self.index:
                        out.append(v)
        return sorted(out)


def correct_cell_barcode(raw_cb, quals, wl: Whitelist, cache):
    """:73-128.  Returns the corrected barcode string or None."""
    if raw_cb in cache:
        return cache[raw_cb]
    if raw_cb in wl.index:
        cache[raw_cb] = raw_cb
        return raw_cb
    cands = wl.candidates(raw_cb)
    if not cands:
        cache[raw_cb] = None
        return None
    if len(cands) == 1:
        cache[raw_cb] = cands[0]
        return cands[0]
    best, lowest = None, float("inf")
    for cand in cands:                       # ascending string order (SPEC)
        for i, (a, b) in enumerate(zip(raw_cb, cand)):
            if a != b:
                if quals[i] < lowest:
                    lowest = quals[i]
                    best = cand
                break
    cache[raw_cb] = best
    return best


def fastq_id(header_line):
    """Bio.SeqIO 'fastq' record.id: the title line after '@' up to the first whitespace."""
    t = header_line[1:] if header_line.startswith("@") else header_line
    parts = t.split(None, 1)
    return parts[0] if parts else ""


def removesuffix(s, suf):
    return s[:-len(suf)] if suf and s.endswith(suf) else s


def process_pairs(pairs, wl: Whitelist, cb_length=16, umi_length=12):
    """process_pair (:144-209) over pairs in file order.

    pairs: iterable of (r1_id, r1_seq, r1_qual, r2_id, r2_seq, r2_qual), quals = lists of ints.
    Returns (records, stats): records = one dict per written pair (name, cb, umi, r1_seq, r1_qual,
    r2_seq, r2_qual), stats = the counters the reference prints (:284-309)."""
    stats = {k: 0 for k in ("total_pairs", "written_pairs", "cb_perfect_match", "cb_corrected", "cb_no_correction",
                            "name_mismatch", "too_short", "no_remaining_seq")}
    cache = {}
    out = []
    for (id1, s1, q1, id2, s2, q2) in pairs:
        stats["total_pairs"] += 1
        n1, n2 = removesuffix(id1, "/1"), removesuffix(id2, "/2")
        if n1 != n2:
            stats["name_mismatch"] += 1
            continue
        if len(s1) < cb_length + umi_length:
            stats["too_short"] += 1
            continue
        raw_cb, umi, rest = s1[:cb_length], s1[cb_length:cb_length + umi_length], s1[cb_length + umi_length:]
        if len(rest) == 0:
            stats["no_remaining_seq"] += 1
            continue
        cb = correct_cell_barcode(raw_cb, q1[:cb_length], wl, cache)
        if cb is None:
            stats["cb_no_correction"] += 1
            continue
        stats["cb_perfect_match" if cb == raw_cb else "cb_corrected"] += 1
        stats["written_pairs"] += 1
        out.append({"name": n1, "cb": cb, "umi": umi, "r1_seq": rest, "r1_qual": list(q1[cb_length + umi_length:]),
                    "r2_seq": s2, "r2_qual": list(q2)})
    stats["cache_size"] = len(cache)
    return out, stats


def correct_batch(cbs, quals, eligible, wl: Whitelist):
    """Array form used by the parity tests: returns (idx, status) lists; idx = whitelist index or -1."""
    cache = {}
    idx, status = [], []
    for i, raw in enumerate(cbs):
        if eligible is not None and not eligible[i]:
            idx.append(-1); status.append(CB_SKIPPED)
            continue
        cb = correct_cell_barcode(raw, quals[i], wl, cache)
        if cb is None:
            idx.append(-1); status.append(CB_NONE)
        else:
            idx.append(wl.index[cb]); status.append(CB_PERFECT if cb == raw else CB_CORRECTED)
    return idx, status, len(cache)
