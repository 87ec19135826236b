"""TEST ORACLE ONLY (never imported by nimble_b200/): restatement of the read trimming `nimble align --trim
<TARGET_LENGTH>:<STRICTNESS>` asks the aligner for (nimble/__main__.py:191-192,400; defaults nimble/types.py:24-25).

PARITY UNPINNED.  The arithmetic lives in the un-vendored nimble-aligner binary.  The flag's two parameters are
exactly those of Trimmomatic's MAXINFO step (targetLength, strictness), so the published MaxInfo criterion is
restated here (Bolger, Lohse & Usadel 2014, Bioinformatics 30:2114, Supplementary methods): keep the prefix of
length l that maximises

    log( 1 / (1 + e^(target - l)) )  +  (1 - strictness) * log(l)  +  strictness * sum_{j < l} log(p_correct(q_j))

with p_correct(q) = 1 - 10^(-(q + 0.5) / 10), qualities clamped to 0..59, the longer prefix on a tie; a read whose
best prefix is empty is dropped (length 0).  float64, terms added in this order, left to right."""
import math

MAXQ = 60


def tables(target, strictness, max_len=500):
    length_score = [math.log(1.0 / (1.0 + math.exp(float(target - i - 1)))) + (1.0 - strictness) * math.log(float(i + 1)) for i in range(max_len)]
    qual_score = [strictness * math.log(1.0 - math.pow(0.1, (0.5 + q) / 10.0)) for q in range(MAXQ)]
    return length_score, qual_score


def trim_maxinfo(quals, target, strictness):
    """quals: phred values (ints).  Returns the number of leading bases to keep."""
    ls, qs = tables(target, strictness, max(1, len(quals)))
    acc, best, pos = 0.0, -math.inf, 0
    for i, q in enumerate(quals):
        acc += qs[min(max(int(q), 0), MAXQ - 1)]
        score = ls[i] + acc
        if score >= best:
            best, pos = score, i + 1
    return pos
