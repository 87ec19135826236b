"""The streaming file pipeline behind nb200_align_files (csrc/stream.cpp) on inputs that span many slabs, and the same
pipeline over several GPUs from one process (nb200_align_files_multi): outputs byte-identical to the one-GPU run."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

from nimble_b200 import frontend, synth
from oracle import oracle as O
from helpers import to_concat

pytestmark = pytest.mark.gpu


def _inputs(tmp_path, n_reads, seed):
    import file_bench
    lib, codes = synth.allele_family_library(n_founders=8, alleles_per_founder=20, length=700, snps_mean=10, seed=seed)
    r1, truth = synth.sample_reads(codes, n_reads, read_len=90, seed=seed + 1)
    key = synth.barcodes_10x(n_reads, n_cells=500, seed=seed + 1, truth=truth)
    lib_path = str(tmp_path / "lib.json")
    with open(lib_path, "w") as f:
        json.dump(lib, f)
    bam = str(tmp_path / "in.bam")
    file_bench.write_bam(bam, r1, key)
    return lib, lib_path, bam, r1, key


def test_stream_many_slabs_matches_oracle(engine, tmp_path):
    """300 k reads = three slabs, BGZF input in many chunks, plain and gzip output: every row of the per-read TSV is the
    oracle's call for that read, in input order."""
    n = 300_000
    lib, lib_path, bam, r1, key = _inputs(tmp_path, n, 71)
    out, out_gz = str(tmp_path / "o.tsv"), str(tmp_path / "o.tsv.gz")
    assert frontend.align(lib_path, out, [bam], 4, "unstranded", "", None, engine=engine) == 0
    assert frontend.align(lib_path, out_gz, [bam], 4, "unstranded", "", None, engine=engine) == 0
    text = open(out).read()
    assert gzip.open(out_gz, "rt").read() == text
    lo = O.Library(lib, k=20)
    ro, fo = O.align(lo, to_concat(r1))
    rows = text.split("\n")
    assert rows[0].split("\t")[:2] == ["nimble_features", "nimble_score"] and rows[-1] == ""
    called = np.nonzero(ro["n_feat"])[0]
    assert len(rows) - 2 == len(called) > 100_000
    acgt = "ACGT"
    for row, i in zip(rows[1:-1:997], called[::997]):          # a sample of rows, all columns
        f = row.split("\t")
        assert f[0] == ",".join(lo.features[j] for j in fo[i, :ro["n_feat"][i]])
        assert f[2:6] == [str(int(x)) for x in ro["score"][i]]
        assert f[6] == "r%09d" % i
        cb, ub = int(key[i] >> np.uint64(32)), int(key[i] & np.uint64(0xFFFFFF))
        assert f[7] == "".join(acgt[(cb >> (2 * (15 - j))) & 3] for j in range(16))
        assert f[8] == "".join(acgt[(ub >> (2 * (11 - j))) & 3] for j in range(12))
    names = [r.split("\t", 7)[6] for r in rows[1:-1]]
    assert names == ["r%09d" % i for i in called]                # every called read once, in input order


def test_multi_gpu_file_align_equals_single_gpu(tmp_path):
    """nb200_align_files_multi (one process, a context per GPU, slabs dealt to the GPUs) writes the same bytes as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 1_200_000
    lib, lib_path, bam, r1, key = _inputs(tmp_path, n, 81)
    one, two = str(tmp_path / "one.tsv"), str(tmp_path / "two.tsv")
    assert frontend.align(lib_path, one, [bam], 0, "unstranded", "", None) == 0
    assert frontend.align(lib_path, two, [bam], 0, "unstranded", "", None, gpus=min(torch.cuda.device_count(), 8)) == 0
    a, b = open(one, "rb").read(), open(two, "rb").read()
    assert len(a) > 10_000_000 and a == b
