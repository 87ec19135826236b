"""Multi-rank host logic on CPU: world_size-2 gloo.  Reads are sharded by
cell-barcode hash, every rank aggregates its own shard (the oracle stands in for the GPU path),
count tables are all-gathered and merged; the result must equal the single-rank table."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from nimble_b200 import shard, synth  # noqa: E402
from nimble_b200.engine import CountTable  # noqa: E402


def test_shard_is_a_function_of_the_cell_only():
    rng = np.random.default_rng(1)
    cells = rng.integers(0, 1 << 32, size=500, dtype=np.uint64)
    key = (cells[rng.integers(0, 500, size=20000)] << np.uint64(32)) | rng.integers(0, 1 << 24, size=20000, dtype=np.uint64)
    for world in (1, 2, 4, 8):
        r = shard.shard_of_key(key, world)
        assert r.min() >= 0 and r.max() < world
        by_cell = {}
        for c, x in zip((key >> np.uint64(32)).tolist(), r.tolist()):
            assert by_cell.setdefault(c, x) == x
        parts = shard.split_by_rank(key, world)
        assert sum(len(p) for p in parts) == len(key)
        if world > 1:
            sizes = np.array([len(p) for p in parts])
            assert sizes.min() > 0.5 * sizes.mean()          # balanced enough
    nb = np.array([0xFFFFFFFFFFFFFFFF], np.uint64)
    assert shard.shard_of_key(nb, 8)[0] == 0


def _oracle_table(lib_json, reads, key):
    from oracle import oracle as O
    from helpers import oracle_counts, to_concat
    lo = O.Library(lib_json)
    ro, fo = O.align(lo, to_concat(reads), n_threads=2)
    cell, cnt, off, ids, dropped = oracle_counts(lo, ro, fo, key)
    return lo, CountTable(cell, cnt, off.astype(np.uint32), ids, dropped, int((ro["n_feat"] > 0).sum()), 0)


def _worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib, codes = synth.allele_family_library(n_founders=4, alleles_per_founder=6, length=400, snps_mean=6, seed=61)
    reads, truth = synth.sample_reads(codes, 6000, read_len=90, seed=62)
    key = synth.barcodes_10x(len(reads), n_cells=40, seed=62, truth=truth)
    mine = shard.split_by_rank(key, world)[rank]
    lo, table = _oracle_table(lib, reads[mine], key[mine])
    width = lo.cfg.max_hits_to_report
    m = torch.from_numpy(shard.table_to_tensor_rows(table, width))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([m.shape[0]], dtype=torch.int64))
    mx = int(max(int(s) for s in sizes))
    pad = torch.full((mx, m.shape[1]), -1, dtype=torch.int64)
    pad[:m.shape[0]] = m
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    dropped = torch.tensor([table.dropped_empty], dtype=torch.int64)
    dist.all_reduce(dropped)
    if rank == 0:
        mats = [g[:int(s)].numpy() for g, s in zip(gathered, sizes)]
        merged = shard.merge_rank_tables(mats, lo.tok_end, lo.tok_comma)
        _, full = _oracle_table(lib, reads, key)
        want = shard.table_to_tensor_rows(full, width)
        out_q.put((np.array_equal(merged, want), int(dropped) == full.dropped_empty, len(want), [len(x) for x in mats]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_rank():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, dropped_ok, n_rows, per_rank = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and dropped_ok and n_rows > 50 and min(per_rank) > 0
