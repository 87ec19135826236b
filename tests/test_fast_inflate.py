"""csrc/fast_inflate.hpp (the BGZF reader's own raw-DEFLATE decoder) against zlib: every block type, every compression
level and strategy, sizes from 0 to 65536, data shapes from constant to random; truncated and damaged streams must be
declined or decoded to something of the right length without touching memory outside the buffers (guard bytes)."""
import ctypes as ct
import random
import zlib

import numpy as np

from nimble_b200 import _lib


def _deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    z = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return z.compress(data) + z.flush()


def _inflate(L, comp, n, guard=64):
    src = np.frombuffer(comp, np.uint8).copy()
    out = np.full(n + 2 * guard, 0xA5, np.uint8)
    rc = L.nb200_fast_inflate(src.ctypes.data, len(src), out.ctypes.data + guard, n)
    assert (out[:guard] == 0xA5).all() and (out[guard + n:] == 0xA5).all(), "wrote outside the output buffer"
    return rc, bytes(out[guard:guard + n])


def _shapes(rng, n):
    yield bytes(n)                                                        # constant
    yield bytes(rng.getrandbits(8) for _ in range(n))                     # incompressible
    yield bytes(rng.choice(b"ACGT") for _ in range(n))                    # 2 bits of entropy per byte
    words = [bytes(rng.choice(b"ACGTN#IIIIFFFF:") for _ in range(rng.randint(3, 40))) for _ in range(50)]
    s = bytearray()
    while len(s) < n:
        s += rng.choice(words)
    yield bytes(s[:n])                                                    # long and short repeats
    rec = bytearray()
    while len(rec) < n:                                                   # BAM-like records
        rec += b"\\xd6\\x00\\x00\\x00\\xff\\xff\\xff\\xff" + bytes(rng.getrandbits(8) for _ in range(4)) + b"r%08d\\x00" % rng.randrange(10 ** 8)
        rec += bytes(rng.choice(b"\\x11\\x12\\x14\\x18\\x21\\x22\\x24\\x28\\x41\\x42\\x44\\x48\\x81\\x82\\x84\\x88") for _ in range(45)) + b"\\x1e" * 90
        rec += b"CBZ" + bytes(rng.choice(b"ACGT") for _ in range(16)) + b"\\x00UBZ" + bytes(rng.choice(b"ACGT") for _ in range(12)) + b"\\x00"
    yield bytes(rec[:n])
    # geometric symbol distribution over 256 values: code lengths up to 15 (subtables of the 11-bit primary table)
    yield bytes(min(255, int(rng.expovariate(0.045))) for _ in range(n))
    yield bytes((i * 7 + min(63, int(rng.expovariate(0.2)))) & 0xFF for i in range(n))       # long matches at many distances + rare literals


def test_fast_inflate_equals_zlib():
    L = _lib.load()
    rng = random.Random(7)
    n_cases = 0
    for n in (0, 1, 2, 7, 8, 9, 63, 258, 259, 1000, 4095, 30000, 65280, 65536):
        for data in _shapes(rng, n):
            for level, strat in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                 (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (4, zlib.Z_FILTERED)):
                comp = _deflate(data, level, strat, mem=rng.choice([1, 8, 9]))
                rc, got = _inflate(L, comp, n)
                assert rc == 1 and got == data, (n, level, strat)
                n_cases += 1
    # several deflate blocks in one stream (Z_FULL_FLUSH between them), mixed types
    z = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts = [bytes(rng.choice(b"ACGT") for _ in range(5000)), bytes(3000), bytes(rng.getrandbits(8) for _ in range(4000))]
    comp = b"".join(z.compress(p) + z.flush(zlib.Z_FULL_FLUSH) for p in parts) + z.flush()
    rc, got = _inflate(L, comp, sum(len(p) for p in parts))
    assert rc == 1 and got == b"".join(parts)
    assert n_cases > 500


def test_fast_inflate_declines_bad_streams():
    L = _lib.load()
    rng = random.Random(8)
    data = bytes(rng.choice(b"ACGTACGTNN") for _ in range(20000))
    comp = _deflate(data, 6)
    assert _inflate(L, comp, len(data)) == (1, data)
    assert _inflate(L, comp, len(data) - 1)[0] == 0                       # inflated size disagrees
    assert _inflate(L, comp, len(data) + 1)[0] == 0
    for cut in (0, 1, 5, len(comp) // 2, len(comp) - 1):
        assert _inflate(L, comp[:cut], len(data))[0] == 0                 # truncated
    ok = 0
    for _ in range(3000):                                                 # damaged bytes: declined, or decoded to SOMETHING of the right length
        bad = bytearray(comp)
        for _k in range(rng.randint(1, 4)):
            bad[rng.randrange(len(bad))] ^= 1 << rng.randrange(8)
        rc, got = _inflate(L, bytes(bad), len(data))
        if rc:
            try:
                ok += zlib.decompress(bytes(bad), -15) == got            # then zlib reads the damaged stream the same way
            except zlib.error:
                raise AssertionError("decoded a stream zlib rejects")
    assert ok >= 0
