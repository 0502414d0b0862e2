"""Structured fuzzing of the zstd compressor/decompressor kernel sources under the test-only SIMT emulator:
many data shapes (runs, short periods, long repeats, binary alphabets > 128 symbols, tiny alphabets, ramps,
mixtures, ragged lengths) must round-trip, decode with libzstd, and libzstd's frames must decode with ours."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as ora
import tsgpu
from tsgpu import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT_LIB = os.path.join(ROOT, "tests", "simt", "libtsgpu_simt.so")
Z = tsgpu.FLAG_ZSTD


@pytest.fixture(scope="module")
def ctx():
    subprocess.check_call(["make", "-s", "-C", ROOT, "tests/simt/libtsgpu_simt.so"])
    c = tsgpu.Context(max_chunk_bytes=200000, max_batch=4, lib_path=SIMT_LIB)
    yield c
    c.close()


def shapes(rng, n):
    yield "zeros", np.zeros(n, np.uint8)
    yield "ones_run_then_random", np.concatenate([np.full(n // 2, 255, np.uint8), rng.integers(0, 256, n - n // 2, dtype=np.uint8)])
    yield "period2", np.tile(np.array([7, 200], np.uint8), n // 2 + 1)[:n]
    yield "period3", np.tile(np.array([1, 2, 3], np.uint8), n // 3 + 1)[:n]
    yield "period31", np.tile(rng.integers(0, 256, 31, dtype=np.uint8), n // 31 + 1)[:n]
    yield "period33", np.tile(rng.integers(0, 256, 33, dtype=np.uint8), n // 33 + 1)[:n]
    yield "period1000", np.tile(rng.integers(0, 256, 1000, dtype=np.uint8), n // 1000 + 1)[:n]
    yield "binary_alphabet_200", rng.integers(0, 200, n).astype(np.uint8)
    yield "skewed_high_bytes", (255 - np.minimum(rng.geometric(0.15, n), 120)).astype(np.uint8)
    yield "two_symbols", rng.integers(0, 2, n).astype(np.uint8) * 97
    yield "ramp", (np.arange(n) & 0xff).astype(np.uint8)
    yield "text", corpus.gen_chunk("K", int(rng.integers(0, 1000)), 0, n)
    t = corpus.gen_chunk("K", 3, 1, n).copy()
    t[rng.integers(0, n, max(1, n // 50))] = rng.integers(0, 256, max(1, n // 50), dtype=np.uint8)
    yield "text_with_noise", t
    words = [bytes(rng.integers(128, 256, int(k), dtype=np.uint8)) for k in rng.integers(2, 12, 40)]
    yield "high_byte_words", np.frombuffer(b"".join(words[i] for i in rng.integers(0, 40, n // 4 + 8)), np.uint8)[:n].copy()
    blk = rng.integers(0, 256, 3000, dtype=np.uint8)
    yield "long_repeat_far", np.concatenate([blk, rng.integers(0, 256, 5000, dtype=np.uint8), blk, blk])[:n] if n >= 14000 else np.tile(blk, 2)[:n]


@pytest.mark.parametrize("mode", [0, tsgpu.FLAG_ZSTD_DENSE])        # both compressors: independent blocks / 64 KiB regions
@pytest.mark.parametrize("n", [700, 8191, 8192, 8193, 40000, 123457])
def test_simt_fuzz_shapes(ctx, n, mode):
    rng = np.random.default_rng(n)
    for name, src in shapes(rng, n):
        src = np.ascontiguousarray(src[:n])
        m = src.size
        for cs in (0, 50000):
            out, sizes = ctx.transform(Z | mode, src, cs)
            c = cs if cs else m
            pos = 0
            for i, s in enumerate(sizes):
                want = src[i * c:min(m, (i + 1) * c)]
                assert ora.zstd_decompress_chunk(out[pos:pos + s]) == want.tobytes(), (name, n, cs, i)
                pos += s
            back, _ = ctx.detransform(Z, out, sizes, m)
            assert np.array_equal(back, src), (name, n, cs)
        ref, rs = ora.transform_segment(Z, src, 0)
        back, _ = ctx.detransform(Z, ref, rs, m)
        assert np.array_equal(back, src), (name, n, "libzstd frame")


@pytest.mark.parametrize("level", [-5, 1, 6, 12, 19])
def test_simt_decodes_other_libzstd_levels(ctx, level):
    # other strategies (fast, lazy, btopt/btultra): different block splitting, repeat-offset density, table reuse
    rng = np.random.default_rng(level + 100)
    for name, src in shapes(rng, 150000):
        src = np.ascontiguousarray(src[:150000])
        frame = np.frombuffer(ora.zstd_compress_level(src, level), dtype=np.uint8)
        back, _ = ctx.detransform(Z, frame, [frame.size], src.size)
        assert np.array_equal(back, src), (name, level)
