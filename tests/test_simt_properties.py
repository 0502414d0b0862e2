"""Property tests (hypothesis) of the emulated compressor / decompressors on structured random inputs: literal runs, repeats at
chosen distances, small alphabets, byte ramps, glued in random order.  Properties: libzstd reads our frames; we read ours
(fast path) and libzstd's at random levels (serial general path and, in a second context, the parallel general path)."""
import os
import subprocess

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as ora
import tsgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT_LIB = os.path.join(ROOT, "tests", "simt", "libtsgpu_simt.so")
Z = tsgpu.FLAG_ZSTD


@st.composite
def structured_bytes(draw):
    rng = np.random.default_rng(draw(st.integers(0, 2**32 - 1)))
    parts, total = [], 0
    for _ in range(draw(st.integers(1, 12))):
        kind = draw(st.sampled_from(["rand", "run", "alpha", "repeat", "ramp", "text"]))
        n = draw(st.one_of(st.integers(1, 300), st.integers(3000, 40000)))
        if kind == "rand":
            p = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == "run":
            p = np.full(n, draw(st.integers(0, 255)), dtype=np.uint8)
        elif kind == "alpha":
            k = draw(st.integers(2, 40))
            p = rng.integers(0, k, n, dtype=np.uint8) * draw(st.integers(1, 6)) + draw(st.integers(0, 15))
        elif kind == "ramp":
            p = (np.arange(n) * draw(st.integers(1, 7)) % 251).astype(np.uint8)
        elif kind == "text":
            words = [bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(50)]
            p = np.frombuffer(b" ".join(words[int(i)] for i in rng.zipf(1.4, n // 4 + 1) % 50)[:n].ljust(n, b"."), dtype=np.uint8)
        else:                                                # repeat an earlier stretch at some distance
            have = np.concatenate(parts) if parts else rng.integers(0, 256, 64, dtype=np.uint8)
            d = draw(st.integers(1, max(1, min(have.size, 70000))))
            seed = have[-d:]
            p = np.resize(seed, n)
        parts.append(p.astype(np.uint8))
        total += n
    out = np.concatenate(parts)
    reps = draw(st.sampled_from([1, 1, 1, 3, 9]))            # now and then long enough for several 128 KiB libzstd blocks
    if reps > 1 and out.size * reps <= (1 << 20):
        out = np.concatenate([out] + [np.roll(out, int(k) * 7) for k in range(1, reps)])
    return out


@pytest.fixture(scope="module")
def ctxs():
    subprocess.check_call(["make", "-s", "-C", ROOT, "tests/simt/libtsgpu_simt.so"])
    a = tsgpu.Context(max_chunk_bytes=1 << 20, max_batch=2, lib_path=SIMT_LIB)
    yield a
    a.close()


EXAMPLES = int(os.environ.get("TSGPU_HYP_EXAMPLES", "40"))     # a soak run sets it higher (scripts/README.md)


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture, HealthCheck.data_too_large])
@given(src=structured_bytes(), level=st.sampled_from([1, 3, 7, 19]))
def test_compressor_and_decode_paths(ctxs, src, level):
    serial = ctxs
    n = src.size
    out, sizes = serial.transform(Z, src, 0)
    assert ora.zstd_content_size(out[:sizes[0]]) == n
    assert ora.zstd_decompress_chunk(out[:sizes[0]]) == src.tobytes()
    back, _ = serial.detransform(Z, out, sizes, n)
    assert np.array_equal(back, src)
    dout, dsizes = serial.transform(Z | tsgpu.FLAG_ZSTD_DENSE, src, 0)          # the dense compressor: same contract
    assert ora.zstd_content_size(dout[:dsizes[0]]) == n
    assert ora.zstd_decompress_chunk(dout[:dsizes[0]]) == src.tobytes()
    back, _ = serial.detransform(Z, dout, dsizes, n)
    assert np.array_equal(back, src)
    ref = np.frombuffer(ora.zstd_compress_level(src, level), dtype=np.uint8)
    back, osz = serial.detransform(Z, ref, [ref.size], n)
    assert osz == [n] and np.array_equal(back, src)
