"""Run by tests/test_simt_asan.py in a subprocess with libasan / libubsan preloaded (argv[1] = the library to load): both directions of the emulated kernels over a
sweep of shapes (block/chunk boundaries, incompressible and constant data, binary alphabets, ragged batches, AES)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import tsgpu  # noqa: E402
from tsgpu import corpus  # noqa: E402
from oracle import oracle as ora  # noqa: E402

lib = os.path.join(ROOT, "tests", "simt", sys.argv[1] if len(sys.argv) > 1 else "libtsgpu_simt_asan.so")
c = tsgpu.Context(max_chunk_bytes=1 << 18, max_batch=4, lib_path=lib)
rng = np.random.default_rng(3)
cases = []
for kind in "KRZ":
    for n in (1, 5, 300, 8191, 8192, 8193, 20000, 70000):
        cases.append(corpus.gen_segment(kind, 0, n, n))
b = rng.integers(0, 256, 50000, dtype=np.uint8); b[::3] = 200; b[1::7] = 131       # byte values above 128: FSE-coded weights
cases.append(b)
cases.append((rng.integers(0, 4, 40000, dtype=np.uint8) * 60 + 3).astype(np.uint8))    # four symbols
cases.append(np.repeat(rng.integers(0, 256, 700, dtype=np.uint8), 37)[:25000].copy())  # long runs
key, aad = rng.bytes(32), rng.bytes(32)
for src in cases:
    for flags in (1, 3, 5, 7):                                                          # 4 = dense compressor (64 KiB regions)
        ivs = rng.bytes(12)
        out, sizes = c.transform(flags, src, 0, key, aad, ivs)
        back, _ = ora.detransform_chunks(flags & 3, out, sizes, src.size, key, aad)
        assert np.array_equal(back, src)
        back, _ = c.detransform(flags & 3, out, sizes, src.size, key, aad)
        assert np.array_equal(back, src)
    ref = np.frombuffer(ora.zstd_compress_chunk(src), dtype=np.uint8)                   # libzstd frames: general path
    back, _ = c.detransform(1, ref, [ref.size], src.size)
    assert np.array_equal(back, src)
seg = corpus.gen_segment("K", 1, 100000, 30000)                                        # several chunks, two batches
for flags in (3, 7):
    out, sizes = c.transform(flags, seg, 30000, key, aad, rng.bytes(12 * 4))
    back, _ = c.detransform(3, out, sizes, seg.size, key, aad)
    assert np.array_equal(back, seg)
big_seg = corpus.gen_segment("K", 2, (1 << 18) + 4097, 1 << 18)                        # a full-size chunk: 4 regions + a ragged tail chunk
out, sizes = c.transform(5, big_seg, 1 << 18)
back, _ = c.detransform(1, out, sizes, big_seg.size)
assert np.array_equal(back, big_seg)
blobs = [rng.integers(0, 256, n, dtype=np.uint8) for n in (10485, 37, 126)]
out, sizes = c.transform_chunks(2, np.concatenate(blobs), [x.size for x in blobs], key, aad, rng.bytes(36))
assert sizes == [x.size + 28 for x in blobs]
bad = out.copy(); bad[20] ^= 1                                                         # tag mismatch path
try:
    c.detransform(2, bad, sizes, sum(x.size for x in blobs), key, aad)
    raise SystemExit("tag mismatch not detected")
except tsgpu.TsgpuError:
    pass
c.close()
# chunks larger than the reading context's chunk size — as a tampered manifest or a config mismatch
# would present them — must be refused before the GCM kernel writes plaintext anywhere (AES only: 4 x 65856 B plaintexts
# into a context sized for 65536 B chunks used to overrun d_orig by 592 B).
big = tsgpu.Context(max_chunk_bytes=1 << 17, max_batch=4, lib_path=lib)
small = tsgpu.Context(max_chunk_bytes=65536, max_batch=4, lib_path=lib)
src = rng.integers(0, 256, 4 * 65856, dtype=np.uint8)
out, sizes = big.transform(2, src, 65856, key, aad, rng.bytes(48))
for flags in (2, 3):
    try:
        small.detransform(flags, out, sizes, src.size + 4096, key, aad)
        raise SystemExit("oversize chunk accepted (flags %d)" % flags)
    except tsgpu.TsgpuError as e:
        # AES only: refused up front; AES+zstd: the decrypted bytes still fit a frame slot but are no zstd frame
        assert e.code == (tsgpu.binding.E_ARG if flags == 2 else tsgpu.binding.E_CORRUPT), e
big.close(); small.close()
print("sanitizer sweep ok", len(cases))
