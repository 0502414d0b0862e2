"""-m gpu parity tests: the CUDA path (through the C-ABI of libtsgpu.so) against the CPU oracle on the same
seeded inputs.  Mirrors the reference's own test pyramid: operator round trips
(core/T/transform/*Test.java), the TransformsEndToEndTest chunk-size grid, the IT's independent per-chunk
decrypt/decompress (core/IT/RemoteStorageManagerTest.java:327-381) and its ranged-fetch grid (:383-423).
Bars: AES-GCM ciphertext and tags bit-exact; zstd frames cross-decodable both ways with Frame_Content_Size;
round trip byte-exact; ChunkIndex inputs identical."""
import numpy as np
import pytest

import tsgpu
from tsgpu import binding, corpus
from oracle import oracle as ora

pytestmark = pytest.mark.gpu

Z, A, D = tsgpu.FLAG_ZSTD, tsgpu.FLAG_AES, tsgpu.FLAG_ZSTD_DENSE      # D modifies Z: the dense (region) compressor
MIB = 1 << 20


@pytest.fixture(scope="module")
def ctx():
    c = tsgpu.Context(max_chunk_bytes=4 * MIB, max_batch=8)
    yield c
    c.close()


@pytest.fixture(scope="module")
def small_ctx():
    c = tsgpu.Context(max_chunk_bytes=400000, max_batch=64)
    yield c
    c.close()


def _material(rng, nch):
    return rng.bytes(32), rng.bytes(32), rng.bytes(12 * max(nch, 1))


# ------------------------------------------------------------------ AES-256-GCM: bit-exact vs OpenSSL
@pytest.mark.parametrize("kind", ["R", "K", "Z"])
def test_aes_bit_exact_4mib_chunks(ctx, kind):
    n, cs = 5 * 4 * MIB - 12345, 4 * MIB          # short final chunk
    src = corpus.gen_segment(kind, 1, n, cs)
    key, aad, ivs = corpus.fixed_key_material(5)
    got, gs = ctx.transform(A, src, cs, key, aad, ivs)
    want, ws = ora.transform_segment(A, src, cs, key, aad, ivs)
    assert gs == ws == [cs + 28] * 4 + [cs - 12345 + 28]
    assert np.array_equal(got, want)
    back, osz = ctx.detransform(A, want, ws, n, key, aad)
    assert osz == [cs] * 4 + [cs - 12345] and np.array_equal(back, src)


@pytest.mark.parametrize("n,cs", [(1, 0), (15, 0), (16, 0), (17, 0), (31, 7), (4096, 1024), (100001, 4099),
                                  (262144 * 3 + 1, 262144), (400000, 0)])
def test_aes_bit_exact_ragged(small_ctx, n, cs):
    rng = np.random.default_rng(n * 31 + cs)
    src = rng.integers(0, 256, n, dtype=np.uint8)
    nch = (n + cs - 1) // cs if cs else 1
    key, aad, ivs = _material(rng, nch)
    got, gs = small_ctx.transform(A, src, cs, key, aad, ivs)
    want, ws = ora.transform_segment(A, src, cs, key, aad, ivs)
    assert gs == ws and np.array_equal(got, want)
    back, _ = small_ctx.detransform(A, got, gs, n, key, aad)
    assert np.array_equal(back, src)


def test_aes_tag_mismatch_is_an_error_and_releases_nothing(ctx):
    # DecryptionChunkEnumeration.java:59-61: AEADBadTagException -> RuntimeException
    n, cs = 2 * MIB + 5, MIB
    src = corpus.gen_segment("R", 2, n, cs)
    key, aad, ivs = corpus.fixed_key_material(3)
    enc, sizes = ctx.transform(A, src, cs, key, aad, ivs)
    for where in (0, 11, 12, len(enc) // 2, len(enc) - 1):     # IV, ciphertext, tag
        bad = enc.copy()
        bad[where] ^= 1
        dst = np.full(n, 0x55, dtype=np.uint8)
        with pytest.raises(tsgpu.TsgpuError) as e:
            ctx.detransform(A, bad, sizes, n, key, aad, dst=dst)
        assert e.value.code == binding.E_AUTH
    with pytest.raises(tsgpu.TsgpuError) as e:             # wrong AAD
        ctx.detransform(A, enc, sizes, n, key, bytes(32))
    assert e.value.code == binding.E_AUTH
    with pytest.raises(tsgpu.TsgpuError) as e:             # "Stream has fewer bytes than expected"
        ctx.detransform(A, enc[:-1], sizes, n, key, aad)
    assert e.value.code == binding.E_SHORT and "fewer bytes" in str(e.value)


def test_aes_empty_aad_and_odd_aad(small_ctx):
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, 5000, dtype=np.uint8)
    for alen in (0, 1, 16, 20, 33, 100):
        key, aad, ivs = rng.bytes(32), rng.bytes(alen), rng.bytes(12)
        got, gs = small_ctx.transform(A, src, 0, key, aad, ivs)
        want, ws = ora.transform_segment(A, src, 0, key, aad, ivs)
        assert np.array_equal(got, want)


# ------------------------------------------------------------------ zstd: cross-decodability both ways
@pytest.mark.parametrize("mode", [0, D])
@pytest.mark.parametrize("kind", ["K", "R", "Z"])
def test_zstd_gpu_frames_decode_with_libzstd(ctx, kind, mode):
    n, cs = 3 * 4 * MIB - 777, 4 * MIB
    src = corpus.gen_segment(kind, 3, n, cs)
    got, gs = ctx.transform(Z | mode, src, cs)
    pos = 0
    for i, s in enumerate(gs):
        frame = got[pos:pos + s]
        lo, hi = i * cs, min(n, (i + 1) * cs)
        assert ora.zstd_content_size(frame) == hi - lo          # Frame_Content_Size is mandatory for the reader
        assert ora.zstd_decompress_chunk(frame) == src[lo:hi].tobytes()
        pos += s
    assert pos == len(got)
    if kind == "K":
        assert sum(gs) < n // 2
    back, osz = ctx.detransform(Z, got, gs, n)
    assert np.array_equal(back, src)


@pytest.mark.parametrize("kind", ["K", "R", "Z"])
def test_zstd_gpu_decodes_libzstd_level3_frames(ctx, kind):
    n, cs = 3 * 4 * MIB - 777, 4 * MIB
    src = corpus.gen_segment(kind, 4, n, cs)
    ref, rs = ora.transform_segment(Z, src, cs)
    back, osz = ctx.detransform(Z, ref, rs, n)
    assert osz == [cs, cs, cs - 777]
    assert np.array_equal(back, src)


def test_zstd_corrupt_frames_are_errors(ctx):
    src = corpus.gen_segment("K", 5, MIB, MIB)
    ref, rs = ora.transform_segment(Z, src, MIB)
    for where, val in ((0, 0x00), (4, 0x04), (len(ref) // 2, None), (len(ref) - 1, None)):
        bad = ref.copy()
        bad[where] = val if val is not None else bad[where] ^ 0xFF
        try:
            back, _ = ctx.detransform(Z, bad, rs, MIB)
        except tsgpu.TsgpuError as e:
            assert e.code == binding.E_CORRUPT
        else:
            # a flipped literal byte can still be a valid frame; it just must not crash or overrun
            assert len(back) == MIB


# ------------------------------------------------------------------ full chain: TransformsEndToEndTest grid
@pytest.mark.parametrize("flags", [A, Z, Z | A, Z | D, Z | A | D])
@pytest.mark.parametrize("cs", [0, 1, 2, 3, 5, 13, 1024, 2048, 5123, 181200 - 1, 181200 * 2])
def test_transforms_end_to_end_grid(small_ctx, flags, cs):
    n = 181200 if cs == 0 or cs >= 13 else 700           # tiny chunk sizes: keep the chunk count sane
    rng = np.random.default_rng(99)
    src = rng.integers(0, 256, n, dtype=np.uint8)
    if flags & Z:                                        # half compressible, half random
        src[: n // 2] = corpus.gen_chunk("K", 9, 0, n // 2)
    nch = (n + cs - 1) // cs if cs else 1
    key, aad, ivs = _material(rng, nch)
    got, gs = small_ctx.transform(flags, src, cs, key, aad, ivs)
    assert len(gs) == nch
    # the reference-side reader recovers the segment from our bytes (IT: RemoteStorageManagerTest.java:327-381)
    back_ref, _ = ora.detransform_chunks(flags & 3, got, gs, n, key, aad)      # (the reader does not care how it was compressed)
    assert np.array_equal(back_ref, src)
    # and we read what the reference writes
    ref, rs = ora.transform_segment(flags & 3, src, cs, key, aad, ivs)
    back, osz = small_ctx.detransform(flags, ref, rs, n, key, aad)
    assert np.array_equal(back, src)
    back2, _ = small_ctx.detransform(flags, got, gs, n, key, aad)
    assert np.array_equal(back2, src)


@pytest.mark.parametrize("mode", [0, D])
@pytest.mark.parametrize("kind", ["K", "R"])
def test_full_pipeline_4mib(ctx, kind, mode):
    n, cs = 16 * 4 * MIB + 4321, 4 * MIB                  # 17 chunks -> 3 batches of 8
    src = corpus.gen_segment(kind, 6, n, cs)
    key, aad, ivs = corpus.fixed_key_material(17)
    got, gs = ctx.transform(Z | A | mode, src, cs, key, aad, ivs)
    back_ref, _ = ora.detransform_chunks(Z | A, got, gs, n, key, aad)
    assert np.array_equal(back_ref, src)
    back, osz = ctx.detransform(Z | A, got, gs, n, key, aad)
    assert np.array_equal(back, src) and sum(osz) == n
    # ChunkIndex inputs: identical encoding for the same size list
    assert binding.chunk_index_json(cs, n, None, sizes=gs) .startswith('{"type":"variable","originalChunkSize":4194304')
    assert ora.transformed_chunks_deserialize(binding.transformed_chunks_serialize(gs)) == gs
    pos = ctx.chunk_positions(gs)
    assert [int(p) for p in pos[:-1]] == [c[3] for c in ora.ChunkIndex.variable(cs, n, gs).chunks()]
    assert int(pos[-1]) == len(got)


def test_ranged_fetch_window(ctx):
    # config C5: detransform only the chunks covering [S, S + 16 MiB) and cut with the FetchChunkEnumeration plan
    n, cs = 12 * 4 * MIB, 4 * MIB
    src = corpus.gen_segment("K", 7, n, cs)
    key, aad, ivs = corpus.fixed_key_material(12)
    obj, sizes = ctx.transform(Z | A, src, cs, key, aad, ivs)
    idx = ora.ChunkIndex.variable(cs, n, sizes)
    chunks = idx.chunks()
    for start in (0, 16 * MIB, n - 16 * MIB, 8 * MIB + 12345):
        end = min(n - 1, start + 16 * MIB - 1)
        plan = idx.fetch_plan(start, end)
        first, last = plan[0][0], plan[-1][0]
        lo = chunks[first][3]
        hi = chunks[last][3] + chunks[last][4]
        part, osz = ctx.detransform(Z | A, obj[lo:hi], sizes[first:last + 1], (last - first + 1) * cs, key, aad)
        out = []
        off = 0
        for (cid, skip, take), o in zip(plan, osz):
            out.append(part[off + skip: off + skip + take])
            off += o
        assert np.array_equal(np.concatenate(out), src[start:end + 1])


def test_deserialize_reference_written_index(ctx):
    # a manifest written by the reference carries a libzstd-compressed size list
    sizes = [1000000 + (i * 7919) % 5000 for i in range(2000)]
    s = ora.transformed_chunks_serialize(sizes)
    assert ctx.transformed_chunks_deserialize(s) == sizes
    assert ctx.transformed_chunks_deserialize("KLUv/SAPeQAAAAAAAwAAAAoBAAoAAAAe") == [10, 20, 30]


def test_transformed_chunks_compressed_like_the_reference(ctx):
    # TransformedChunksSerializer.java:40-48: the manifest field is Base64(zstd(codec bytes)); the ctx-taking serializer
    # compresses with the dense kernel — both readers accept it, the golden vector is unchanged, sizes track libzstd's
    rng = np.random.default_rng(5)
    for name, sizes in {"equal": [4194332] * 256, "clustered": (1350000 + rng.integers(-70000, 70000, 256)).tolist(),
                        "tight": (1350000 + rng.integers(-700, 700, 4096)).tolist(), "golden": [10, 20, 30]}.items():
        mine, ref, raw = ctx.transformed_chunks_serialize(sizes), ora.transformed_chunks_serialize(sizes), binding.transformed_chunks_serialize(sizes)
        assert ora.transformed_chunks_deserialize(mine) == sizes and ctx.transformed_chunks_deserialize(mine) == sizes, name
        assert len(mine) <= len(raw), name
        if name == "golden":
            assert mine == ref == "KLUv/SAPeQAAAAAAAwAAAAoBAAoAAAAe"
        else:
            assert len(mine) <= len(ref) + max(40, len(ref) // 10), (name, len(mine), len(ref), len(raw))
    js = ctx.chunk_index_json(4 * MIB, 255 * 4 * MIB + 5, None, sizes=[4194332] * 256)
    assert js.startswith('{"type":"variable","originalChunkSize":4194304,"originalFileSize":%d,"transformedChunks":"' % (255 * 4 * MIB + 5))


def test_frame_header_fields_refused_like_libzstd(ctx):
    # header fields a reader without a dictionary has to refuse, exactly where libzstd does: a Dictionary_ID other than 0
    # ("Dictionary mismatch"), a Window_Descriptor above windowLog 31; a zero Dictionary_ID field and windowLog 31 are fine
    src = corpus.gen_segment("K", 1, 600000, 600000)
    f = np.frombuffer(ora.zstd_compress_level(src, 1), dtype=np.uint8).copy()
    assert f[4] & 0x20 == 0                                          # not Single_Segment: byte 5 is the Window_Descriptor
    small = corpus.gen_segment("K", 1, 9000, 9000)
    f2 = np.frombuffer(ora.zstd_compress_level(small, 3), dtype=np.uint8).copy()
    assert f2[4] & 0x20
    u8 = lambda *v: np.array(v, np.uint8)
    cases = [("plain", f, src), ("windowLog 32", np.concatenate([f[:5], u8(22 << 3), f[6:]]), src),
             ("windowLog 31", np.concatenate([f[:5], u8(21 << 3), f[6:]]), src),
             ("dictionary id 7", np.concatenate([f[:4], u8(f[4] | 1), f[5:6], u8(7), f[6:]]), src),
             ("dictionary id field 0", np.concatenate([f[:4], u8(f[4] | 1), f[5:6], u8(0), f[6:]]), src),
             ("single segment, dictionary id 257", np.concatenate([f2[:4], u8(f2[4] | 2), u8(1, 1), f2[5:]]), small),
             ("reserved bit", np.concatenate([f2[:4], u8(f2[4] | 8), f2[5:]]), small)]
    for name, frame, want in cases:
        frame = np.ascontiguousarray(frame)
        try:
            ref_ok = ora.zstd_decompress_chunk(frame) == want.tobytes()
        except Exception:
            ref_ok = False
        try:
            back, _ = ctx.detransform(Z, frame, [frame.size], want.size)
            mine_ok = bool(np.array_equal(back, want))
        except tsgpu.TsgpuError as e:
            assert e.code == binding.E_CORRUPT
            mine_ok = False
        assert mine_ok == ref_ok, (name, ref_ok, mine_ok)
        assert ref_ok == (name in ("plain", "windowLog 31", "dictionary id field 0")), name


def test_content_checksum_is_verified_like_libzstd(small_ctx):
    # zstd-jni's Zstd.decompress verifies the Content_Checksum (low 32 bits of XXH64 of the content) when a frame announces one; the
    # reference's writer never does, a foreign writer (zstd CLI default) may: accepted when right, refused when wrong — as libzstd does
    for n in (1, 31, 32, 33, 100, 8192, 70001, 300000):
        src = corpus.gen_segment("K", n % 97, n, n)
        for level in (1, 3):
            f = np.frombuffer(ora.zstd_compress_checksum(src, level), dtype=np.uint8)
            assert f[4] & 4
            back, _ = small_ctx.detransform(Z, f, [f.size], n)
            assert np.array_equal(back, src), (n, level)
            bad = f.copy()
            bad[-1 - (n % 4)] ^= 0x21
            with pytest.raises(Exception):
                ora.zstd_decompress_chunk(bad)
            with pytest.raises(tsgpu.TsgpuError) as e:
                small_ctx.detransform(Z, bad, [bad.size], n)
            assert e.value.code == binding.E_CORRUPT


def test_bytes_after_the_frame_are_refused_like_libzstd(small_ctx):
    # DecompressionChunkEnumeration.java:41-45 hands the whole chunk to Zstd.decompress(chunk, size): libzstd refuses bytes after the
    # frame ("Src size is incorrect") and a second frame ("Destination buffer is too small"), and steps over skippable frames
    src = corpus.gen_segment("K", 3, 9000, 9000)
    skippable = np.frombuffer(bytes([0x53, 0x2A, 0x4D, 0x18, 3, 0, 0, 0, 9, 9, 9]), dtype=np.uint8)
    frames = {"libzstd": np.frombuffer(ora.zstd_compress_level(src, 3), dtype=np.uint8)}
    for name, flags in (("speed", Z), ("dense", Z | tsgpu.FLAG_ZSTD_DENSE)):
        out, sizes = small_ctx.transform(flags, src, 0)
        frames[name] = np.array(out[:sizes[0]], copy=True)
    for name, f in frames.items():
        for what, tail, accepted in (("exact", np.zeros(0, np.uint8), True), ("garbage", np.array([1, 2, 3, 4, 5], np.uint8), False),
                                     ("one byte", np.zeros(1, np.uint8), False), ("second frame", f, False),
                                     ("skippable frame", skippable, True), ("truncated skippable frame", skippable[:9], False)):
            chunk = np.concatenate([f, tail])
            try:
                ref_ok = ora.zstd_decompress_chunk(chunk) == src.tobytes()
            except Exception:
                ref_ok = False
            assert ref_ok == accepted, (name, what)                       # the expectation IS libzstd's behaviour
            try:
                back, _ = small_ctx.detransform(Z, chunk, [chunk.size], 9000)
                mine_ok = np.array_equal(back, src)
            except tsgpu.TsgpuError as e:
                assert e.code == binding.E_CORRUPT
                mine_ok = False
            assert mine_ok == accepted, (name, what)


def test_empty_segment_and_argument_errors(ctx):
    out, sizes = ctx.transform(Z | A, np.zeros(0, np.uint8), 4 * MIB, bytes(32), b"", bytes(12))
    assert sizes == [] and len(out) == 0
    with pytest.raises(tsgpu.TsgpuError) as e:
        ctx.transform(A, np.zeros(10, np.uint8), 8 * MIB + 1, bytes(32), b"", bytes(12))
    assert e.value.code == binding.E_ARG


def test_single_process_multi_device_context():
    # the JVM case: one process, several GPUs; batches are dealt round-robin to (device, slot) pairs
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    c = tsgpu.Context(max_chunk_bytes=MIB, max_batch=4, devices=[0, 1])
    n, cs = 37 * MIB + 999, MIB                      # 38 chunks -> 10 batches over 2 devices x 4 slots
    src = corpus.gen_segment("K", 8, n, cs)
    key, aad, ivs = corpus.fixed_key_material(38)
    got, gs = c.transform(Z | A, src, cs, key, aad, ivs)
    one = tsgpu.Context(max_chunk_bytes=MIB, max_batch=4, devices=[0])
    want, ws = one.transform(Z | A, src, cs, key, aad, ivs)
    assert gs == ws and np.array_equal(got, want)     # same bytes whichever device handled a batch
    back, _ = c.detransform(Z | A, got, gs, n, key, aad)
    assert np.array_equal(back, src)
    c.close(); one.close()


@pytest.mark.parametrize("mode", [0, D])
@pytest.mark.parametrize("shape", ["skewed_high_bytes", "gauss_around_128", "alphabet_200", "period_1000", "long_runs"])
def test_zstd_binary_payload_shapes(ctx, shape, mode):
    # binary Kafka payloads: alphabets above byte value 128 (FSE-compressed Huffman weights), literal-only blocks
    # (zero sequences), long-distance repeats and long runs (warp-wide match extension)
    rng = np.random.default_rng(42)
    n, cs = 2 * 4 * MIB + 4097, 4 * MIB
    if shape == "skewed_high_bytes":
        src = (255 - np.minimum(rng.geometric(0.15, n), 120)).astype(np.uint8)
    elif shape == "gauss_around_128":
        src = np.clip(rng.normal(128, 20, n), 0, 255).astype(np.uint8)
    elif shape == "alphabet_200":
        src = rng.integers(0, 200, n).astype(np.uint8)
    elif shape == "period_1000":
        src = np.tile(rng.integers(0, 256, 1000, dtype=np.uint8), n // 1000 + 1)[:n].copy()
    else:
        src = np.repeat(rng.integers(0, 256, n // 3000 + 1, dtype=np.uint8), 3000)[:n].copy()
    got, gs = ctx.transform(Z | mode, src, cs)
    pos = 0
    for i, s in enumerate(gs):
        assert ora.zstd_decompress_chunk(got[pos:pos + s]) == src[i * cs:min(n, (i + 1) * cs)].tobytes()
        pos += s
    if shape in ("skewed_high_bytes", "gauss_around_128", "period_1000", "long_runs"):
        assert sum(gs) < 0.9 * n
    back, _ = ctx.detransform(Z, got, gs, n)
    assert np.array_equal(back, src)
    ref, rs = ora.transform_segment(Z, src, cs)
    back, _ = ctx.detransform(Z, ref, rs, n)
    assert np.array_equal(back, src)


def test_index_files_ride_one_ragged_batch(ctx):
    # RemoteStorageManager.transformIndex (RemoteStorageManager.java:455-490): each Kafka index file is ONE chunk,
    # encryption only; fetchIndexBytes (:624-652) reads one back
    rng = np.random.default_rng(77)
    blobs = [rng.integers(0, 256, n, dtype=np.uint8) for n in (10 * MIB // 8, 10 * MIB // 12, 37, 126, 20480)]
    src = np.concatenate(blobs)
    key, aad, ivs = rng.bytes(32), rng.bytes(32), rng.bytes(12 * 5)
    out, sizes = ctx.transform_chunks(A, src, [b.size for b in blobs], key, aad, ivs)
    assert sizes == [b.size + 28 for b in blobs]
    pos = 0
    for i, b in enumerate(blobs):
        assert bytes(out[pos:pos + sizes[i]]) == ora.aesgcm_encrypt_chunk(key, ivs[12 * i:12 * i + 12], aad, b)
        pos += sizes[i]
    back, osz = ctx.detransform(A, out, sizes, src.size, key, aad)
    assert np.array_equal(back, src) and osz == [b.size for b in blobs]


def test_large_chunk_multi_pass_assembly():
    # chunk.size is configurable up to 2^30-1 (RemoteStorageManagerConfig.java:123-130): a 20 MiB chunk has 2560 zstd
    # blocks (the frame assembler scans 1024 block sizes per pass) and 80 AES ranges
    c = tsgpu.Context(max_chunk_bytes=20 * MIB, max_batch=2)
    n, cs = 2 * 20 * MIB + 54321, 20 * MIB
    src = corpus.gen_segment("K", 10, n, 4 * MIB)
    key, aad, ivs = corpus.fixed_key_material(3)
    got, gs = c.transform(Z | A, src, cs, key, aad, ivs)
    back_ref, _ = ora.detransform_chunks(Z | A, got, gs, n, key, aad)
    assert np.array_equal(back_ref, src)
    back, _ = c.detransform(Z | A, got, gs, n, key, aad)
    assert np.array_equal(back, src)
    ref, rs = ora.transform_segment(Z | A, src, cs, key, aad, ivs)
    back, _ = c.detransform(Z | A, ref, rs, n, key, aad)
    assert np.array_equal(back, src)
    c.close()
