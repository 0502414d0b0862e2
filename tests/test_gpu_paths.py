"""-m gpu tests of which path runs: key tables across calls, libzstd-written frames through the whole-frame executor, this
library's frames region by region, identical frames on repeated runs."""
import numpy as np
import pytest

from oracle import oracle as ora
import tsgpu
from tsgpu import corpus

Z, A = tsgpu.FLAG_ZSTD, tsgpu.FLAG_AES


@pytest.mark.gpu
def test_gpu_key_tables_follow_the_key_across_calls():
    # the slot's GHASH tables are kept while the key stays the same and rebuilt when it changes (bit-exact either way)
    rng = np.random.default_rng(99)
    src = rng.integers(0, 256, 3 * (1 << 20) + 5, dtype=np.uint8)
    k1, k2, aad = rng.bytes(32), rng.bytes(32), rng.bytes(32)
    ivs = rng.bytes(12 * 4)
    ctx = tsgpu.Context(max_chunk_bytes=1 << 20, max_batch=2)
    try:
        for key in (k1, k1, k2, k1, k2, k2):
            want, wsizes = ora.transform_segment(ora.FLAG_AES, src, 1 << 20, key, aad, ivs)
            got, gsizes = ctx.transform(A, src, 1 << 20, key, aad, ivs)
            assert gsizes == wsizes and np.array_equal(got, want)
            back, _ = ctx.detransform(A, want, wsizes, src.size, key, aad)
            assert np.array_equal(back, src)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_decode_paths_libzstd_frames_and_own_regions():
    # libzstd-written frames: entropy stage per block in parallel + one CTA executing the frame; own frames: region by region
    ctx = tsgpu.Context(max_chunk_bytes=4 << 20, max_batch=4)
    try:
        ctx.profile_enable(True)
        for kind, n, level in (("K", 4 << 20, 3), ("K", 3000000, 19), ("K", 1500000, 1), ("R", 600000, 3), ("Z", 2000000, 3)):
            src = corpus.gen_chunk(kind, 3, 0, n)
            frame = np.frombuffer(ora.zstd_compress_level(src, level), dtype=np.uint8)
            back, osz = ctx.detransform(Z, frame, [frame.size], n)
            assert osz == [n] and np.array_equal(back, src), (kind, n, level)
        st0 = ctx.decode_path_stats()
        assert st0["whole_frames"] == 5 and st0["regions"] == 0 and st0["serial_frames"] == 0
        # one batch: two libzstd frames, a speed-mode frame (a warp per self-contained block), a dense-mode frame (a CTA per region)
        srcs = [corpus.gen_chunk("K", 9 + i, 0, 1 << 20) for i in range(2)]
        frames = [np.frombuffer(ora.zstd_compress_chunk(s), dtype=np.uint8) for s in srcs]
        mine, msz = ctx.transform(Z, srcs[0], 0)
        frames.append(mine[:msz[0]])
        dense, dsz = ctx.transform(Z | 4, srcs[1], 0)
        frames.append(dense[:dsz[0]])
        back, _ = ctx.detransform(Z, np.concatenate(frames), [f.size for f in frames], 4 << 20)
        assert np.array_equal(back, np.concatenate(srcs + srcs))
        names = set(ctx.profile_report())
        assert {"zstd_dec_entropy", "zstd_dec_frame_exec", "zstd_dec_regions", "zstd_dec_blk_literals", "zstd_dec_blk_sequences", "zstd_dec_blk_exec"} <= names
        st1 = ctx.decode_path_stats()
        assert st1["whole_frames"] == 7 and st1["regions"] == 16 and st1["blocks"] == 128 and st1["region_fallback_frames"] == 0
        rng = np.random.default_rng(4)
        for trial in range(40):
            bad = frames[1].copy()
            bad[int(rng.integers(0, bad.size))] ^= 1 << int(rng.integers(0, 8))
            try:
                ctx.detransform(Z, bad, [bad.size], 1 << 20)
            except tsgpu.TsgpuError as e:
                assert e.code == tsgpu.binding.E_CORRUPT
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_frames_are_identical_across_runs():
    # dense mode: hash-slot winners are deterministic (highest position of a step) — the same input gives the same object
    ctx = tsgpu.Context(max_chunk_bytes=4 << 20, max_batch=8)
    try:
        src = corpus.gen_segment("K", 5, 24 << 20, 4 << 20)
        first, fs = ctx.transform(Z | tsgpu.FLAG_ZSTD_DENSE, src, 4 << 20)
        first = first.copy()
        for _ in range(4):
            again, sz = ctx.transform(Z | tsgpu.FLAG_ZSTD_DENSE, src, 4 << 20)
            assert sz == fs and np.array_equal(again, first)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_frame_executor_pointer_jumping_shapes():
    # giant matches (periodic source / rounds), steps cut at the span limit, deep copy chains: every branch of the executor's
    # pointer-jumping steps, on frames libzstd wrote at three levels; plus a 4 MiB frame of each shape family
    from executor_shapes import executor_shapes
    ctx = tsgpu.Context(max_chunk_bytes=4 << 20, max_batch=2)
    try:
        shapes = executor_shapes()
        shapes["records_4MiB"] = np.resize(shapes["records"], 4 << 20)
        shapes["mixed_4MiB"] = np.resize(shapes["mixed"], 4 << 20)
        n = 0
        for name, src in shapes.items():
            for level in (1, 3, 19):
                frame = np.frombuffer(ora.zstd_compress_level(src, level), dtype=np.uint8)
                back, osz = ctx.detransform(Z, frame, [frame.size], src.size)
                assert osz == [src.size] and np.array_equal(back, src), (name, level)
                n += 1
        st = ctx.decode_path_stats()
        assert st["whole_frames"] == n and st["serial_frames"] == 0
    finally:
        ctx.close()
