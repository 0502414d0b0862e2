"""One context, many caller threads (Kafka's copier pool + the fetch pool, SURVEY.md §8b "Threading"): calls claim work slots
instead of serialising on a context-wide lock.  On the emulator kernels take turns (it is single-
threaded by construction), so this checks the slot bookkeeping; the -m gpu variant runs the same thing for real."""
import os
import threading

import numpy as np
import pytest

import tsgpu
from tsgpu import corpus
from oracle import oracle as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT_LIB = os.path.join(ROOT, "tests", "simt", "libtsgpu_simt.so")


def _hammer(ctx, threads, base, cs, rounds):
    key, aad, ivs = corpus.fixed_key_material(64)
    errs = []

    def work(seed):
        try:
            src = corpus.gen_segment("K", seed, base + seed * 1000, cs)
            for r in range(rounds):
                flags = 3 if (seed + r) & 1 else 2
                out, sizes = ctx.transform(flags, src, cs, key, aad, ivs)
                back, _ = ora.detransform_chunks(flags, out, sizes, src.size, key, aad)
                assert np.array_equal(back, src)
                mine, _ = ctx.detransform(flags, out, sizes, src.size, key, aad)
                assert np.array_equal(mine, src)
        except Exception as e:       # noqa: BLE001 - collected and re-raised by the caller
            errs.append(repr(e))
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


def test_simt_many_threads_share_one_context():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", ROOT, "tests/simt/libtsgpu_simt.so"])
    ctx = tsgpu.Context(max_chunk_bytes=1 << 16, max_batch=2, lib_path=SIMT_LIB)
    try:
        _hammer(ctx, threads=6, base=200000, cs=1 << 16, rounds=2)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_many_threads_share_one_context():
    ctx = tsgpu.Context(max_chunk_bytes=1 << 20, max_batch=4)
    try:
        _hammer(ctx, threads=10, base=24 << 20, cs=1 << 20, rounds=3)     # ten copier threads, 24 MiB segments
    finally:
        ctx.close()
