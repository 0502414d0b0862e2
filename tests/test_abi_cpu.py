"""CPU-side checks of the product library: it loads, exports every symbol include/tsgpu.h declares, the
host-side index plumbing matches the reference's golden vectors, and it refuses to work without a GPU."""
import os
import re
import subprocess

import pytest

import tsgpu
from tsgpu import binding
from oracle import oracle as ora

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENCODED_CHUNKS = "KLUv/SAPeQAAAAAAAwAAAAoBAAoAAAAe"


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-s", "-C", ROOT, "tiered-storage-for-apache-kafka_b200/libtsgpu.so"])
    return binding.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "tsgpu.h")).read()
    declared = set(re.findall(r"\b(tsgpu_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("tsgpu_ctx")
    assert declared == set(binding.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tsgpu.TsgpuError) as e:
        tsgpu.Context(max_chunk_bytes=1 << 20)
    assert e.value.code == binding.E_NODEVICE
    assert "no CPU fallback" in str(e.value)


def test_host_codec_matches_reference_vectors(lib):
    assert binding.chunk_sizes_encode([10, 20, 30]) == bytes.fromhex("000000030000000a01000a0000001e")
    assert binding.transformed_chunks_serialize([10, 20, 30]) == ENCODED_CHUNKS
    assert binding.chunk_index_json(100, 250, 110, 30) == ora.ChunkIndex.fixed(100, 250, 110, 30).to_json()
    assert binding.chunk_index_json(100, 250, None, sizes=[10, 20, 30]) == \
        ora.ChunkIndex.variable(100, 250, [10, 20, 30]).to_json()
    for vals in ([], [213], [2**31 - 1], [0, 1000, 2, 44002, 369], [1, 0xFFFFFF + 10, 0xFFFFFF + 20, 2**31 - 1]):
        enc = binding.chunk_sizes_encode(vals)
        assert enc == ora.codec_encode(vals)
        assert binding.chunk_sizes_decode(enc) == vals
    with pytest.raises(tsgpu.TsgpuError, match="Values cannot be negative"):
        binding.chunk_sizes_encode([1, -2, 3])
    # our serializer frames the codec bytes as a raw zstd block: the reference-side reader must accept it
    big = [4194332 - (i * 7919) % 100000 for i in range(256)]
    assert ora.transformed_chunks_deserialize(binding.transformed_chunks_serialize(big)) == big
