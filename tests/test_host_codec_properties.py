"""Property tests (hypothesis) of the host-side index code behind the C-ABI against the oracle's line-by-line restatement
of core/M/manifest/index/serde/ChunkSizesBinaryCodec.java:104-202 and AbstractChunkIndex.java:52-123.  No GPU involved:
these entry points are plain host code in libtsgpu.so."""
import json

import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as ora
import tsgpu
from tsgpu import binding

INT_MAX = 2**31 - 1
sizes_lists = st.lists(st.integers(min_value=0, max_value=INT_MAX), min_size=0, max_size=300)
clustered = st.integers(min_value=0, max_value=INT_MAX - 70000).flatmap(
    lambda base: st.lists(st.integers(min_value=0, max_value=70000), min_size=1, max_size=300).map(lambda d: [base + x for x in d]))
SET = settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@pytest.fixture(scope="module", autouse=True)
def _lib():
    binding.load()


@SET
@given(st.one_of(sizes_lists, clustered))
def test_codec_bytes_equal_the_reference_layout(values):
    enc = binding.chunk_sizes_encode(values)
    assert enc == ora.codec_encode(values)
    assert binding.chunk_sizes_decode(enc) == values
    b64 = binding.transformed_chunks_serialize(values) if values else None
    if values:
        assert ora.transformed_chunks_deserialize(b64) == values            # the reference-side reader accepts our framing


@SET
@given(st.binary(min_size=0, max_size=64))
def test_codec_decode_never_crashes_on_garbage(blob):
    try:
        out = binding.chunk_sizes_decode(blob)
    except tsgpu.TsgpuError:
        return
    assert binding.chunk_sizes_encode(out) is not None


@SET
@given(st.integers(min_value=1, max_value=1 << 22), st.integers(min_value=0, max_value=1 << 26), st.data())
def test_index_json_equals_the_oracle(ocs, ofs, data):
    n = max(1, -(-ofs // ocs))
    if n > 400:
        ocs = max(ocs, ofs // 400 + 1)
        n = max(1, -(-ofs // ocs))
    if data.draw(st.booleans()):
        tcs = data.draw(st.integers(min_value=0, max_value=1 << 23))
        ftcs = data.draw(st.integers(min_value=0, max_value=1 << 23))
        assert binding.chunk_index_json(ocs, ofs, tcs, ftcs) == ora.ChunkIndex.fixed(ocs, ofs, tcs, ftcs).to_json()
    else:
        # variable index: field names / order / numbers are the reference's; the `transformedChunks` string is Base64 of a zstd
        # frame of the codec bytes, and this library frames them as a Raw block while libzstd compresses when that pays, so
        # the strings are equal only when libzstd also stores raw (as in the reference's golden vector) — what must hold
        # always is that the reference-side reader recovers the sizes
        sizes = data.draw(st.lists(st.integers(min_value=0, max_value=1 << 23), min_size=n, max_size=n))
        mine = json.loads(binding.chunk_index_json(ocs, ofs, None, sizes=sizes))
        ref = json.loads(ora.ChunkIndex.variable(ocs, ofs, sizes).to_json())
        assert list(mine) == list(ref) == ["type", "originalChunkSize", "originalFileSize", "transformedChunks"]
        assert {k: v for k, v in mine.items() if k != "transformedChunks"} == {k: v for k, v in ref.items() if k != "transformedChunks"}
        assert ora.transformed_chunks_deserialize(mine["transformedChunks"]) == sizes
