"""ctypes binding of the CPU oracle (oracle/libtsoracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
The product package never does.  See oracle/tsoracle.h for what each call restates (reference file:line).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtsoracle.so")

FLAG_ZSTD = 1
FLAG_AES = 2
IV_SIZE = 12
TAG_SIZE = 16
E_ARG, E_STATE, E_AUTH, E_CORRUPT, E_SHORT = -1, -2, -3, -4, -5


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class Chunk(C.Structure):
    """core/M/Chunk.java:21-36"""
    _fields_ = [("id", C.c_int32), ("original_position", C.c_int32), ("original_size", C.c_int32),
                ("transformed_position", C.c_int32), ("transformed_size", C.c_int32)]

    def tuple(self):
        return (self.id, self.original_position, self.original_size, self.transformed_position,
                self.transformed_size)


class FetchPiece(C.Structure):
    _fields_ = [("chunk_id", C.c_int32), ("skip", C.c_int32), ("take", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p = C.c_void_p
        L.ora_last_error.restype = C.c_char_p
        L.ora_zstd_version.restype = C.c_char_p
        L.ora_zstd_bound.restype = C.c_size_t
        L.ora_zstd_bound.argtypes = [C.c_size_t]
        for f in (L.ora_zstd_compress_chunk, L.ora_zstd_decompress_chunk):
            f.restype = C.c_int64
            f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t]
        L.ora_zstd_compress_level.restype = C.c_int64
        L.ora_zstd_compress_level.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int]
        L.ora_zstd_compress_checksum.restype = C.c_int64
        L.ora_zstd_compress_checksum.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_int]
        L.ora_zstd_content_size.restype = C.c_int64
        L.ora_zstd_content_size.argtypes = [u8p, C.c_size_t]
        L.ora_aesgcm_encrypt_chunk.argtypes = [u8p, u8p, u8p, C.c_size_t, u8p, C.c_size_t, u8p]
        L.ora_aesgcm_decrypt_chunk.argtypes = [u8p, u8p, C.c_size_t, u8p, C.c_size_t, u8p]
        L.ora_aesgcm_plain_encrypt.argtypes = [u8p, u8p, u8p, C.c_size_t, u8p, C.c_size_t, u8p, u8p]
        L.ora_aes256_encrypt_block.argtypes = [u8p, u8p, u8p]
        L.ora_transform_bound.restype = C.c_uint64
        L.ora_transform_bound.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32]
        L.ora_transform_segment.argtypes = [C.c_uint32, u8p, C.c_uint64, C.c_uint32, u8p, u8p, C.c_uint32, u8p,
                                            u8p, C.c_uint64, u8p, u8p]
        L.ora_detransform_chunks.argtypes = [C.c_uint32, u8p, u8p, C.c_uint32, u8p, u8p, C.c_uint32, u8p,
                                             C.c_uint64, u8p]
        L.ora_builder_new.restype = C.c_void_p
        L.ora_builder_new.argtypes = [C.c_int32] * 3
        L.ora_builder_add_chunk.argtypes = [C.c_void_p, C.c_int32]
        L.ora_builder_finish.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.ora_builder_free.argtypes = [C.c_void_p]
        L.ora_index_new_fixed.argtypes = [C.c_int32] * 4 + [C.POINTER(C.c_void_p)]
        L.ora_index_new_variable.argtypes = [C.c_int32, C.c_int32, u8p, C.c_int32, C.POINTER(C.c_void_p)]
        L.ora_index_free.argtypes = [C.c_void_p]
        L.ora_index_materialized_count.argtypes = [C.c_void_p]
        L.ora_index_chunks.argtypes = [C.c_void_p, C.POINTER(Chunk)]
        L.ora_index_find.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Chunk)]
        L.ora_index_chunks_for_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Chunk), C.c_int32]
        L.ora_index_to_json.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.ora_codec_encode.restype = C.c_int64
        L.ora_codec_encode.argtypes = [u8p, C.c_int32, u8p, C.c_size_t]
        L.ora_codec_decode.argtypes = [u8p, C.c_size_t, u8p, C.c_int32]
        L.ora_transformed_chunks_serialize.restype = C.c_int64
        L.ora_transformed_chunks_serialize.argtypes = [u8p, C.c_int32, C.c_char_p, C.c_size_t]
        L.ora_transformed_chunks_deserialize.argtypes = [C.c_char_p, u8p, C.c_int32]
        L.ora_fetch_plan.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(FetchPiece), C.c_int32]
        _lib = L
    return _lib


def _err(code):
    raise OracleError(code, lib().ora_last_error().decode())


def _p(a):
    return a.ctypes.data if a is not None and a.size else None


def _u8(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8)


# ---------------------------------------------------------------- per-chunk primitives
def zstd_compress_chunk(data):
    d = _u8(data)
    out = np.empty(lib().ora_zstd_bound(d.size) + 64, dtype=np.uint8)
    r = lib().ora_zstd_compress_chunk(_p(d), d.size, _p(out), out.size)
    if r < 0:
        _err(r)
    return out[:r].tobytes()


def zstd_compress_level(data, level):
    d = _u8(data)
    out = np.empty(lib().ora_zstd_bound(d.size) + 64, dtype=np.uint8)
    r = lib().ora_zstd_compress_level(_p(d), d.size, _p(out), out.size, level)
    if r < 0:
        _err(r)
    return out[:r].tobytes()


def zstd_compress_checksum(data, level=3):
    """a frame with Content_Checksum (the reference never writes one; its reader verifies it when present)"""
    d = _u8(data)
    out = np.empty(lib().ora_zstd_bound(d.size) + 64, dtype=np.uint8)
    r = lib().ora_zstd_compress_checksum(_p(d), d.size, _p(out), out.size, level)
    if r < 0:
        _err(r)
    return out[:r].tobytes()


def zstd_content_size(frame):
    f = _u8(frame)
    r = lib().ora_zstd_content_size(_p(f), f.size)
    if r < 0:
        _err(r)
    return r


def zstd_decompress_chunk(frame):
    f = _u8(frame)
    n = zstd_content_size(f)
    out = np.empty(max(n, 1), dtype=np.uint8)
    r = lib().ora_zstd_decompress_chunk(_p(f), f.size, _p(out), n)
    if r < 0:
        _err(r)
    return out[:r].tobytes()


def aesgcm_encrypt_chunk(key, iv, aad, pt):
    k, i, a, p = _u8(key), _u8(iv), _u8(aad), _u8(pt)
    out = np.empty(p.size + 28, dtype=np.uint8)
    rc = lib().ora_aesgcm_encrypt_chunk(_p(k), _p(i), _p(a), a.size, _p(p), p.size, _p(out))
    if rc:
        _err(rc)
    return out.tobytes()


def aesgcm_decrypt_chunk(key, aad, chunk):
    k, a, c = _u8(key), _u8(aad), _u8(chunk)
    out = np.empty(max(c.size - 28, 1), dtype=np.uint8)
    rc = lib().ora_aesgcm_decrypt_chunk(_p(k), _p(a), a.size, _p(c), c.size, _p(out))
    if rc:
        _err(rc)
    return out[:c.size - 28].tobytes()


def aesgcm_plain_encrypt(key, iv, aad, pt):
    k, i, a, p = _u8(key), _u8(iv), _u8(aad), _u8(pt)
    ct = np.empty(max(p.size, 1), dtype=np.uint8)
    tag = np.empty(16, dtype=np.uint8)
    lib().ora_aesgcm_plain_encrypt(_p(k), _p(i), _p(a), a.size, _p(p), p.size, _p(ct), _p(tag))
    return ct[:p.size].tobytes(), tag.tobytes()


def aes256_encrypt_block(key, block):
    k, b = _u8(key), _u8(block)
    out = np.empty(16, dtype=np.uint8)
    lib().ora_aes256_encrypt_block(_p(k), _p(b), _p(out))
    return out.tobytes()


# ---------------------------------------------------------------- segment chains
def transform_bound(flags, src_len, chunk_size):
    return lib().ora_transform_bound(flags, src_len, chunk_size)


def transform_segment(flags, src, chunk_size, key=None, aad=b"", ivs=None):
    """Returns (transformed bytes as np.uint8 array, list of transformed chunk sizes)."""
    s = _u8(src)
    cs = chunk_size if chunk_size else max(s.size, 1)
    n_max = (s.size + cs - 1) // cs + 1
    dst = np.empty(transform_bound(flags, s.size, chunk_size) + 64, dtype=np.uint8)
    sizes = np.zeros(n_max, dtype=np.uint32)
    n = C.c_uint32(0)
    k = _u8(key) if key is not None else np.zeros(32, np.uint8)
    a = _u8(aad)
    iv = _u8(ivs) if ivs is not None else np.zeros(12 * n_max, np.uint8)
    rc = lib().ora_transform_segment(flags, _p(s), s.size, chunk_size, _p(k), _p(a), a.size, _p(iv),
                                     _p(dst), dst.size, _p(sizes), C.addressof(n))
    if rc:
        _err(rc)
    sizes = sizes[:n.value]
    return dst[:int(sizes.sum())], [int(x) for x in sizes]


def detransform_chunks(flags, src, transformed_sizes, dst_cap, key=None, aad=b""):
    s = _u8(src)
    ts = np.asarray(transformed_sizes, dtype=np.uint32)
    dst = np.empty(max(dst_cap, 1), dtype=np.uint8)
    osz = np.zeros(max(ts.size, 1), dtype=np.uint32)
    k = _u8(key) if key is not None else np.zeros(32, np.uint8)
    a = _u8(aad)
    rc = lib().ora_detransform_chunks(flags, _p(s), _p(ts), ts.size, _p(k), _p(a), a.size, _p(dst), dst_cap,
                                      _p(osz))
    if rc:
        _err(rc)
    osz = osz[:ts.size]
    return dst[:int(osz.sum())], [int(x) for x in osz]


# ---------------------------------------------------------------- ChunkIndex
class ChunkIndex:
    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_index_free(self._h)
            self._h = None

    @staticmethod
    def fixed(ocs, ofs, tcs, ftcs):
        h = C.c_void_p()
        rc = lib().ora_index_new_fixed(ocs, ofs, tcs, ftcs, C.byref(h))
        if rc:
            _err(rc)
        return ChunkIndex(h)

    @staticmethod
    def variable(ocs, ofs, sizes):
        a = np.asarray(sizes, dtype=np.int32)
        h = C.c_void_p()
        rc = lib().ora_index_new_variable(ocs, ofs, _p(a), a.size, C.byref(h))
        if rc:
            _err(rc)
        return ChunkIndex(h)

    def chunks(self):
        n = lib().ora_index_materialized_count(self._h)
        arr = (Chunk * n)()
        lib().ora_index_chunks(self._h, arr)
        return [c.tuple() for c in arr]

    def find_chunk_for_original_offset(self, offset):
        c = Chunk()
        rc = lib().ora_index_find(self._h, offset, C.byref(c))
        if rc < 0:
            _err(rc)
        return c.tuple() if rc else None

    def chunks_for_range(self, first, last):
        cap = lib().ora_index_materialized_count(self._h) + 1
        arr = (Chunk * cap)()
        n = lib().ora_index_chunks_for_range(self._h, first, last, arr, cap)
        if n < 0:
            _err(n)
        return [arr[i].tuple() for i in range(n)]

    def to_json(self):
        buf = C.create_string_buffer(1 << 20)
        n = lib().ora_index_to_json(self._h, buf, len(buf))
        if n < 0:
            _err(n)
        return buf.value.decode()

    def fetch_plan(self, first, last):
        cap = lib().ora_index_materialized_count(self._h) + 1
        arr = (FetchPiece * cap)()
        n = lib().ora_fetch_plan(self._h, first, last, arr, cap)
        if n < 0:
            _err(n)
        return [(arr[i].chunk_id, arr[i].skip, arr[i].take) for i in range(n)]


class ChunkIndexBuilder:
    """transformed_chunk_size=None -> VariableSizeChunkIndexBuilder else FixedSizeChunkIndexBuilder
    (core/M/transform/TransformFinisher.java:75-93)."""

    def __init__(self, original_chunk_size, original_file_size, transformed_chunk_size=None):
        t = -1 if transformed_chunk_size is None else transformed_chunk_size
        if transformed_chunk_size is not None and transformed_chunk_size < 0:
            raise OracleError(E_ARG, "Transformed chunk size must be non-negative, %d given" % transformed_chunk_size)
        self._h = lib().ora_builder_new(original_chunk_size, original_file_size, t)
        if not self._h:
            _err(E_ARG)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_builder_free(self._h)
            self._h = None

    def add_chunk(self, size):
        rc = lib().ora_builder_add_chunk(self._h, size)
        if rc:
            _err(rc)

    def finish(self, size):
        h = C.c_void_p()
        rc = lib().ora_builder_finish(self._h, size, C.byref(h))
        if rc:
            _err(rc)
        return ChunkIndex(h)


# ---------------------------------------------------------------- codec
def codec_encode(values):
    a = np.asarray(values, dtype=np.int32)
    out = np.empty(32 + 4 * a.size, dtype=np.uint8)
    r = lib().ora_codec_encode(_p(a), a.size, _p(out), out.size)
    if r < 0:
        _err(r)
    return out[:r].tobytes()


def codec_decode(data, cap=1 << 22):
    d = _u8(data)
    out = np.empty(cap, dtype=np.int32)
    n = lib().ora_codec_decode(_p(d), d.size, _p(out), cap)
    if n < 0:
        _err(n)
    return [int(x) for x in out[:n]]


def transformed_chunks_serialize(values):
    a = np.asarray(values, dtype=np.int32)
    buf = C.create_string_buffer(256 + 8 * a.size)
    r = lib().ora_transformed_chunks_serialize(_p(a), a.size, buf, len(buf))
    if r < 0:
        _err(r)
    return buf.value.decode()


def transformed_chunks_deserialize(s, cap=1 << 22):
    out = np.empty(cap, dtype=np.int32)
    n = lib().ora_transformed_chunks_deserialize(s.encode(), _p(out), cap)
    if n < 0:
        _err(n)
    return [int(x) for x in out[:n]]
