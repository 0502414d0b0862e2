"""ctypes binding of the C-ABI in include/tsgpu.h (libtsgpu.so, built in-tree by `make` / __graft_entry__.build()).

This is the same surface a JNI shim binds.  There is no CPU fallback: if the CUDA library
is missing or no device is visible, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libtsgpu.so")

FLAG_ZSTD = 1
FLAG_AES = 2
FLAG_ZSTD_DENSE = 4
IV_SIZE = 12
TAG_SIZE = 16
SLOT_HEAD = 4
OK, E_ARG, E_STATE, E_AUTH, E_CORRUPT, E_SHORT, E_NODEVICE, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4, -5, -6, -7, -8

SYMBOLS = [
    "tsgpu_last_error", "tsgpu_version", "tsgpu_create", "tsgpu_destroy", "tsgpu_host_alloc", "tsgpu_host_free",
    "tsgpu_transform_bound", "tsgpu_transform", "tsgpu_transform_chunks", "tsgpu_detransform", "tsgpu_slot_stride",
    "tsgpu_transform_device", "tsgpu_detransform_device", "tsgpu_launch_count", "tsgpu_chunk_positions",
    "tsgpu_chunk_sizes_encode", "tsgpu_chunk_sizes_decode", "tsgpu_transformed_chunks_serialize",
    "tsgpu_transformed_chunks_deserialize", "tsgpu_chunk_index_json", "tsgpu_profile_enable", "tsgpu_profile_report",
    "tsgpu_decode_path_stats", "tsgpu_transformed_chunks_serialize_ctx", "tsgpu_chunk_index_json_ctx",
]


class TsgpuError(RuntimeError):
    """Maps to the unchecked RuntimeException the reference's operators throw."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_libs = {}


def load(path=None):
    path = path or os.environ.get("TSGPU_LIB") or DEFAULT_LIB
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise TsgpuError(E_NODEVICE, "libtsgpu.so not built at %s (run `make` or __graft_entry__.build()); "
                                     "there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    L.tsgpu_last_error.restype = C.c_char_p
    L.tsgpu_version.restype = C.c_char_p
    L.tsgpu_create.argtypes = [vp, C.c_int, u32, u32, C.POINTER(vp)]
    L.tsgpu_destroy.argtypes = [vp]
    L.tsgpu_host_alloc.restype = vp
    L.tsgpu_host_alloc.argtypes = [C.c_size_t]
    L.tsgpu_host_free.argtypes = [vp]
    L.tsgpu_transform_bound.restype = u64
    L.tsgpu_transform_bound.argtypes = [u32, u64, u32]
    L.tsgpu_slot_stride.restype = u64
    L.tsgpu_slot_stride.argtypes = [u32, u32]
    L.tsgpu_transform.argtypes = [vp, u32, vp, u64, u32, vp, vp, u32, vp, vp, u64, vp, vp]
    L.tsgpu_transform_chunks.argtypes = [vp, u32, vp, vp, u32, vp, vp, u32, vp, vp, u64, vp]
    L.tsgpu_detransform.argtypes = [vp, u32, vp, u64, vp, u32, vp, vp, u32, vp, u64, vp]
    L.tsgpu_transform_device.argtypes = [vp, C.c_int, u32, vp, u64, u32, vp, vp, u32, vp, vp, u64, vp, vp]
    L.tsgpu_detransform_device.argtypes = [vp, C.c_int, u32, vp, u64, vp, u32, u32, vp, vp, u32, vp, vp, vp, vp]
    L.tsgpu_launch_count.restype = u64
    L.tsgpu_launch_count.argtypes = [vp]
    L.tsgpu_decode_path_stats.argtypes = [vp, vp]
    L.tsgpu_profile_enable.argtypes = [vp, C.c_int]
    L.tsgpu_profile_report.argtypes = [vp, C.c_char_p, C.POINTER(u32)]
    L.tsgpu_chunk_positions.argtypes = [vp, vp, u32, vp]
    L.tsgpu_chunk_sizes_encode.argtypes = [vp, u32, vp, C.POINTER(u32)]
    L.tsgpu_chunk_sizes_decode.argtypes = [vp, u32, vp, C.POINTER(u32)]
    L.tsgpu_transformed_chunks_serialize.argtypes = [vp, u32, C.c_char_p, C.POINTER(u32)]
    L.tsgpu_transformed_chunks_deserialize.argtypes = [vp, C.c_char_p, vp, C.POINTER(u32)]
    L.tsgpu_chunk_index_json.argtypes = [i32, i32, i32, i32, vp, u32, C.c_char_p, C.POINTER(u32)]
    L.tsgpu_transformed_chunks_serialize_ctx.argtypes = [vp, vp, u32, C.c_char_p, C.POINTER(u32)]
    L.tsgpu_chunk_index_json_ctx.argtypes = [vp, i32, i32, i32, i32, vp, u32, C.c_char_p, C.POINTER(u32)]
    _libs[path] = L
    return L


def _u8(b):
    if b is None:
        return np.zeros(0, np.uint8)
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(b), dtype=np.uint8)


def _p(a):
    return a.ctypes.data if a is not None and a.size else None


class Context:
    """tsgpu_ctx: one per process (JVM); thread-safe."""

    def __init__(self, max_chunk_bytes, max_batch=32, devices=None, lib_path=None):
        self.lib = load(lib_path)
        self._h = C.c_void_p()
        ids = (C.c_int * len(devices))(*devices) if devices else None
        rc = self.lib.tsgpu_create(ids, len(devices) if devices else 0, max_chunk_bytes, max_batch, C.byref(self._h))
        if rc:
            raise TsgpuError(rc, self.lib.tsgpu_last_error().decode())
        self.max_chunk_bytes = max_chunk_bytes
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.tsgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise TsgpuError(rc, self.lib.tsgpu_last_error().decode())

    @property
    def handle(self):
        return self._h

    def launch_count(self):
        return int(self.lib.tsgpu_launch_count(self._h))

    def decode_path_stats(self):
        """{'regions': n, 'region_fallback_frames': n, 'whole_frames': n, 'serial_frames': n, 'blocks': n} since the context was created"""
        v = (C.c_uint64 * 8)()
        self._check(self.lib.tsgpu_decode_path_stats(self._h, v))
        return dict(zip(("regions", "region_fallback_frames", "whole_frames", "serial_frames", "blocks"), [int(x) for x in v]))

    def profile_enable(self, on=True):
        self._check(self.lib.tsgpu_profile_enable(self._h, 1 if on else 0))

    def profile_report(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        n = C.c_uint32(len(buf))
        self._check(self.lib.tsgpu_profile_report(self._h, buf, C.byref(n)))
        return json.loads(buf.value.decode())

    # ---- host-buffer API (what the JNI shim calls) ----
    def transform(self, flags, src, chunk_size, key=None, aad=b"", ivs=None, dst=None):
        """Returns (np.uint8 view of the transformed bytes, [transformed sizes])."""
        s = _u8(src)
        cs = chunk_size if chunk_size else max(s.size, 1)
        n_max = (s.size + cs - 1) // cs + 1
        cap = int(self.lib.tsgpu_transform_bound(flags, s.size, chunk_size)) + 64
        if dst is None:
            dst = np.empty(cap, dtype=np.uint8)
        sizes = np.zeros(n_max, dtype=np.uint32)
        n = C.c_uint32(n_max)
        k, a, iv = _u8(key), _u8(aad), _u8(ivs)
        self._check(self.lib.tsgpu_transform(self._h, flags, _p(s), s.size, chunk_size, _p(k), _p(a), a.size, _p(iv),
                                             dst.ctypes.data, dst.size, sizes.ctypes.data, C.addressof(n)))
        sizes = sizes[:n.value]
        return dst[:int(sizes.sum(dtype=np.uint64))], [int(x) for x in sizes]

    def transform_chunks(self, flags, src, chunk_lens, key=None, aad=b"", ivs=None):
        """Ragged chunks (e.g. the index files of a segment): returns (transformed bytes, [transformed sizes])."""
        s = _u8(src)
        lens = np.ascontiguousarray(chunk_lens, dtype=np.uint32)
        cap = int(sum(int(self.lib.tsgpu_transform_bound(flags, int(x), int(x))) for x in lens)) + 64
        dst = np.empty(cap, dtype=np.uint8)
        sizes = np.zeros(max(lens.size, 1), dtype=np.uint32)
        k, a, iv = _u8(key), _u8(aad), _u8(ivs)
        self._check(self.lib.tsgpu_transform_chunks(self._h, flags, _p(s), _p(lens), lens.size, _p(k), _p(a), a.size, _p(iv),
                                                    dst.ctypes.data, dst.size, sizes.ctypes.data))
        sizes = sizes[:lens.size]
        return dst[:int(sizes.sum(dtype=np.uint64))], [int(x) for x in sizes]

    def detransform(self, flags, src, transformed_sizes, dst_cap, key=None, aad=b"", dst=None):
        """Returns (np.uint8 view of the original bytes, [original sizes])."""
        s = _u8(src)
        ts = np.ascontiguousarray(transformed_sizes, dtype=np.uint32)
        if dst is None:
            dst = np.empty(max(int(dst_cap), 1), dtype=np.uint8)
        osz = np.zeros(max(ts.size, 1), dtype=np.uint32)
        k, a = _u8(key), _u8(aad)
        self._check(self.lib.tsgpu_detransform(self._h, flags, _p(s), s.size, _p(ts), ts.size, _p(k), _p(a), a.size,
                                               dst.ctypes.data, int(dst_cap), osz.ctypes.data))
        osz = osz[:ts.size]
        return dst[:int(osz.sum(dtype=np.uint64))], [int(x) for x in osz]

    # ---- device-resident API (raw device pointers as ints; bench.py passes torch data_ptr()) ----
    def slot_stride(self, flags, chunk_size):
        return int(self.lib.tsgpu_slot_stride(flags, chunk_size))

    def transform_device(self, flags, d_src, src_len, chunk_size, key, aad, ivs, d_slots, slot_stride, d_sizes,
                         stream=0, device_index=0):
        k, a, iv = _u8(key), _u8(aad), _u8(ivs)
        self._check(self.lib.tsgpu_transform_device(self._h, device_index, flags, d_src, src_len, chunk_size, _p(k), _p(a),
                                                    a.size, _p(iv), d_slots, slot_stride, d_sizes, stream))

    def detransform_device(self, flags, d_slots, slot_stride, d_sizes, n_chunks, chunk_size, key, aad, d_dst,
                           d_original_sizes, d_status, stream=0, device_index=0):
        k, a = _u8(key), _u8(aad)
        self._check(self.lib.tsgpu_detransform_device(self._h, device_index, flags, d_slots, slot_stride, d_sizes, n_chunks,
                                                      chunk_size, _p(k), _p(a), a.size, d_dst, d_original_sizes, d_status,
                                                      stream))

    # ---- ChunkIndex plumbing ----
    def chunk_positions(self, sizes):
        s = np.ascontiguousarray(sizes, dtype=np.uint32)
        pos = np.zeros(s.size + 1, dtype=np.uint64)
        self._check(self.lib.tsgpu_chunk_positions(self._h, _p(s), s.size, pos.ctypes.data))
        return pos

    def transformed_chunks_deserialize(self, b64, cap=1 << 22):
        out = np.zeros(cap, dtype=np.int32)
        n = C.c_uint32(cap)
        self._check(self.lib.tsgpu_transformed_chunks_deserialize(self._h, b64.encode(), out.ctypes.data, C.byref(n)))
        return [int(x) for x in out[:n.value]]

    def transformed_chunks_serialize(self, values):
        """TransformedChunksSerializer with the codec bytes compressed (dense compressor), like the reference's manifests"""
        v = np.ascontiguousarray(values, dtype=np.int32)
        buf = C.create_string_buffer(256 + 8 * v.size)
        n = C.c_uint32(len(buf))
        self._check(self.lib.tsgpu_transformed_chunks_serialize_ctx(self._h, _p(v), v.size, buf, C.byref(n)))
        return buf.value.decode()

    def chunk_index_json(self, original_chunk_size, original_file_size, transformed_chunk_size=None,
                         final_transformed_chunk_size=0, sizes=None):
        v = np.ascontiguousarray(sizes if sizes is not None else [], dtype=np.int32)
        buf = C.create_string_buffer(1024 + 8 * v.size)
        n = C.c_uint32(len(buf))
        tcs = -1 if transformed_chunk_size is None else transformed_chunk_size
        self._check(self.lib.tsgpu_chunk_index_json_ctx(self._h, original_chunk_size, original_file_size, tcs,
                                                        final_transformed_chunk_size, _p(v), v.size, buf, C.byref(n)))
        return buf.value.decode()


def _raise(L, rc):
    raise TsgpuError(rc, L.tsgpu_last_error().decode())


def chunk_sizes_encode(values, lib_path=None):
    L = load(lib_path)
    v = np.ascontiguousarray(values, dtype=np.int32)
    out = np.zeros(32 + 4 * v.size, dtype=np.uint8)
    n = C.c_uint32(out.size)
    rc = L.tsgpu_chunk_sizes_encode(_p(v), v.size, out.ctypes.data, C.byref(n))
    if rc:
        _raise(L, rc)
    return out[:n.value].tobytes()


def chunk_sizes_decode(data, cap=1 << 22, lib_path=None):
    L = load(lib_path)
    d = _u8(data)
    out = np.zeros(cap, dtype=np.int32)
    n = C.c_uint32(cap)
    rc = L.tsgpu_chunk_sizes_decode(_p(d), d.size, out.ctypes.data, C.byref(n))
    if rc:
        _raise(L, rc)
    return [int(x) for x in out[:n.value]]


def transformed_chunks_serialize(values, lib_path=None):
    L = load(lib_path)
    v = np.ascontiguousarray(values, dtype=np.int32)
    buf = C.create_string_buffer(256 + 8 * v.size)
    n = C.c_uint32(len(buf))
    rc = L.tsgpu_transformed_chunks_serialize(_p(v), v.size, buf, C.byref(n))
    if rc:
        _raise(L, rc)
    return buf.value.decode()


def chunk_index_json(original_chunk_size, original_file_size, transformed_chunk_size=None,
                     final_transformed_chunk_size=0, sizes=None, lib_path=None):
    L = load(lib_path)
    v = np.ascontiguousarray(sizes if sizes is not None else [], dtype=np.int32)
    buf = C.create_string_buffer(1024 + 8 * v.size)
    n = C.c_uint32(len(buf))
    tcs = -1 if transformed_chunk_size is None else transformed_chunk_size
    rc = L.tsgpu_chunk_index_json(original_chunk_size, original_file_size, tcs, final_transformed_chunk_size,
                                  _p(v), v.size, buf, C.byref(n))
    if rc:
        _raise(L, rc)
    return buf.value.decode()
