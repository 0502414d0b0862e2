"""The in-process multi-device context on a 2-GPU lease, without importing torch (a fresh box pays up to a minute for that and a
2-GPU lease is charged twice): the body of tests/test_gpu_parity.py::test_single_process_multi_device_context plus the device
API on device 1.  Prints one JSON line; exits non-zero on a mismatch.  usage (gpurun --gpus 2): python scripts/gpu_multidev_check.py"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsgpu  # noqa: E402
from tsgpu import corpus  # noqa: E402

MIB = 1 << 20
Z, A = tsgpu.FLAG_ZSTD, tsgpu.FLAG_AES
n_dev = ctypes.c_int(0)
ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n_dev))
if n_dev.value < 2:
    print(json.dumps({"devices": n_dev.value, "skipped": "needs 2 GPUs"}))
    sys.exit(0)
t0 = time.time()
c = tsgpu.Context(max_chunk_bytes=MIB, max_batch=4, devices=[0, 1])
n, cs = 37 * MIB + 999, MIB                          # 38 chunks -> 10 batches over 2 devices x 4 slots
src = corpus.gen_segment("K", 8, n, cs)
key, aad, ivs = corpus.fixed_key_material(38)
res = {"devices": n_dev.value}
for name, flags in (("zstd_dense+aes", Z | A | tsgpu.FLAG_ZSTD_DENSE), ("aes", A)):
    got, gs = c.transform(flags, src, cs, key, aad, ivs)
    one = tsgpu.Context(max_chunk_bytes=MIB, max_batch=4, devices=[0])
    want, ws = one.transform(flags, src, cs, key, aad, ivs)
    one.close()
    back, _ = c.detransform(flags & 3, got, gs, n, key, aad)
    res[name] = {"same_bytes_as_single_device": bool(gs == ws and np.array_equal(got, want)),
                 "round_trip": bool(np.array_equal(back, src)), "transformed_bytes": int(sum(gs))}
got, gs = c.transform(Z | A, src, cs, key, aad, ivs)   # speed mode: bytes may differ between runs, the round trip may not
back, _ = c.detransform(Z | A, got, gs, n, key, aad)
res["zstd_speed+aes"] = {"round_trip": bool(np.array_equal(back, src)), "transformed_bytes": int(sum(gs))}
res["launches"] = int(c.launch_count())
res["seconds"] = round(time.time() - t0, 2)
c.close()
ok = all(all(v for k, v in d.items() if isinstance(v, bool)) for d in res.values() if isinstance(d, dict))
res["ok"] = ok
print(json.dumps(res))
sys.exit(0 if ok else 1)
