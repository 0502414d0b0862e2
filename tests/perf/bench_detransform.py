"""Detransform (fetch side) throughput probe: device-resident decrypt+decompress of (a) this library's frames
(per-block fast path), (b) libzstd level-3 frames (general path), and (c) the config-5 ranged fetch
(16 MiB window = 4 chunks) through the host C-ABI."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import tsgpu
from tsgpu import corpus
from oracle import oracle as ora
MIB = 1 << 20
seg_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seg, cs = seg_mib * MIB, 4 * MIB
nch = seg // cs
src = corpus.gen_segment('K', 0, seg, cs)
key, aad, ivs = corpus.fixed_key_material(nch)
flags = 3
ctx = tsgpu.Context(max_chunk_bytes=cs, max_batch=nch)
dev = torch.device('cuda', 0)
stride = ctx.slot_stride(flags, cs)
res = {}
def to_slots(obj, sizes):
    slots = np.zeros(nch * stride, dtype=np.uint8)
    pos = 0
    for i, s in enumerate(sizes):
        slots[i * stride + 4: i * stride + 4 + s] = obj[pos:pos + s]; pos += s
    return torch.from_numpy(slots).to(dev)
hctx = tsgpu.Context(max_chunk_bytes=cs, max_batch=16)
mine, msz = hctx.transform(flags, src, cs, key, aad, ivs)
ref, rsz = ora.transform_segment(flags, src[: 64 * MIB], cs, key, aad, ivs[: 12 * 16])
for name, obj, sizes, n in (('own_frames_fast_path', mine, msz, nch), ('libzstd_frames_general_path', ref, rsz, 16)):
    d_slots = to_slots(obj, sizes + [0] * (nch - len(sizes)))
    d_sizes = torch.tensor(sizes + [0] * (nch - len(sizes)), dtype=torch.int32, device=dev)
    d_dst = torch.zeros(n * cs, dtype=torch.uint8, device=dev)
    d_osz = torch.zeros(nch, dtype=torch.int32, device=dev); d_st = torch.zeros(nch, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def step():
        ctx.detransform_device(flags, d_slots.data_ptr(), stride, d_sizes.data_ptr(), n, cs, key, aad, d_dst.data_ptr(), d_osz.data_ptr(), d_st.data_ptr(), st)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ok = bool(np.array_equal(d_dst.cpu().numpy(), src[: n * cs])) and int(d_st.sum().item()) == 0
    ctx.profile_enable(True); step(); rep = ctx.profile_report(); ctx.profile_enable(False)
    res[name] = {'GiB_per_s': n * cs / 2**30 / (ms / 1000), 'ms': ms, 'chunks': n, 'bit_exact': ok, 'kernels_ms': {k: round(v['ms'], 3) for k, v in rep.items()}}
# config 5: ranged fetch of 16 MiB through the host API (H2D of the 4 transformed chunks, kernels, D2H of 16 MiB)
pos = np.concatenate([[0], np.cumsum(msz)])
lat = []
for first in (0, nch // 2, nch - 4):
    part = mine[pos[first]: pos[first + 4]]
    for _ in range(2): hctx.detransform(flags, part, msz[first:first + 4], 4 * cs, key, aad)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, osz = hctx.detransform(flags, part, msz[first:first + 4], 4 * cs, key, aad)
    lat.append(time.perf_counter() - t0)
    assert np.array_equal(out, src[first * cs:(first + 4) * cs])
res['ranged_fetch_16MiB_host_api'] = {'ms': [round(1000 * x, 3) for x in lat], 'GiB_per_s': 16 / 1024 / (sum(lat) / len(lat))}
t0 = time.perf_counter(); ora.detransform_chunks(flags, mine[pos[0]:pos[4]], msz[:4], 4 * cs, key, aad); res['ranged_fetch_16MiB_cpu_oracle_ms'] = round(1000 * (time.perf_counter() - t0), 2)
print(json.dumps(res))
