#!/usr/bin/env python
"""bench.py — segment transform throughput (BASELINE.json metric) on N B200s, one process per GPU.

A "step" = one pass of the hot path over one synthetic segment (default: 1 GiB, 4 MiB chunks, Zstd + AES-256-GCM,
i.e. BASELINE.json configs[3], which is the configuration the metric is quoted on and fits one GPU).
  value   device-resident: the segment is already in HBM, output slots stay in HBM (CUDA events, max over ranks)
  e2e     the reference-facing C-ABI call tsgpu_transform with HOST (pinned) buffers: H2D + kernels + D2H timed
  roofline  dominant kernel: algorithmic bytes / CUDA-event duration of that kernel vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (libzstd + OpenSSL stand-in for zstd-jni + JCE, see oracle/tsoracle.h) on host cores
`--impl reference` times that CPU path with all host threads on the same config (rank 0 only).
Weak scaling: every rank transforms its own segment; value = total original bytes / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GIB = float(1 << 30)
MIB = 1 << 20
METRIC = "segment_transform_GiB_per_s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tsgpu", choices=["tsgpu", "reference"])
    ap.add_argument("--workload", default="zstd+aes", choices=["zstd+aes", "aes", "zstd"])
    ap.add_argument("--corpus", default="K", choices=["K", "R", "Z"])
    ap.add_argument("--segment-mib", type=int, default=1024)
    ap.add_argument("--chunk-mib", type=int, default=4)
    ap.add_argument("--cpu-sample-mib", type=int, default=0, help="0 = auto (aim for ~10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def flags_of(workload):
    return {"zstd+aes": 3, "aes": 2, "zstd": 1}[workload]


def config_of(args, n_gpus):
    return {
        "workload": "%d MiB segment per GPU, %d MiB chunks, %s, corpus %s (BASELINE configs[%d])" % (
            args.segment_mib, args.chunk_mib, args.workload, args.corpus,
            {"zstd+aes": 3, "aes": 2, "zstd": 1}[args.workload]),
        "segment_bytes": args.segment_mib * MIB, "chunk_bytes": args.chunk_mib * MIB,
        "transform": args.workload, "corpus": args.corpus,
        "parallelism": "segments sharded across %d GPU(s), no data-path collective" % n_gpus,
        "l2_policy": "inputs (segment >= 1 GiB) larger than the 126 MB L2; no explicit flush",
    }


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []          # (arrival time, csv line)
        self.proc = None
        self.t0 = self.t1 = None

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # samples that arrived inside the timed region; if the region was shorter than the sampling period, the samples
        # taken under the same load right before it (warm-up runs the identical step) are used and counted separately
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        rows = inside if inside else [r for (t, r) in self.rows][-5:]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_segment(args, segment_id):
    from tsgpu import corpus
    return corpus.gen_segment(args.corpus, segment_id, args.segment_mib * MIB, args.chunk_mib * MIB)


# ------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_transform_chunks(ora, flags, src, cs, key, aad, ivs, lo, hi):
    """virtual chunks [lo, hi) through the oracle's chain, one chunk at a time like the reference's pull pipeline;
    virtual chunk v is chunk v mod nch of the sample (several passes over the sample keep every worker busy)"""
    n = src.size
    nch = max(1, (n + cs - 1) // cs)
    out = 0
    for v in range(lo, hi):
        i = v % nch
        a, b = i * cs, min(n, (i + 1) * cs)
        t, sizes = ora.transform_segment(flags, src[a:b], cs, key, aad, ivs[12 * i:12 * i + 12])
        out += sizes[0]
    return out


_POOL_STATE = {}


def _pool_job(r):
    st = _POOL_STATE
    c0 = time.process_time()
    cpu_transform_chunks(st["ora"], st["flags"], st["src"], st["cs"], st["key"], st["aad"], st["ivs"], r[0], r[1])
    return time.process_time() - c0


class CpuArm:
    """The oracle's chain over the first `sample_bytes` of a segment on `threads` host cores.  threads > 1 uses
    forked worker processes kept alive across steps (per-chunk contexts and buffers are freshly allocated, as in
    the reference; separate address spaces keep page-fault handling off one mm lock, and warm-up passes populate
    the forked page tables before anything is timed).  Every worker gets at least `min_chunks_per_worker` chunks
    per step (the sample is passed over several times if needed) so that dispatch latency does not dominate."""

    def __init__(self, args, flags, src, threads, sample_bytes, min_chunks_per_worker=1):
        from oracle import oracle as ora
        from tsgpu import corpus
        self.cs = args.chunk_mib * MIB
        self.nch = max(1, min(src.size, sample_bytes) // self.cs)
        key, aad, ivs = corpus.fixed_key_material(self.nch)
        self.rounds = max(1, -(-min_chunks_per_worker * threads // self.nch)) if threads > 1 else 1
        total = self.nch * self.rounds
        per = (total + threads - 1) // threads
        self.ranges = [(k * per, min(total, (k + 1) * per)) for k in range(threads) if k * per < total]
        _POOL_STATE.update(ora=ora, flags=flags, src=src[:self.nch * self.cs], cs=self.cs, key=key, aad=aad, ivs=ivs)
        self.pool = None
        self.cpu_seconds = 0.0
        if len(self.ranges) > 1:
            import multiprocessing as mp
            self.pool = mp.get_context("fork").Pool(len(self.ranges))

    @property
    def cores(self):
        return len(self.ranges)

    def step(self):
        """one pass; returns (bytes, seconds)"""
        t0 = time.perf_counter()
        if self.pool is None:
            cpu = [_pool_job(self.ranges[0])]
        else:
            cpu = self.pool.map(_pool_job, self.ranges, chunksize=1)
        dt = time.perf_counter() - t0
        self.cpu_seconds += float(sum(cpu))
        return self.nch * self.rounds * self.cs, dt

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as ora
    flags = flags_of(args.workload)
    threads = os.cpu_count() or 1
    # a bounded sample of the same workload: the first sample_mib of the segment, all host cores
    sample_mib = args.cpu_sample_mib or min(args.segment_mib, 64 * threads)
    seg_args = argparse.Namespace(**vars(args))
    seg_args.segment_mib = sample_mib
    src = make_segment(seg_args, 0)
    arm = CpuArm(args, flags, src, threads, src.size, min_chunks_per_worker=8)
    for _ in range(max(1, args.warmup)):
        arm.step()
    arm.cpu_seconds = 0.0
    t_tot, b_tot = 0.0, 0
    for _ in range(args.steps):
        nbytes, dt = arm.step()
        t_tot += dt; b_tot += nbytes
    used = arm.cores
    busy = arm.cpu_seconds / t_tot if t_tot > 0 else 0.0      # CPU-seconds burnt per wall second = cores actually granted
    rounds = arm.rounds
    arm.close()
    val = b_tot / GIB / t_tot
    line = {
        "metric": METRIC, "value": val, "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "impl": "reference", "config": config_of(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": "GiB/s", "cores": used, "kind": "port",
                         "effective_cores": round(busy, 1),
                         "sample": "first %d MiB of the segment, %d pass(es) per step, on %d worker processes (%.1f CPU-seconds "
                                   "consumed per wall second: what the host actually granted); libzstd %s level 3 + OpenSSL "
                                   "EVP AES-256-GCM standing in for zstd-jni 1.5.6-9 + SunJCE (no JVM in the image)" % (
                                       sample_mib, rounds, used, busy, ora.lib().ora_zstd_version().decode())},
        "e2e": {"value": val, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(torch, index):
    """Pin this rank to the CPUs next to its GPU so the pinned staging buffers are first-touched on the local NUMA
    node (matters for the PCIe-bound e2e number when 8 ranks share one host).  Best effort: any failure is ignored."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        if ids:
            os.sched_setaffinity(0, ids)
    except Exception:
        pass


def main_tsgpu(args):
    import torch
    import torch.distributed as dist
    import tsgpu
    from tsgpu import corpus

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl tsgpu needs a GPU: libtsgpu has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flags = flags_of(args.workload)
    seg, cs = args.segment_mib * MIB, args.chunk_mib * MIB
    nch = seg // cs
    key, aad, ivs = corpus.fixed_key_material(nch)

    src_np = make_segment(args, rank)
    h_src = torch.empty(seg, dtype=torch.uint8).pin_memory()
    h_src.numpy()[:] = src_np
    d_src = h_src.to(dev, non_blocking=True)
    ctx = tsgpu.Context(max_chunk_bytes=cs, max_batch=nch, devices=[local])
    stride = ctx.slot_stride(flags, cs)
    d_slots = torch.empty(nch * stride, dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(nch, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        ctx.transform_device(flags, d_src.data_ptr(), seg, cs, key, aad, ivs, d_slots.data_ptr(), stride,
                             d_sizes.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events on the launching stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                    # started before warm-up so nvidia-smi is already streaming
    for _ in range(args.warmup):
        step_device()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    sampler.end()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * seg * args.steps / GIB / (ms_max / 1000.0)
    sizes = d_sizes.cpu().numpy().astype(np.int64)
    transformed_total = int(sizes.sum())

    # ---- roofline: per-kernel CUDA-event timing, separate steps so the events do not perturb `value`
    roofline, kernels = None, None
    ctx.profile_enable(True)
    for _ in range(max(1, min(args.steps, 3))):
        step_device()
    rep = ctx.profile_report()
    ctx.profile_enable(False)
    psteps = max(1, min(args.steps, 3))
    if rep:
        kernels = {k: {"launches": v["launches"] // psteps, "ms": v["ms"] / psteps} for k, v in rep.items()}
        top = max(rep.items(), key=lambda kv: kv[1]["ms"])
        name, ms_k = top[0], top[1]["ms"] / top[1]["launches"]
        # algorithmic bytes of the dominant kernel per launch (DESIGN.md "Kernels"): the bytes it must read + write
        if name.startswith("zstd_compress") or name.startswith("zstd_enc"):
            frame_total = transformed_total - (28 * nch if flags & 2 else 0)
            alg = seg + frame_total
        elif name.startswith("gcm_main"):
            payload = transformed_total - 28 * nch
            alg = 2 * payload
        else:
            alg = seg + transformed_total
        peak, how = load_peaks()
        achieved = alg / 1e9 / (ms_k / 1000.0)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(name)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                    "algorithmic_bytes_per_launch": alg, "kernel_ms": ms_k,
                    "note": "integer/LDS-bound kernels: see DESIGN.md for the ALU/shared-memory ceilings"}

    # ---- e2e: the host-buffer C-ABI call (pinned src/dst), H2D + kernels + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        ectx = tsgpu.Context(max_chunk_bytes=cs, max_batch=4, devices=[local])
        cap = int(ectx.lib.tsgpu_transform_bound(flags, seg, cs)) + 64
        h_dst = torch.empty(cap, dtype=torch.uint8).pin_memory()
        dst_np = h_dst.numpy()
        out_sizes = None
        for _ in range(max(1, args.warmup)):
            _, out_sizes = ectx.transform(flags, src_np if False else h_src.numpy(), cs, key, aad, ivs, dst=dst_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, out_sizes = ectx.transform(flags, h_src.numpy(), cs, key, aad, ivs, dst=dst_np)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_max = float(tt.item())
        e2e = {"value": world * seg * args.steps / GIB / dt_max, "unit": "GiB/s",
               "h2d_bytes_per_step": seg, "d2h_bytes_per_step": int(sum(out_sizes)),
               "timer": "host wall clock around tsgpu_transform (it synchronises internally), max over ranks"}
        ectx.close()

    # ---- cpu_baseline: the oracle on this box's host cores, rank 0, bounded sample
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as ora
        probe = CpuArm(args, flags, src_np, 1, 2 * cs)
        probe.step()
        pb, pdt = probe.step()
        probe_v = pb / GIB / pdt
        target_s = 10.0
        sample = args.cpu_sample_mib * MIB if args.cpu_sample_mib else int(min(seg, max(4 * cs, probe_v * GIB * target_s)))
        sample = (sample // cs) * cs
        arm = CpuArm(args, flags, src_np, 1, sample)
        nbytes, dt = arm.step()
        v = nbytes / GIB / dt
        cpu = {"value": v, "unit": "GiB/s", "cores": 1, "kind": "port",
               "sample": "first %d MiB of the same segment, chunk-sequential on 1 thread (the reference's per-segment "
                         "pipeline is single-threaded); libzstd %s level 3 + OpenSSL EVP AES-256-GCM standing in for "
                         "zstd-jni 1.5.6-9 + SunJCE; %.1f s" % (nbytes // MIB, ora.lib().ora_zstd_version().decode(), dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GiB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config_of(args, world),
            "compression_ratio": seg / max(1, transformed_total - (28 * nch if flags & 2 else 0)) if flags & 1 else None,
            "transformed_bytes_per_segment": transformed_total,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "kernels_ms_per_step": kernels,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(main_reference(a) if a.impl == "reference" else main_tsgpu(a))
