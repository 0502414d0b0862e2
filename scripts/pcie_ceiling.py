"""H2D / D2H ceilings with N ranks copying at the same time (torchrun; one rank per GPU, pinned buffers first-touched
after binding the rank to its GPU's NUMA node like bench.py does).  VERDICT r1 #5: the e2e number at N >= 4 is limited by
how many GPUs share a NUMA node / PCIe root, not by the kernels — this is the measurement."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bind_to_gpu_numa_node(torch, local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(rank + 1)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2.fill_(0)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return reps * n / 2**30 / float(t.item())


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d, non_blocking=True)


res = {"ranks": world,
       "h2d_GiB_per_s_per_rank": timed(lambda: d.copy_(h, non_blocking=True)),
       "d2h_GiB_per_s_per_rank": timed(lambda: h2.copy_(d, non_blocking=True)),
       "h2d_and_d2h_together_GiB_per_s_per_rank_each_way": timed(both)}
try:
    p = torch.cuda.get_device_properties(local)
    res["numa_node_of_gpu0"] = open("/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).read().strip()
except Exception:
    pass
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
