"""World-size-2 gloo test of the N>1 plumbing bench.py uses: segments are sharded across ranks with no data-path
collective; the only communication is the barrier and the MAX reduction of the per-rank time, and rank 0 prints the
aggregate.  Runs on CPU (the per-rank "transform" is the oracle here: this checks the sharding/aggregation logic,
not the kernels)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys, time
    sys.path.insert(0, %r)
    import numpy as np
    import torch, torch.distributed as dist
    from oracle import oracle as ora
    from tsgpu import corpus
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cs, n = 1 << 16, 5 * (1 << 16) + 123
    src = corpus.gen_segment("K", rank, n, cs)              # every rank owns a different segment (weak scaling)
    key, aad, ivs = corpus.fixed_key_material(6)
    dist.barrier()
    t0 = time.perf_counter()
    out, sizes = ora.transform_segment(3, src, cs, key, aad, ivs)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)                 # max over ranks, as bench.py does
    back, _ = ora.detransform_chunks(3, out, sizes, n, key, aad)
    ok = torch.tensor([1 if np.array_equal(back, src) else 0])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    tot = torch.tensor([float(n)], dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"world": world, "ok": int(ok.item()), "bytes": tot.item(), "value": tot.item() / dt.item()}))
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_sharding_and_aggregation(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["ok"] == 1 and d["bytes"] == 2 * (5 * (1 << 16) + 123) and d["value"] > 0
