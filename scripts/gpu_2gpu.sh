# 2-GPU lease: the in-process multi-device context, torchrun weak scaling (NCCL barrier / max-over-ranks), H2D/D2H ceilings with two ranks
set -x
R=${1:-r02}
nvidia-smi topo -m 2>/dev/null | head -8
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_2gpu.json 2> gpurun_out/${R}_bench_2gpu.err; tail -c 1500 gpurun_out/${R}_bench_2gpu.json; tail -3 gpurun_out/${R}_bench_2gpu.err
python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_1gpu_same_box.json 2>/dev/null; python - <<PY
import json
a = json.loads(open('gpurun_out/${R}_bench_1gpu_same_box.json').read().strip().splitlines()[-1]); b = json.loads(open('gpurun_out/${R}_bench_2gpu.json').read().strip().splitlines()[-1])
print('1 GPU value %.1f e2e %.1f | 2 GPUs value %.1f e2e %.1f | e2e scaling %.3f' % (a['value'], a['e2e']['value'], b['value'], b['e2e']['value'], b['e2e']['value'] / (2 * a['e2e']['value'])))
PY
python scripts/pcie_ceiling.py > gpurun_out/${R}_pcie_1rank.json 2>/dev/null; cat gpurun_out/${R}_pcie_1rank.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/pcie_ceiling.py > gpurun_out/${R}_pcie_2ranks.json 2>/dev/null; cat gpurun_out/${R}_pcie_2ranks.json
