set -x
R=${1:-r02h}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -q -x -k "zstd or full_pipeline or grid or ranged or paths or identical or libzstd" 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/${R}_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('K speed: value %.1f GiB/s e2e %.1f ratio %.3f' % (d['value'], d['e2e']['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()}, d['verified'])"
for F in libzstd own; do
python bench.py --direction fetch --frames $F --steps 12 --warmup 3 --no-cpu-baseline 2>>gpurun_out/${R}_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fetch $F: value %.2f GiB/s (%.2f ms/window) e2e %.2f ms/window' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_window']), {k: round(v['ms'], 3) for k, v in d['kernels_ms_per_step'].items()}, d['verified'])"
done
python tests/perf/bench_detransform.py 256 2>>gpurun_out/${R}_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('own_frames_fast_path', 'libzstd_frames_general_path'): print(k, round(d[k]['GiB_per_s'], 1), 'GiB/s', d[k]['kernels_ms'], d[k]['bit_exact'])"
tail -3 gpurun_out/${R}_bench.err
