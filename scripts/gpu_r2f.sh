# round 2: both compressor modes, full GPU suite, e2e
set -x
R=${1:-r02f}
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 8 --warmup 3 > gpurun_out/${R}_bench_K_speed.json 2> gpurun_out/${R}_bench.err; python - <<PY
import json
d = json.loads(open('gpurun_out/${R}_bench_K_speed.json').read().strip().splitlines()[-1])
print('K speed: value %.1f GiB/s e2e %.1f ratio %.3f' % (d['value'], d['e2e']['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()}, d['verified'], d['cpu_baseline'] and d['cpu_baseline']['copier_pool'])
PY
python bench.py --zstd-mode dense --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_K_dense.json 2>> gpurun_out/${R}_bench.err; python - <<PY
import json
d = json.loads(open('gpurun_out/${R}_bench_K_dense.json').read().strip().splitlines()[-1])
print('K dense: value %.1f GiB/s e2e %.1f ratio %.3f' % (d['value'], d['e2e']['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()}, d['verified'])
PY
python bench.py --corpus R --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_R_speed.json 2>> gpurun_out/${R}_bench.err; python - <<PY
import json
d = json.loads(open('gpurun_out/${R}_bench_R_speed.json').read().strip().splitlines()[-1])
print('R speed: value %.1f GiB/s e2e %.1f' % (d['value'], d['e2e']['value']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()})
PY
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
tail -5 gpurun_out/${R}_bench.err
