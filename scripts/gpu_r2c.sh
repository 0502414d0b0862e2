# round 2, quick iteration: correctness of the zstd paths + kernel times
set -x
R=${1:-r02c}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -q -x -k "zstd or full_pipeline or grid or ranged or paths or identical" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/${R}_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('K value %.1f GiB/s  ratio %.3f' % (d['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()}, d['verified'])"
python bench.py --corpus R --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('R value %.1f GiB/s  ratio %.3f' % (d['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()})"
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
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:zstd_enc_regions -s 1 -c 1 --csv --log-file gpurun_out/${R}_enc_metrics.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify --segment-mib 256 > /dev/null 2>&1; tail -7 gpurun_out/${R}_enc_metrics.csv | cut -d, -f5,12- 
tail -5 gpurun_out/${R}_bench.err
