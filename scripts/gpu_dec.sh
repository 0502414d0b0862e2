# fetch-side loop while tuning the decoder: parity subset, per-kernel times on own / libzstd frames, ncu captures of the block kernels
set -x
R=${1:-r02k}
python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -q -x -k "zstd or full_pipeline or grid or ranged or paths or identical or libzstd or executor" 2>&1 | tail -3
python tests/perf/bench_detransform.py 256 2>gpurun_out/${R}_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('own_frames_fast_path', 'libzstd_frames_general_path'): print(k, round(d[k]['GiB_per_s'], 1), 'GiB/s', d[k]['kernels_ms'], d[k]['bit_exact'])
print({k: v for k, v in d.items() if k.startswith('ranged')})"
for F in own libzstd; do
python bench.py --direction fetch --frames $F --steps 12 --warmup 3 --no-cpu-baseline 2>>gpurun_out/${R}_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fetch $F: value %.2f GiB/s (%.2f ms/window) e2e %.2f ms/window' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_window']), {k: round(v['ms'], 3) for k, v in d['kernels_ms_per_step'].items()}, d['verified'])"
done
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_par_execute -s 0 -c 1 -o gpurun_out/prof_${R}_dec_frame_exec -f python bench.py --direction fetch --frames libzstd --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --segment-mib 64 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_indep_execute -s 1 -c 1 -o gpurun_out/prof_${R}_dec_blk_exec -f python tests/perf/bench_detransform.py 64 > /dev/null 2>&1
fi
tail -3 gpurun_out/${R}_bench.err
