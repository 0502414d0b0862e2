# quick perf probe: correctness subset + kernel breakdown (no e2e / cpu arms)
set -x
TSGPU_TEST_EXPERIMENTAL=1 python -m pytest tests/test_zz_gpu_experimental.py -m gpu -q 2>&1 | tail -3     # off-by-default variants
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "zstd or full_pipeline or grid or ranged" 2>&1 | tail -4
python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('K value %.1f GiB/s  ratio %.3f' % (d['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()})"
python bench.py --corpus R --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('R value %.1f GiB/s  ratio %.3f' % (d['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()})"
TSGPU_ENC_SPLIT=1 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('K (two-launch compressor) value %.1f GiB/s  ratio %.3f' % (d['value'], d['compression_ratio']), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()})"
python tests/perf/bench_detransform.py 64 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fetch: own frames %.1f GiB/s, libzstd frames %.1f ms' % (d['own_frames_fast_path']['GiB_per_s'], d['libzstd_frames_general_path']['ms']))"
TSGPU_DEC_PARALLEL=1 python tests/perf/bench_detransform.py 64 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fetch (parallel general path): libzstd frames %.1f ms' % d['libzstd_frames_general_path']['ms'], d['libzstd_frames_general_path']['kernels_ms'])"
