#!/bin/bash
# Times every ab/libtsgpu_*.so with the same bench line (device-resident value, kernel breakdown).  Args are passed to bench.py.
D=${AB_DIR:-ab}; for lib in $D/libtsgpu_*.so; do
  TSGPU_LIB=$PWD/$lib python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'value %.1f ratio %.3f' % (d['value'], d['compression_ratio'] or 0), {k: round(v['ms'], 2) for k, v in d['kernels_ms_per_step'].items()})"
done
