# round 2, run A: full GPU suite (incl. bench-shape tests of the device API), fetch direction, verification in bench
set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_bench_K.json 2> gpurun_out/r02a_bench_K.err; tail -c 900 gpurun_out/r02a_bench_K.json; tail -3 gpurun_out/r02a_bench_K.err
python bench.py --direction fetch --frames libzstd --steps 16 --warmup 3 > gpurun_out/r02a_fetch_libzstd.json 2> gpurun_out/r02a_fetch_libzstd.err; tail -c 1800 gpurun_out/r02a_fetch_libzstd.json; tail -3 gpurun_out/r02a_fetch_libzstd.err
python bench.py --direction fetch --frames own --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_fetch_own.json 2> gpurun_out/r02a_fetch_own.err; tail -c 1500 gpurun_out/r02a_fetch_own.json; tail -3 gpurun_out/r02a_fetch_own.err
python tests/perf/bench_detransform.py 64 > gpurun_out/r02a_detransform.json 2> gpurun_out/r02a_detransform.err; tail -c 1500 gpurun_out/r02a_detransform.json; tail -3 gpurun_out/r02a_detransform.err
python bench.py --config 0 --steps 3 --warmup 1 | cut -c1-600
