# bench lines only (no ncu): refresh profiles/ after a kernel change
set -x
R=${1:-r01}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${R}_zstdaes_K.json 2> gpurun_out/bench_${R}.err; tail -c 600 gpurun_out/bench_${R}_zstdaes_K.json
python bench.py --corpus R --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_zstdaes_R.json 2>/dev/null
python bench.py --workload aes --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_aes_K.json 2>/dev/null
python bench.py --workload zstd --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_zstd_K.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
