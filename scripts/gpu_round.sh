# full end-of-round measurement: tests, bench lines, launch list, ncu captures of the two dominant kernels
set -x
R=${1:-r01}
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
TSGPU_TEST_EXPERIMENTAL=1 python -m pytest tests/test_zz_gpu_experimental.py -m gpu -q 2>&1 | tail -5     # off-by-default variants
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${R}_zstdaes_K.json 2> gpurun_out/bench_${R}.err; tail -c 1200 gpurun_out/bench_${R}_zstdaes_K.json
python bench.py --corpus R --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_zstdaes_R.json 2>/dev/null
python bench.py --workload aes --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_aes_K.json 2>/dev/null
python bench.py --workload zstd --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_zstd_K.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>/dev/null; tail -c 700 gpurun_out/bench_${R}_reference.json
python tests/perf/bench_detransform.py 256 > gpurun_out/detransform_${R}.json 2> gpurun_out/detransform_${R}.err; tail -c 1500 gpurun_out/detransform_${R}.json; tail -3 gpurun_out/detransform_${R}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_blocks -s 1 -c 1 -o gpurun_out/prof_${R}_zstd_enc_blocks -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
#timeout 900 ncu --set full --clock-control none --import-source on -k regex:gcm_main -s 1 -c 1 -o gpurun_out/prof_${R}_gcm_main -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
