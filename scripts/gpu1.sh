set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "aes" -x 2>&1 | tail -15
python bench.py --workload aes --steps 5 --warmup 3 > gpurun_out/bench_aes_r1.json 2> gpurun_out/bench_aes_r1.err; tail -c 3000 gpurun_out/bench_aes_r1.json; tail -5 gpurun_out/bench_aes_r1.err
python bench.py --workload aes --corpus R --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_aes_R_r1.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_aes_r1.csv python bench.py --workload aes --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --segment-mib 256 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gcm_main -s 1 -c 1 -o gpurun_out/prof_gcm_main_r1 -f python bench.py --workload aes --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --segment-mib 256 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "aes_bit_exact_ragged or aes_empty" -x 2>&1 | tail -8
