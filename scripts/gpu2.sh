set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_zstdaes_K_r1.json 2> gpurun_out/bench_zstdaes_K_r1.err; tail -c 2500 gpurun_out/bench_zstdaes_K_r1.json; tail -5 gpurun_out/bench_zstdaes_K_r1.err
python bench.py --corpus R --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zstdaes_R_r1.json 2>/dev/null; tail -c 1500 gpurun_out/bench_zstdaes_R_r1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_zstdaes_r1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --segment-mib 256 > gpurun_out/ncu_bench2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_blocks -s 1 -c 1 -o gpurun_out/prof_zstd_enc_blocks_r1 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --segment-mib 256 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "grid and 5123" -x 2>&1 | tail -8
