# ncu captures of the round-2 kernels (one GPU; numbers printed under ncu are not bench values)
set -x
R=${1:-r02b}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_regions -s 1 -c 1 -o gpurun_out/prof_${R}_enc_regions -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify --segment-mib 256 > gpurun_out/ncu_${R}_enc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_frame_exec -s 1 -c 1 -o gpurun_out/prof_${R}_dec_frame_exec -f python bench.py --direction fetch --frames libzstd --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --segment-mib 64 > gpurun_out/ncu_${R}_decf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_regions -s 1 -c 1 -o gpurun_out/prof_${R}_dec_regions -f python tests/perf/bench_detransform.py 64 > gpurun_out/ncu_${R}_decr.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_par_entropy -s 3 -c 1 -o gpurun_out/prof_${R}_dec_entropy -f python tests/perf/bench_detransform.py 64 > gpurun_out/ncu_${R}_dece.log 2>&1
ls -la gpurun_out | tail -6
