set -x
R=${1:-r02e}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_regions -s 1 -c 1 -o gpurun_out/prof_${R}_enc_K -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify --segment-mib 256 > gpurun_out/ncu_${R}_encK.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_regions -s 1 -c 1 -o gpurun_out/prof_${R}_enc_R -f python bench.py --corpus R --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify --segment-mib 256 > gpurun_out/ncu_${R}_encR.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_regions -s 1 -c 1 -o gpurun_out/prof_${R}_dec_regions -f python tests/perf/bench_detransform.py 64 > gpurun_out/ncu_${R}_decr.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_par_execute -s 3 -c 1 -o gpurun_out/prof_${R}_dec_frame -f python tests/perf/bench_detransform.py 64 > gpurun_out/ncu_${R}_decf.log 2>&1
ls -la gpurun_out/*${R}*
