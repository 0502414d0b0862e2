timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_blocks -s 2 -c 1 -o gpurun_out/prof_zdec_cur -f python tests/perf/bench_detransform.py 64 > gpurun_out/ncu_zdec_cur.log 2>&1
tail -2 gpurun_out/ncu_zdec_cur.log
