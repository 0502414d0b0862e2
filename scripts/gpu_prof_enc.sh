set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_blocks -s 1 -c 1 -o gpurun_out/prof_zenc_cur -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --segment-mib 128 > gpurun_out/ncu_zenc_cur.log 2>&1
tail -2 gpurun_out/ncu_zenc_cur.log
