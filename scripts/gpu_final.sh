# after a kernel change late in the round: the headline bench line, the launch list and one full capture of the dominant kernel
set -x
R=${1:-r02b}; O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/${R}_bench_config3_speed.json 2> $O/${R}_bench.err; tail -c 400 $O/${R}_bench_config3_speed.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_blocks -s 1 -c 1 -o $O/prof_${R}_zstd_enc_blocks -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
python scripts/ncu_summary.py $O/prof_${R}_zstd_enc_blocks.ncu-rep $O/${R}_zstd_enc_blocks.ncu.json > /dev/null
ncu -i $O/prof_${R}_zstd_enc_blocks.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv 2>/dev/null
python scripts/ncu_by_line.py /tmp/src.csv 60 > $O/${R}_zstd_enc_blocks.by_line.txt 2>&1
ls -la $O | tail -8
