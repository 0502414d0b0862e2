# end-of-round evidence: full GPU suite, driver-shaped bench lines for every BASELINE config, launch list, ncu captures at the bench configuration
set -x
R=${1:-r02}
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_config3_speed.json 2> gpurun_out/${R}_bench.err; tail -c 600 gpurun_out/${R}_bench_config3_speed.json
python bench.py --zstd-mode dense --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_config3_dense.json 2>> gpurun_out/${R}_bench.err
python bench.py --corpus R --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_config3_R.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 0 --steps 3 --warmup 1 > gpurun_out/${R}_bench_config0.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 1 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_config1.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_config2.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 4 --steps 16 --warmup 3 > gpurun_out/${R}_bench_config4_own.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 4 --frames libzstd --steps 16 --warmup 3 > gpurun_out/${R}_bench_config4_libzstd.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 4 --zstd-mode dense --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_config4_own_dense.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 4 --window-mib 1024 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_bulk_fetch_own.json 2>> gpurun_out/${R}_bench.err
python bench.py --config 4 --window-mib 1024 --zstd-mode dense --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_bulk_fetch_own_dense.json 2>> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err; tail -c 500 gpurun_out/${R}_bench_reference.json
python tests/perf/bench_detransform.py 256 > gpurun_out/${R}_detransform.json 2>> gpurun_out/${R}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_blocks -s 1 -c 1 -o gpurun_out/prof_${R}_zstd_enc_blocks -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_enc_regions -s 1 -c 1 -o gpurun_out/prof_${R}_zstd_enc_regions -f python bench.py --zstd-mode dense --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:gcm_main -s 1 -c 1 -o gpurun_out/prof_${R}_gcm_main -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
for K in zstd_dec_blk_literals zstd_dec_blk_sequences zstd_dec_indep_execute; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_${R}_$K -f python tests/perf/bench_detransform.py 256 > /dev/null 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_regions -s 1 -c 1 -o gpurun_out/prof_${R}_zstd_dec_regions -f python bench.py --direction fetch --frames own --zstd-mode dense --segment-mib 256 --window-mib 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_par_entropy -s 0 -c 1 -o gpurun_out/prof_${R}_zstd_dec_entropy -f python bench.py --direction fetch --frames libzstd --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --segment-mib 64 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_dec_par_execute -s 0 -c 1 -o gpurun_out/prof_${R}_zstd_dec_frame_exec -f python bench.py --direction fetch --frames libzstd --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --segment-mib 64 > /dev/null 2>&1
ls -la gpurun_out | grep ${R} | tail -30; tail -5 gpurun_out/${R}_bench.err
# compute-sanitizer over the smoke invocation (both compressors, every decode path, AES both ways): logs go under profiles/
timeout 420 compute-sanitizer --tool memcheck --log-file gpurun_out/${R}_sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_memcheck.out 2>&1; echo "memcheck exit $?" >> gpurun_out/${R}_sanitizer_memcheck.out
timeout 420 compute-sanitizer --tool racecheck --log-file gpurun_out/${R}_sanitizer_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_racecheck.out 2>&1; echo "racecheck exit $?" >> gpurun_out/${R}_sanitizer_racecheck.out
tail -3 gpurun_out/${R}_sanitizer_memcheck.log gpurun_out/${R}_sanitizer_racecheck.log
