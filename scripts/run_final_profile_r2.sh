# round-2 end-state measurement on one B200: the default bench line, the ncu launch list and one full capture of the
# dominant kernel from the same command (each ncu pass only after the command exited 0 without ncu)
set -e
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drivers > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drivers > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_ -s 3 -c 1 -f -o gpurun_out/prof_r2_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-drivers > gpurun_out/r2_ncu_full.log 2>&1
cut -c1-300 gpurun_out/r2_bench_n1_final.json
