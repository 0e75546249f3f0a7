set -x
CMD="python bench.py --steps 3 --warmup 3 --no-also --no-cold --no-cpu-baseline"
$CMD > gpurun_out/plain_prof.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launch_r02.log 2>&1
tail -2 gpurun_out/ncu_launch_r02.log
$CMD > gpurun_out/plain_prof2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -o gpurun_out/prof_search_b6_r02 $CMD > gpurun_out/ncu_full_r02.log 2>&1
tail -3 gpurun_out/ncu_full_r02.log
CMD2="python tools/wave_trace.py partialorder_16 0"
$CMD2 > gpurun_out/plain_prof3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:expand_quad_kernel -s 30 -c 1 -o gpurun_out/prof_expand_quad_po16_r02 $CMD2 > gpurun_out/ncu_full2_r02.log 2>&1
tail -3 gpurun_out/ncu_full2_r02.log
ls -la gpurun_out/*.ncu-rep | tail -3
