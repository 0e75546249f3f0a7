./tools/micro/alloc_cost > gpurun_out/alloc_cost.log 2>&1; cat gpurun_out/alloc_cost.log
python tools/wave_trace.py juggling_b6_f6_nosym 3 > gpurun_out/trace_b6_multi.txt 2>&1; tail -12 gpurun_out/trace_b6_multi.txt | cut -c1-600
python tools/wave_trace.py juggling_b6_f6_nosym 3 single_branch=1 > gpurun_out/trace_b6_single.txt 2>&1; tail -10 gpurun_out/trace_b6_single.txt | cut -c1-200
python tools/wave_trace.py partialorder_18 1 > gpurun_out/trace_po18.txt 2>&1; tail -70 gpurun_out/trace_po18.txt | cut -c1-250
