set -x
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_bridge.py -x -q -k "golden or postprocessing or bridge" ) > gpurun_out/pytest_post_r02h.log 2>&1; tail -4 gpurun_out/pytest_post_r02h.log
( time python -m pytest tests/test_gpu_fuzz.py -x -q -k "postprocessing" ) > gpurun_out/pytest_postfuzz_r02h.log 2>&1; tail -4 gpurun_out/pytest_postfuzz_r02h.log
python tools/wave_trace.py juggling_b4_f4 3 > gpurun_out/trace_b4f4_h.txt 2>&1; tail -3 gpurun_out/trace_b4f4_h.txt
python tools/wave_trace.py juggling_b6_f6_nosym 3 > gpurun_out/trace_b6_h.txt 2>&1; tail -3 gpurun_out/trace_b6_h.txt
python tools/wave_trace.py digitinvader9 0 > gpurun_out/trace_di9_h.txt 2>&1
python tools/wave_trace.py digitinvader9 0 lookahead=2 >> gpurun_out/trace_di9_h.txt 2>&1
python tools/wave_trace.py digitinvader9 0 lookahead=1 >> gpurun_out/trace_di9_h.txt 2>&1
python tools/wave_trace.py digitinvader9 0 wide_wave_nodes=-1 >> gpurun_out/trace_di9_h.txt 2>&1
python tools/wave_trace.py digitinvader9 0 single_branch=1 >> gpurun_out/trace_di9_h.txt 2>&1
python tools/wave_trace.py partialorder_14 0 lookahead=2 >> gpurun_out/trace_di9_h.txt 2>&1
python tools/wave_trace.py partialorder_14 0 >> gpurun_out/trace_di9_h.txt 2>&1
python tools/wave_trace.py juggling_b6_f6_nosym 0 lookahead=2 >> gpurun_out/trace_di9_h.txt 2>&1
cat gpurun_out/trace_di9_h.txt
python tools/cold_trace.py > gpurun_out/cold_h.txt 2>&1; grep -v "^\[stcsp\] block0" gpurun_out/cold_h.txt | tail -60
