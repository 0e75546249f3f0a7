set -x
( time timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -x -q -k "not semantic and not partialorder_14 and not partialorder_13 and not partialorder_12 and not digitinvader9 and not digitinvader8 and not digitinvader7" ) > gpurun_out/pytest_fast_r02j.log 2>&1; tail -4 gpurun_out/pytest_fast_r02j.log
for n in juggling_b4_f4 juggling_b6_f6_nosym juggling_b5_f6 digitinvader5 digitinvader9 partialorder_14 juggling_b8_f8_nosym partialorder_16; do
  python tools/wave_trace.py $n 0 > gpurun_out/t.txt 2>&1; tail -1 gpurun_out/t.txt
done
python tools/wave_trace.py digitinvader9 0 lookahead=2 > gpurun_out/t.txt 2>&1; tail -1 gpurun_out/t.txt
python tools/wave_trace.py juggling_b4_f4 3 > gpurun_out/trace_b4f4_j.txt 2>&1
python tools/wave_trace.py juggling_b6_f6_nosym 3 > gpurun_out/trace_b6_j.txt 2>&1
grep -v block0 gpurun_out/trace_b6_j.txt | tail -9
grep block0 gpurun_out/trace_b6_j.txt | head -8
python tools/cold_trace.py juggling_b6_f6_nosym > gpurun_out/cold_j.txt 2>&1; grep -v "block0" gpurun_out/cold_j.txt | tail -8
