set -x
( time timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -x -q -k "not semantic and not partialorder_14 and not partialorder_13 and not partialorder_12 and not digitinvader9 and not digitinvader8 and not digitinvader7" ) > gpurun_out/pytest_fast_r02i.log 2>&1; tail -4 gpurun_out/pytest_fast_r02i.log
for n in juggling_b4_f4 juggling_b6_f6_nosym juggling_b5_f6 digitinvader5; do
  python tools/wave_trace.py $n 0 > gpurun_out/t.txt 2>&1; tail -1 gpurun_out/t.txt
  STCSP_DBG_FLAGS=1 python tools/wave_trace.py $n 0 > gpurun_out/t.txt 2>&1; echo "unfused: $(tail -1 gpurun_out/t.txt)"
done
python tools/wave_trace.py juggling_b4_f4 3 > gpurun_out/trace_b4f4_i.txt 2>&1
python tools/wave_trace.py juggling_b6_f6_nosym 3 > gpurun_out/trace_b6_i.txt 2>&1
grep -v block0 gpurun_out/trace_b6_i.txt | tail -9
for kw in "" "lookahead=2" "lookahead=1" "wide_wave_nodes=-1" "single_branch=1" "expand_mode=3" "expand_mode=1"; do
  python tools/wave_trace.py digitinvader9 0 $kw > gpurun_out/t.txt 2>&1; echo "di9 [$kw]: $(tail -1 gpurun_out/t.txt)"
done
for kw in "" "lookahead=2"; do
  python tools/wave_trace.py partialorder_14 0 $kw > gpurun_out/t.txt 2>&1; echo "po14 [$kw]: $(tail -1 gpurun_out/t.txt)"
  python tools/wave_trace.py juggling_b6_f6_nosym 0 $kw > gpurun_out/t.txt 2>&1; echo "b6 [$kw]: $(tail -1 gpurun_out/t.txt)"
done
python tools/cold_trace.py > gpurun_out/cold_i.txt 2>&1; grep -v "block0" gpurun_out/cold_i.txt | grep -v "wave [0-9]*:" | tail -40
