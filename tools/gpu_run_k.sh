for w in 192 64 32 16 8; do
  for n in juggling_b6_f6_nosym juggling_b5_f6 digitinvader5 digitinvader9 partialorder_12; do
    STCSP_SCALAR_WALK=$w python tools/wave_trace.py $n 0 > gpurun_out/t.txt 2>&1; echo "walk $w: $(tail -1 gpurun_out/t.txt)"
  done
done
