#!/bin/bash
# round 2, session 5: timelines of the chain instance (digitinvader9) and the mid-wide one (partialorder_14)
python tools/wave_trace.py digitinvader9 3 > gpurun_out/trace_di9_p.txt 2>&1
python tools/wave_trace.py digitinvader5 3 > gpurun_out/trace_di5_p.txt 2>&1
python tools/wave_trace.py partialorder_14 3 > gpurun_out/trace_po14_p.txt 2>&1
tail -2 gpurun_out/trace_di9_p.txt
