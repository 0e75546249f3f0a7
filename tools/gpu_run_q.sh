#!/bin/bash
# ncu capture of the search kernel on the chain instance (digitinvader9)
python tools/solve_once.py digitinvader9 > gpurun_out/plain_di9.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 4 -o gpurun_out/prof_search_di9_r02p python tools/solve_once.py digitinvader9 > gpurun_out/ncu_di9.log 2>&1
tail -3 gpurun_out/ncu_di9.log
