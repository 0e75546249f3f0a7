#!/bin/bash
# round 2, last session: launch list + full ncu capture of the search kernel (headline) and the default bench line
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-also --no-cold --no-cpu-baseline"
$CMD > gpurun_out/plain_prof_h.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02h.csv $CMD > gpurun_out/ncu_launch_r02h.log 2>&1
tail -2 gpurun_out/ncu_launch_r02h.log
$CMD > gpurun_out/plain_prof_h2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -o gpurun_out/prof_search_b6_r02h $CMD > gpurun_out/ncu_full_r02h.log 2>&1
tail -3 gpurun_out/ncu_full_r02h.log
( time python bench.py --steps 200 --warmup 5 ) > gpurun_out/bench_r02r.json 2> gpurun_out/bench_r02r.err; tail -4 gpurun_out/bench_r02r.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02r.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','parity','gpu_launches','steps')}); print('cold', d.get('cold')); print('e2e', d['e2e']); print('roofline', d['roofline'])
for a in d.get('also',[]): print(a['workload'], round(a.get('device_ms',0),3), round(a.get('e2e_ms',0),2), round(a.get('first_solve_e2e_ms',0),1), a.get('parity',{}).get('sha256_ok'), a.get('search_nodes'), a.get('waves'), round(a.get('parity',{}).get('host_check_s',0),1))
PY
