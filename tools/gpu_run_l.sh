set -x
( time timeout 1300 python -m pytest tests -m gpu -x -q --durations=40 ) > gpurun_out/pytest_gpu_r02l.log 2>&1; tail -60 gpurun_out/pytest_gpu_r02l.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err; tail -4 gpurun_out/bench_r02l.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02l.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','parity')}); print('cold', d.get('cold')); print('e2e', d['e2e']); print('cpu', {k:v for k,v in d.get('cpu_baseline',{}).items() if k!='port_sample'})
for a in d.get('also',[]): print(a['workload'], round(a.get('device_ms',0),3), round(a.get('e2e_ms',0),2), round(a.get('first_solve_e2e_ms',0),1), a.get('parity',{}).get('sha256_ok'), a.get('search_nodes'), a.get('waves'), round(a.get('parity',{}).get('host_check_s',0),1))
PY
