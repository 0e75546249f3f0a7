set -x
python tools/cold_trace.py juggling_b6_f6_nosym partialorder_14 digitinvader9 > gpurun_out/cold_trace_b.log 2>&1
grep -E "solve [0-9]|wall:" gpurun_out/cold_trace_b.log
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r02b.log 2>&1; tail -5 gpurun_out/pytest_gpu_r02b.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -5 gpurun_out/bench_r02b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02b.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','cold')})
print(d['e2e'])
for a in d.get('also',[]): print(a['workload'], a.get('device_ms'), a.get('e2e_ms'), a.get('first_solve_e2e_ms'), a.get('parity',{}).get('sha256_ok'), a.get('search_nodes'), a.get('waves'))
PY
