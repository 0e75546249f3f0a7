python bench.py --steps 20 --warmup 5 --no-also --no-cpu-baseline > gpurun_out/bench_cold.json 2> gpurun_out/bench_cold.err; tail -3 gpurun_out/bench_cold.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_cold.json') if l.startswith('{')][-1])
print('cold', json.dumps(d.get('cold'), indent=1)); print('e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['ms_per_step'])
PY
python tools/cold_trace.py juggling_b6_f6_nosym 2>&1 | grep -v block0 | tail -12
