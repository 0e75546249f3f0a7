set -x
( time timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q -k "not semantic and not partialorder_14 and not partialorder_13 and not partialorder_12 and not digitinvader9 and not digitinvader8 and not digitinvader7 and not cli" ) > gpurun_out/pytest_fast_r02m.log 2>&1; tail -4 gpurun_out/pytest_fast_r02m.log
for n in juggling_b4_f4 juggling_b6_f6_nosym juggling_b5_f6 digitinvader3; do
  python tools/wave_trace.py $n 0 > gpurun_out/t.txt 2>&1; echo "push: $(tail -1 gpurun_out/t.txt)"
  STCSP_NO_PUSH=1 python tools/wave_trace.py $n 0 > gpurun_out/t.txt 2>&1; echo "nopush: $(tail -1 gpurun_out/t.txt)"
done
python bench.py --steps 50 --warmup 5 --no-also --no-cold --no-cpu-baseline > gpurun_out/bench_push.json 2> gpurun_out/bench_push.err; tail -2 gpurun_out/bench_push.err
STCSP_NO_PUSH=1 python bench.py --steps 50 --warmup 5 --no-also --no-cold --no-cpu-baseline > gpurun_out/bench_nopush.json 2> gpurun_out/bench_nopush.err
python - <<'PY'
import json
for f in ('push','nopush'):
    d=json.loads([l for l in open('gpurun_out/bench_%s.json'%f) if l.startswith('{')][-1])
    print(f, 'device', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['parity']['sha256_ok'], d['gpu_launches'])
PY
