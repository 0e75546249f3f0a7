set -x
N=${NGPU:-4}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-cold ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','parity')}); print(d['e2e'])
for a in d.get('sharded',[]): print(json.dumps(a)[:900])
PY
python tools/threads_probe.py partialorder_18 $N 2>&1 | tail -2
