set -x
nvidia-smi topo -m | head -8
N=${NGPU:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py partialorder_11 probe_first_capture partialorder_14 partialorder_16 partialorder_18 > gpurun_out/multi_check_$N.log 2>&1; grep -v "^\[W\|^$" gpurun_out/multi_check_$N.log | tail -12 | cut -c1-700
python tools/wave_trace.py partialorder_18 0 > gpurun_out/po18_single.log 2>&1; tail -1 gpurun_out/po18_single.log
printf 'x' > /dev/null
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from stcsp_solver_b200 import binding, instances
import torch
n = torch.cuda.device_count()
for name in ("partialorder_14", "partialorder_18"):
    model = binding.Model(instances.by_name(name))
    for w in ([2] if n >= 2 else []) + ([4] if n >= 4 else []) + ([8] if n >= 8 else []):
        for i in range(3):
            t0 = time.perf_counter()
            a, xs = binding.solve_multi(model, w, binding.default_options(shard_mode=1))
            wall = (time.perf_counter() - t0) * 1e3
        print("threads: %s x%d device %.2f ms e2e %.2f ms states %d edges %d %s" % (name, w, a.c.solve_ms, wall, a.c.n_states, a.c.n_edges, xs), flush=True)
PY
