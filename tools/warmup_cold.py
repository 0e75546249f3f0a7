"""First solves of models after stcsp_gpu_warmup() in a fresh process.  usage: warmup_cold.py [MB] NAME ..."""
import sys
import time

sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances

mb = int(sys.argv[1])
t0 = time.perf_counter()
binding.warmup(-1, mb << 20)
print("warmup(%d MiB): %.1f ms" % (mb, (time.perf_counter() - t0) * 1e3), flush=True)
for name in sys.argv[2:]:
    model = binding.Model(instances.by_name(name))
    for i in range(3):
        t0 = time.perf_counter()
        a = binding.solve(model, binding.default_options(verbosity=1 if i == 0 else 0))
        w = (time.perf_counter() - t0) * 1e3
        print("%s solve %d: e2e %.3f ms device %.3f ms launches %d waves %d" % (name, i, w, a.c.solve_ms, a.c.n_kernel_launches, a.c.n_waves), flush=True)
        del a
