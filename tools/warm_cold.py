"""First solve of a model in a WARM process (another model solved first, then a pause so that one-time background set-up has
finished), with the library's timeline on stderr.  usage: warm_cold.py NAME [NAME ...]"""
import sys
import time

sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances

names = sys.argv[1:] or ["partialorder_14", "digitinvader9"]
binding.solve(binding.Model(instances.by_name("juggling_b4_f5_nosym")))
time.sleep(float(__import__("os").environ.get("PAUSE", "1.0")))
for name in names:
    model = binding.Model(instances.by_name(name))
    for i in range(3):
        t0 = time.perf_counter()
        a = binding.solve(model, binding.default_options(verbosity=1 if i == 0 else 0))
        w = (time.perf_counter() - t0) * 1e3
        print("%s solve %d: e2e %.3f ms device %.3f ms launches %d waves %d" % (name, i, w, a.c.solve_ms, a.c.n_kernel_launches, a.c.n_waves), flush=True)
        del a
