import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
name, mode = sys.argv[1], sys.argv[2]
m = binding.Model(instances.by_name(name))
t0 = time.time()
opts = binding.default_options(time_limit_s=100, verbosity=1, profile_kernels=1 if mode == "stepwise" else 0)
a = binding.solve(m, opts)
st = a.stats()
print(name, mode, "states", st["n_states"], "edges", st["n_edges"], "dev_ms %.1f wall_s %.2f" % (st["solve_ms"], time.time() - t0), flush=True)
