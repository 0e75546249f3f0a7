"""Device ms (best of 15 steady solves) per instance; run with STCSP_FAN_WARPS=n to vary the multi-variable branching fan."""
import os, sys
sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances
names = sys.argv[1:]
out = []
for name in names:
    m = binding.Model(instances.by_name(name))
    best = 1e9
    for i in range(18):
        a = binding.solve(m)
        if i >= 3:
            best = min(best, a.c.solve_ms)
        w, n = a.c.n_waves, a.c.n_search_nodes
        del a
    out.append("%s %.3f ms (%d waves, %d nodes)" % (name, best, w, n))
print("fan_warps=%s: " % os.environ.get("STCSP_FAN_WARPS", "default") + "; ".join(out), flush=True)
