"""Device / e2e ms of steady-state solves under values of one stcsp_options_t field.  usage: opt_sweep.py FIELD v1,v2,... NAME ..."""
import sys
import time

sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances

field, values, names = sys.argv[1], [int(v) for v in sys.argv[2].split(",")], sys.argv[3:]
for name in names:
    m = binding.Model(instances.by_name(name))
    for v in values:
        opts = lambda: binding.default_options(**{field: v})
        for _ in range(2):
            binding.solve(m, opts())
        best = None
        for _ in range(4):
            t0 = time.perf_counter()
            a = binding.solve(m, opts())
            w = (time.perf_counter() - t0) * 1e3
            if best is None or a.c.solve_ms < best[0]:
                best = (a.c.solve_ms, w, a.c.n_search_nodes, a.c.n_fails, a.c.n_waves, a.c.n_states, a.c.n_edges)
            del a
        print("%-24s %s=%d device %8.3f ms  e2e %8.3f ms  nodes %d fails %d waves %d states %d edges %d" % ((name, field, v) + best), flush=True)
