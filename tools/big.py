"""Large generated instances on one GPU: time, counts, memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
for name in sys.argv[1:]:
    m = binding.Model(instances.by_name(name))
    t0 = time.time()
    try:
        a = binding.solve(m, binding.default_options(time_limit_s=600))
    except Exception as e:
        print(name, "FAILED", e, flush=True); continue
    st = a.stats()
    t1 = time.time()
    sol = binding.Solution(m, a)
    print("%-22s states %9d edges %10d (reachable %d/%d) nodes %10d waves %4d dev_ms %10.1f wall_s %7.2f post_s %6.2f alg_GB %.2f" % (
        name, st["n_states"], st["n_edges"], sol.n_states, sol.n_edges, st["n_search_nodes"], st["n_waves"], st["solve_ms"],
        t1 - t0, time.time() - t1, st["algorithmic_bytes"] / 1e9), flush=True)
