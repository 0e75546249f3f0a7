"""Best-of-N device time per instance (steady state: caches warm): default path (persistent search kernel) and
the step-wise path with per-launch expand timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
names = sys.argv[1:] or ["juggling_b6_f6_nosym", "partialorder_14", "digitinvader9", "juggling_b5_f6", "digitinvader5"]
for name in names:
    m = binding.Model(instances.by_name(name))
    binding.solve(m)
    best = prof = None
    for _ in range(5):
        st = binding.solve(m).stats()
        if best is None or st["solve_ms"] < best["solve_ms"]:
            best = st
        st = binding.solve(m, binding.default_options(profile_kernels=1)).stats()
        if prof is None or st["solve_ms"] < prof["solve_ms"]:
            prof = st
    print("%-22s dev_ms %8.3f wall_ms %8.3f launches %3d | stepwise dev_ms %8.3f expand_ms %8.3f launches %3d | nodes %8d tuples %10d rev %9d waves %4d" % (
        name, best["solve_ms"], best["wall_ms"], best["n_kernel_launches"], prof["solve_ms"], prof["expand_ms"],
        prof["n_kernel_launches"], best["n_search_nodes"], best["n_tuples"], best["n_revisions"], best["n_waves"]), flush=True)
