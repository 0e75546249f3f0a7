"""Best-of-N device time per instance (steady state: allocator cache warm)."""
import sys
from stcsp_solver_b200 import binding, instances
names = sys.argv[1:] or ["juggling_b6_f6_nosym", "partialorder_14", "digitinvader9", "juggling_b5_f6", "digitinvader5"]
for name in names:
    m = binding.Model(instances.by_name(name))
    binding.solve(m)
    best = None
    for _ in range(5):
        a = binding.solve(m, binding.default_options(profile_kernels=1))
        st = a.stats()
        if best is None or st["solve_ms"] < best["solve_ms"]:
            best = st
    print("%-24s dev_ms %8.3f expand_ms %8.3f wall_ms %8.3f nodes %8d tuples %10d revisions %9d waves %4d launches %d" % (
        name, best["solve_ms"], best["expand_ms"], best["wall_ms"], best["n_search_nodes"], best["n_tuples"],
        best["n_revisions"], best["n_waves"], best["n_kernel_launches"]), flush=True)
