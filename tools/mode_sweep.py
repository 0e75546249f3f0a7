"""Device ms per forced expand mode (0 auto, 1 warp, 2 CTA, 3 quad)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
for name in sys.argv[1:]:
    m = binding.Model(instances.by_name(name))
    binding.solve(m)
    out = []
    for mode in (0, 1, 3):
        best = None
        for _ in range(3):
            st = binding.solve(m, binding.default_options(expand_mode=mode)).stats()
            best = st["solve_ms"] if best is None else min(best, st["solve_ms"])
        out.append("mode %d: %8.3f ms" % (mode, best))
    print("%-22s %s" % (name, "  ".join(out)), flush=True)
