"""Device time per instance for several values of wide_wave_nodes (the width above which waves leave the persistent
kernel and run as separate launches); -1 = never."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
names = sys.argv[1:] or ["partialorder_14", "partialorder_16", "partialorder_18", "digitinvader9", "juggling_b8_f8_nosym"]
for name in names:
    m = binding.Model(instances.by_name(name))
    binding.solve(m)
    row = []
    for wide in (-1, 16384, 32768, 65536, 131072, 262144):
        best = min(binding.solve(m, binding.default_options(wide_wave_nodes=wide)).stats()["solve_ms"] for _ in range(4))
        row.append("%d: %.3f" % (wide, best))
    print("%-22s %s" % (name, "  ".join(row)), flush=True)
