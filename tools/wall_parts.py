"""Host-side wall-time split of one steady-state solve (verbosity 1 prints init / waves / download / release)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
name = sys.argv[1] if len(sys.argv) > 1 else "juggling_b6_f6_nosym"
m = binding.Model(instances.by_name(name))
for _ in range(5):
    binding.solve(m)
for _ in range(3):
    a = binding.solve(m, binding.default_options(verbosity=1))
    print(a.stats()["solve_ms"], a.stats()["wall_ms"])
