import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
name = sys.argv[1] if len(sys.argv) > 1 else "juggling_b6_f6_nosym"
m = binding.Model(instances.by_name(name))
binding.solve(m); binding.solve(m)
a = binding.solve(m, binding.default_options(verbosity=3))
print(a.stats())
