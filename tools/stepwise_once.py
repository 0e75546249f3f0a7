"""One warm-up solve, then ONE solve on the step-wise path (one expand launch per wave): the target of ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
name = sys.argv[1] if len(sys.argv) > 1 else "partialorder_14"
m = binding.Model(instances.by_name(name))
binding.solve(m)
a = binding.solve(m, binding.default_options(profile_kernels=1))
print(a.stats())
