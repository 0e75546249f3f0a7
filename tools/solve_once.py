"""Two warm-up solves, then ONE solve on the default path: the target of ncu captures of search_kernel
(use --launch-skip 2 --launch-count 1 -k regex:search_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
name = sys.argv[1] if len(sys.argv) > 1 else "juggling_b6_f6_nosym"
m = binding.Model(instances.by_name(name))
for _ in range(3):
    a = binding.solve(m)
print(a.stats())
