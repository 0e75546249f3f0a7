"""Per-wave log of one solve (library verbosity 1 + per-launch CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
name = sys.argv[1] if len(sys.argv) > 1 else "juggling_b6_f6_nosym"
kw = {}
if len(sys.argv) > 2:
    kw["enum_limit_now"] = int(sys.argv[2])
if len(sys.argv) > 3:
    kw["enum_limit_ahead"] = int(sys.argv[3])
m = binding.Model(instances.by_name(name))
binding.solve(m)        # warm-up
a = binding.solve(m, binding.default_options(profile_kernels=1, verbosity=int(__import__('os').environ.get('V','1')), **kw))
print(a.stats())
