"""Steady-state solve with the library's timeline on stderr: verbosity 3 = per-wave stamps of the search kernel and block 0's
node timeline, verbosity 1 = one line per kernel return / stand-alone wave.  usage: wave_trace.py NAME VERBOSITY [key=value ...]"""
import sys
import time

sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances

name, verb = sys.argv[1], int(sys.argv[2])
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[3:])}
model = binding.Model(instances.by_name(name))
for i in range(3):
    a = binding.solve(model, binding.default_options(**kw))
    del a
t0 = time.perf_counter()
a = binding.solve(model, binding.default_options(verbosity=verb, **kw))
print("%s: e2e %.3f ms device %.3f ms search-kernel %.3f ms launches %d waves %d nodes %d" % (
    name, (time.perf_counter() - t0) * 1e3, a.c.solve_ms, a.c.expand_ms, a.c.n_kernel_launches, a.c.n_waves, a.c.n_search_nodes), flush=True)
