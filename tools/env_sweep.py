"""Device / e2e ms of steady-state solves under an environment knob of the library (read once per process, so one process per
value).  usage: env_sweep.py VAR v1,v2,... NAME [NAME ...]"""
import os
import subprocess
import sys

SNIPPET = r"""
import sys, time
sys.path.insert(0, '.')
from stcsp_solver_b200 import binding, instances
for name in sys.argv[1:]:
    m = binding.Model(instances.by_name(name))
    for _ in range(3):
        binding.solve(m)
    best = None
    for _ in range(5):
        t0 = time.perf_counter()
        a = binding.solve(m)
        w = (time.perf_counter() - t0) * 1e3
        if best is None or a.c.solve_ms < best[0]:
            best = (a.c.solve_ms, w, a.c.n_search_nodes, a.c.n_waves, a.c.n_states, a.c.n_edges)
        del a
    print("%-24s device %8.3f ms  e2e %8.3f ms  nodes %d waves %d states %d edges %d" % ((name,) + best), flush=True)
"""

var, values, names = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
for v in values:
    env = dict(os.environ)
    if v != "-":
        env[var] = v
    print("== %s=%s" % (var, v), flush=True)
    subprocess.run([sys.executable, "-c", SNIPPET] + names, env=env)
