import sys, time
sys.path.insert(0, '.')
from stcsp_solver_b200 import binding, instances
name, w = sys.argv[1], int(sys.argv[2])
model = binding.Model(instances.by_name(name))
for i in range(3):
    t0 = time.perf_counter()
    a, xs = binding.solve_multi(model, w, binding.default_options(shard_mode=1))
    wall = (time.perf_counter() - t0) * 1e3
    print("threads: %s x%d device %.2f ms e2e %.2f ms states %d edges %d %s" % (name, w, a.c.solve_ms, wall, a.c.n_states, a.c.n_edges, xs), flush=True)
