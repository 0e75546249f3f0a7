"""One group solve with W ranks on device 0 (threads of this process); STCSP_TRACE_GROUP=1 shows who waits where."""
import sys
import time
sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances
name, world = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = binding.Model(instances.by_name(name))
for i in range(reps):
    t0 = time.perf_counter()
    a, xs = binding.solve_multi(model, world, binding.default_options(shard_mode=1), devices=[0] * world)
    print("%s world %d: e2e %.2f ms device %.2f ms states %d edges %d %s" % (name, world, (time.perf_counter() - t0) * 1e3, a.c.solve_ms,
                                                                             a.c.n_states, a.c.n_edges, xs), flush=True)
