"""Repeat solves of several instances (deadlock / race hunt for the persistent kernel); prints counts that must not vary."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
plan = [("juggling_b6_f6_nosym", 3000), ("juggling_b5_f6", 500), ("digitinvader5", 300), ("partialorder_12", 300),
        ("partialorder_16", 20), ("digitinvader9", 30), ("partialorder_18", 3)]
for name, reps in plan:
    m = binding.Model(instances.by_name(name))
    seen = set()
    t0 = time.time()
    for i in range(reps):
        st = binding.solve(m).stats()
        seen.add((st["n_states"], st["n_edges"], st["n_search_nodes"], st["n_fails"], st["n_leaves"], st["n_dominance"], st["n_waves"]))
    print("%-22s reps %5d distinct results %d %s  %.1f s" % (name, reps, len(seen), sorted(seen)[0], time.time() - t0), flush=True)
