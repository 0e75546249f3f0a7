"""Repeated group solves of different models with W ranks on device 0 (the sequence of test_group_solve_repeated_and_mixed)."""
import json
import os
import sys
sys.path.insert(0, ".")
from stcsp_solver_b200 import binding, instances
world = int(sys.argv[1]) if len(sys.argv) > 1 else 3
names = ["partialorder_11", "probe_first_capture", "juggling_b5_f6", "partialorder_11", "digitinvader3", "probe_first_capture"]
for name in names * 2:
    g = json.load(open(os.path.join("tests", "golden", name + ".json")))
    model = binding.Model(g["model"] if "model" in g else instances.by_name(name))
    a, xs = binding.solve_multi(model, world, binding.default_options(shard_mode=1), devices=[0] * world)
    ok = binding.Solution(model, a).canonical_sha256() == g["sha256"]
    print(name, "OK" if ok else "BAD", a.c.n_states, a.c.n_edges, xs["waves"], xs["exchanges"], flush=True)
