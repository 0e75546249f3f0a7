"""Eager vs lazy look-ahead: automata must agree; prints device ms, nodes, states for both."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDENS, golden_flags, golden_text
from stcsp_solver_b200 import binding
only = sys.argv[1:] 
bad = 0
for key, g in sorted(GOLDENS.items(), key=lambda kv: kv[1].get("wall_s", 0)):
    if "sha256" not in g or g.get("flags"):
        continue
    if only and not any(o in key for o in only):
        continue
    model = binding.Model(golden_text(g))
    res = []
    for mode in (1, 2):
        best = None
        for _ in range(3):
            a = binding.solve(model, binding.default_options(lookahead=mode))
            st = a.stats()
            if best is None or st["solve_ms"] < best[0]["solve_ms"]:
                best = (st, binding.Solution(model, a).canonical_sha256())
        res.append(best)
    ok = res[0][1] == res[1][1] == g["sha256"]
    bad += not ok
    print("%-26s %s eager %8.3f ms nodes %8d states %6d | lazy %8.3f ms nodes %8d states %6d" % (
        key, "OK " if ok else "BAD", res[0][0]["solve_ms"], res[0][0]["n_search_nodes"], res[0][0]["n_states"],
        res[1][0]["solve_ms"], res[1][0]["n_search_nodes"], res[1][0]["n_states"]), flush=True)
print("bad:", bad)
