"""Quick GPU sanity run: every golden, states/edges/sha vs the reference; prints one line per case."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDENS, golden_flags, golden_text
from stcsp_solver_b200 import binding

only = sys.argv[1] if len(sys.argv) > 1 else ""
bad = 0
for key, g in sorted(GOLDENS.items(), key=lambda kv: kv[1].get("wall_s", 0)):
    if "sha256" not in g or only not in key:
        continue
    flags = golden_flags(g)
    k = next((int(f[2:]) for f in flags if f.startswith("-k")), 2)
    t0 = time.time()
    try:
        model = binding.Model(golden_text(g), k)
        a = binding.solve(model)
        sol = binding.Solution(model, a, "-a" in flags, "-z" in flags)
        ok = sol.canonical_sha256() == g["sha256"]
        st = a.stats()
        print("%-28s %s states %d/%d edges %d/%d nodes %d fails %d waves %d tuples %d dev_ms %.2f wall_ms %.1f ref_s %.2f" % (
            key, "OK " if ok else "BAD", sol.n_states, g["states"], sol.n_edges, g["edges"], st["n_search_nodes"],
            st["n_fails"], st["n_waves"], st["n_tuples"], st["solve_ms"], (time.time() - t0) * 1e3, g.get("wall_s", 0)), flush=True)
        bad += not ok
    except Exception as e:
        print("%-28s EXC %s" % (key, e), flush=True)
        bad += 1
print("bad:", bad)
