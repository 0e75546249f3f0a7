"""Sweep the enumeration budgets; prints device ms per setting (best of 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcsp_solver_b200 import binding, instances
names = sys.argv[1:] or ["juggling_b6_f6_nosym", "partialorder_14", "digitinvader9", "juggling_b5_f6"]
for name in names:
    m = binding.Model(instances.by_name(name))
    binding.solve(m)
    for now in (2, 4, 8, 16, 64):
        for ahead in (1, 2, 4, 8):
            best = None
            for _ in range(3):
                a = binding.solve(m, binding.default_options(enum_limit_now=now, enum_limit_ahead=ahead))
                st = a.stats()
                if best is None or st["solve_ms"] < best["solve_ms"]:
                    best = st
            print("%-24s now %6d ahead %5d: dev_ms %8.3f nodes %8d fails %7d tuples %10d rev %9d states %6d" % (
                name, now, ahead, best["solve_ms"], best["n_search_nodes"], best["n_fails"], best["n_tuples"], best["n_revisions"], best["n_states"]), flush=True)
