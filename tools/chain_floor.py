"""Per-wave floor of the persistent kernel: a 512-state cycle whose waves hold one or two nodes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from stcsp_solver_b200 import binding
from test_gpu_parity import LONG_CHAIN
m = binding.Model(LONG_CHAIN)
for _ in range(3):
    binding.solve(m)
best = min((binding.solve(m).stats() for _ in range(5)), key=lambda st: st["solve_ms"])
print("waves %d nodes %d dev_ms %.3f -> %.2f us per wave" % (best["n_waves"], best["n_search_nodes"], best["solve_ms"],
                                                              best["solve_ms"] * 1e3 / best["n_waves"]))
