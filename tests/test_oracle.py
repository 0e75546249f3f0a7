"""The oracle (oracle/stcsp_oracle.cpp) against the reference itself.

tests/golden/*.json were produced by oracle/make_goldens.py running oracle/_ref/stcsp_ref -- the
unmodified reference solver -- so agreement here pins the CPU restatement: same canonical
automaton AND same search statistics (the reference's stat line), which requires the same search
tree and the same propagation strength.
"""
import pytest

from conftest import GOLDENS, golden_flags, golden_text
from stcsp_solver_b200 import binding

import _oracle

# wall seconds of the reference run; cases above the bound are covered by the GPU parity tests only
REF = {k: g for k, g in GOLDENS.items() if "sha256" in g and g.get("source") != "semantic_oracle"}   # reference output only
FAST = sorted(k for k, g in REF.items() if g.get("wall_s", 0) <= 2.5)
SLOW = sorted(k for k, g in REF.items() if 2.5 < g.get("wall_s", 0) <= 30)


def check(key):
    g = GOLDENS[key]
    flags = golden_flags(g)
    k = next((int(f[2:]) for f in flags if f.startswith("-k")), 2)
    model = binding.Model(golden_text(g), k)
    automaton, stats = _oracle.solve(model)
    sol = binding.Solution(model, automaton, "-a" in flags, "-z" in flags)
    assert sol.canonical_sha256() == g["sha256"]
    if "canonical" in g:
        assert sol.canonical_text() == g["canonical"]
    assert (sol.n_states, sol.n_edges) == (g["states"], g["edges"])
    st = g["stat"]
    assert model.n_vars == st["vars"] and model.n_constraints == st["cons"]
    assert stats["num_nodes"] == st["nodes"]
    assert stats["num_fails"] == st["fails"]
    assert stats["num_dominance"] == st["dominance"]
    if "-a" in flags:
        assert g["stdout"].startswith("adver1: %d; " % sol.adver1)
    if "-z" in flags:
        assert g["stdout"].startswith("adver2: %d\n" % sol.adver2)


@pytest.mark.parametrize("key", FAST)
def test_oracle_matches_reference(key):
    check(key)


@pytest.mark.slow
@pytest.mark.parametrize("key", SLOW)
def test_oracle_matches_reference_slow(key):
    check(key)


def test_error_goldens_are_errors():
    """Cases where the reference exits with an error must be rejected by the front end too."""
    for key, g in GOLDENS.items():
        if "sha256" in g:
            continue
        with pytest.raises(binding.StcspError):
            binding.Model(golden_text(g))
