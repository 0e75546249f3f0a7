"""The independent semantic oracle (oracle/semantic_oracle.cpp) against the reference itself.

The semantic oracle computes the automaton from its definition (BFS over signatures, every
complete assignment checked against every constraint, fail rule and liveness as fixpoints) and
shares no search code with oracle/stcsp_oracle.cpp.  It is PINNED here against every golden that
oracle/_ref/stcsp_ref (the unmodified reference) produced: all shipped models, the -a / -z sweep
over digitinvader1-9, the feature probes.  Only because of that may tests/golden/semantic_*.json
(instances the reference cannot finish) serve as goldens for the CUDA path.
"""
import hashlib

import pytest

from conftest import GOLDENS, golden_flags, golden_text
from stcsp_solver_b200 import binding, instances

import _oracle

REFERENCE = sorted(k for k, g in GOLDENS.items() if "sha256" in g and g.get("source") != "semantic_oracle")
SEMANTIC = sorted(k for k, g in GOLDENS.items() if g.get("source") == "semantic_oracle")


def run(g, want_text=False):
    flags = golden_flags(g)
    k = next((int(f[2:]) for f in flags if f.startswith("-k")), 2)
    model = binding.Model(golden_text(g), k)
    return _oracle.semantic(model, "-a" in flags, "-z" in flags, want_text=want_text)


@pytest.mark.parametrize("key", REFERENCE)
def test_semantic_oracle_matches_reference(key):
    g = GOLDENS[key]
    r = run(g, want_text="canonical" in g)
    assert r["sha256"] == g["sha256"]
    assert (r["n_states"], r["n_edges"]) == (g["states"], g["edges"])
    if "canonical" in g:
        assert r["text"] == g["canonical"]
        assert hashlib.sha256(r["text"].encode()).hexdigest() == r["sha256"]     # the oracle's own SHA-256 routine
    flags = golden_flags(g)
    if "-a" in flags:
        assert g["stdout"].startswith("adver1: %d; " % r["adver1"])
    if "-z" in flags:
        assert g["stdout"].startswith("adver2: %d\n" % r["adver2"])


def test_reference_goldens_cover_the_adversarial_sweep():
    """BASELINE.json config 2: digitinvader1-9 incl. -a (and -z), all from the reference binary."""
    for n in range(1, 10):
        for suffix in ("", "_a", "_z"):
            g = GOLDENS["digitinvader%d%s" % (n, suffix)]
            assert "sha256" in g and g.get("source") != "semantic_oracle"


@pytest.mark.parametrize("key", [k for k in SEMANTIC if GOLDENS[k].get("oracle_seconds", 1e9) <= 15])
def test_semantic_goldens_reproduce(key):
    """The committed semantic goldens are what the pinned oracle computes (small ones re-derived here)."""
    g = GOLDENS[key]
    r = run(g)
    assert (r["sha256"], r["n_states"], r["n_edges"]) == (g["sha256"], g["states"], g["edges"])


def test_semantic_goldens_agree_with_closed_forms():
    """juggling _nosym has 1 + F!/(F-B)! states; every non-root state of b == f has exactly
    one successor per ... (edges = 2 per state for b == f); partialorder doubles per step."""
    import math
    for key in SEMANTIC:
        g = GOLDENS[key]
        name = g["name"]
        if name.startswith("juggling"):
            b, f = int(name.split("_")[1][1:]), int(name.split("_")[2][1:])
            assert g["states"] == 1 + math.factorial(f) // math.factorial(f - b)
            if b == f:
                assert g["edges"] == 2 * (g["states"] - 1)
        if name.startswith("partialorder_"):
            n = int(name.split("_")[1])
            # states: 2^(n+1) - 2^(K+1)... measured law on the reference goldens 10..14: states(n) = 63 * 2^(n-5) for odd K
            prev = GOLDENS.get("semantic_partialorder_%d" % (n - 1)) or GOLDENS.get("partialorder_%d" % (n - 1))
            if prev:
                assert 1.9 < g["states"] / prev["states"] < 2.1
                assert 2.0 < g["edges"] / prev["edges"] < 2.4


def test_every_generated_benchmark_instance_has_a_golden():
    """BASELINE.json config 5 and SURVEY.md 8(d): the synthetic instances are pinned by canonical hash, not by counts."""
    for name in ["juggling_b7_f7_nosym", "juggling_b8_f8_nosym"] + ["partialorder_%d" % n for n in (15, 16, 17, 18)]:
        assert "semantic_" + name in GOLDENS, name
        assert instances.by_name(name)


def test_two_independent_oracles_agree_on_random_models():
    """700 seeded random models (tests/model_fuzz.py: every operator, arrays, fby, @, until, variable-free
    constraints): the DFS restatement of the reference and the semantic oracle give the same canonical automaton."""
    from model_fuzz import random_model
    compared = 0
    for seed in range(700):
        try:
            model = binding.Model(random_model(seed))
        except binding.StcspError:
            continue
        automaton, _ = _oracle.solve(model, 2.0)
        if automaton is None:
            continue
        want = binding.Solution(model, automaton).canonical_sha256_streamed()
        assert _oracle.semantic(model)["sha256"] == want, seed
        compared += 1
    assert compared >= 650
