"""The random-model generator itself (CPU): models parse, the oracle solves them, output is a valid automaton."""
import pytest

import _oracle
from model_fuzz import random_model
from stcsp_solver_b200 import binding


@pytest.mark.parametrize("seed", list(range(0, 40)) + list(range(300, 340)))
def test_random_model_runs_through_front_end_and_oracle(seed):
    text = random_model(seed)
    try:
        model = binding.Model(text)
    except binding.StcspError as e:
        assert e.status in (binding.ERR_PARSE, binding.ERR_INVALID)     # e.g. `@` on next: rejected by design
        return
    automaton, stats = _oracle.solve(model, 2.0)
    if automaton is None:
        return                                                          # too large for a unit test
    sol = binding.Solution(model, automaton)
    assert sol.n_states >= 0 and sol.n_edges >= 0
    assert sol.canonical_text() == binding.Solution(model, _oracle.solve(model, 2.0)[0]).canonical_text()
