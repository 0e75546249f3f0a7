"""Front end + normaliser against the reference: variable order (= edge-label columns), signature
variables, variable and constraint counts of every golden; language corner cases of SURVEY.md Appendix A."""
import pytest

from conftest import GOLDENS, golden_flags, golden_text
from stcsp_solver_b200 import binding, instances


@pytest.mark.parametrize("key", sorted(k for k, g in GOLDENS.items()
                                        if "sha256" in g and not g.get("flags") and "stat" in g))     # reference output only
def test_variables_and_counts_match_reference(key):
    g = GOLDENS[key]
    model = binding.Model(golden_text(g))
    assert "# " + " ".join(model.var_names) == g["header_vars"]
    assert model.n_vars == g["stat"]["vars"]
    assert model.n_constraints == g["stat"]["cons"]
    dump = model.dump()
    sig = []
    for line in dump.split("\n"):
        if line.startswith("NEXT"):
            x = line.split()[1]
            if x not in sig:
                sig.append(x)
    order = {n: i for i, n in enumerate(model.var_names)}
    sig.sort(key=order.get)
    assert ("# " + " ".join(sig)).rstrip() == g["header_sig"].rstrip()


def test_shipped_examples_regenerate_exactly():
    """instances.py must reproduce the reference's example files where they exist (this container only)."""
    import os
    ex = "/root/reference/examples"
    if not os.path.isdir(ex):
        pytest.skip("reference not present")
    for name in instances.SHIPPED:
        a = binding.Model(open(os.path.join(ex, name + ".csp")).read()).dump()
        b = binding.Model(instances.by_name(name)).dump()
        assert a == b, name


@pytest.mark.parametrize("text,msg", [
    ("var X : [0, 3];\nX < ;\n", "Line 2: syntax error"),
    ("var X : [0, 3];\nB0 -1 == X;\n", "Line 2: syntax error"),             # `-1` lexes as a constant
    ("var X : [0, 3];\nY == X;\n", "Variable 'Y' has not been defined."),
    ("var X : [3, 0];\n", "Invalid domain"),
    ("var X_1 : [0, 3];\n", "syntax error"),                                  # no underscore in identifiers
])
def test_errors(text, msg):
    with pytest.raises(binding.StcspError) as e:
        binding.Model(text)
    assert msg in str(e.value)


def test_precedence_and_comments():
    m = binding.Model("// c\nvar X : [0, 3]; ' q\n/* b */ var Y : [0, 3];\nX + 2 * Y - 1 == 3;\nnot X eq Y or X lt Y and Y gt 1 -> X;\n")
    d = m.dump()
    assert "((X + (2 * Y)) - 1) == 3" in d
    assert "not(((X eq Y) or ((X lt Y) and (Y gt 1)))) -> X" in d


def test_fby_and_next_normalisation_order():
    d = binding.Model("var X : [0, 3];\nX == 1 fby 2 fby X;\n").dump()
    lines = [ln for ln in d.split("\n") if ln[:4] in ("POIN", "NEXT")]
    # operands are named left to right, auxiliaries in creation order (pinned by golden probe_fby)
    assert [ln.split(" ;")[0] for ln in lines] == [
        "POINT _V0 == 1", "POINT _V1 == 2", "POINT* first(_V2) == first(_V1)", "NEXT X == next(_V2)",
        "POINT _V3 == _V2", "POINT* first(_V4) == first(_V0)", "NEXT _V3 == next(_V4)", "POINT X == _V4"]
