"""The C-ABI library loads and exports every symbol include/*.h declares; host-only entry points work
without a GPU; GPU entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from stcsp_solver_b200 import binding, instances


def declared_symbols():
    names = set()
    for h in ("stcsp_b200.h", "stcsp_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(stcsp_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(binding.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s


def test_struct_sizes_match_header():
    """ctypes mirrors must have the layout the C compiler gives the header's structs."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "stcsp_host.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(stcsp_problem_t),' \
          'sizeof(stcsp_options_t), sizeof(stcsp_automaton_t), sizeof(stcsp_solution_t), sizeof(stcsp_tok_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(binding.Problem), C.sizeof(binding.Options), C.sizeof(binding.AutomatonC),
                     C.sizeof(binding.SolutionC), C.sizeof(binding.Tok)]


def test_header_is_plain_c():
    import subprocess
    for h in ("stcsp_b200.h", "stcsp_host.h"):
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                               os.path.join(ROOT, "include", h)])


@pytest.mark.skipif(binding.lib().stcsp_gpu_device_count() > 0, reason="a GPU is present")
def test_solve_without_gpu_fails_loudly():
    model = binding.Model(instances.by_name("juggling_b4_f4"))
    with pytest.raises(binding.StcspError) as e:
        binding.solve(model)
    assert e.value.status == binding.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


@pytest.mark.skipif(binding.lib().stcsp_gpu_device_count() > 0, reason="a GPU is present")
def test_warmup_without_gpu_fails_loudly():
    with pytest.raises(binding.StcspError) as e:
        binding.warmup(-1, 1 << 20)
    assert e.value.status == binding.ERR_CUDA


def test_cli_without_gpu_or_with(tmp_path):
    import subprocess
    p = tmp_path / "m.csp"
    p.write_text(instances.by_name("juggling_b4_f4"))
    r = subprocess.run([os.path.join(ROOT, "bin", "stcsp"), "-s", str(p)], cwd=tmp_path, capture_output=True, text=True)
    if binding.lib().stcsp_gpu_device_count() > 0:
        assert r.returncode == 0 and r.stdout.count("\t") == 7
        assert (tmp_path / "solutions.dot").exists()
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
    r = subprocess.run([os.path.join(ROOT, "bin", "stcsp"), str(tmp_path / "bad.csp")], capture_output=True, text=True)
    assert r.returncode == 1
    q = tmp_path / "syntax.csp"
    q.write_text("var X : [0, 3];\nX < ;\n")
    r = subprocess.run([os.path.join(ROOT, "bin", "stcsp"), str(q)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.strip() == "Line 2: syntax error"
