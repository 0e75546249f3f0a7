"""The reference-side binding of the C ABI, compiled (oracle/gpu_bridge.cpp, shown in INTEGRATION.md).

oracle/_ref/stcsp_ref_gpu = the reference's own objects (front end stand-in, normaliser, post-processing, DOT writer)
with the single call solverSolve(Solver*, bool) (reference src/solveralgorithm.h:11) redirected to the bridge, which
flattens `Solver`, calls stcsp_gpu_solve and rebuilds the reference's `Graph`.  Its solutions.dot must canonicalise to the
goldens the unmodified reference produced -- which checks the boundary (b) and, independently of the product's own
post-processing, the automaton the GPU returns (the reference's graphTraverse / adversarialTraverse run on it).

Without a GPU the same bridge is exercised with the CPU oracle behind the C-ABI symbols (stcsp_ref_bridge_cpu).
"""
import os
import resource
import subprocess

import pytest

from conftest import GOLDENS, ROOT, golden_flags, golden_text
from stcsp_solver_b200 import canonical

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REFERENCE = {k: g for k, g in GOLDENS.items() if "sha256" in g and g.get("source") != "semantic_oracle"}
FAST = sorted(k for k, g in REFERENCE.items() if g.get("wall_s", 99) <= 1.0)
ALL = sorted(REFERENCE)


def run_bridge(binary, g, tmp_path):
    exe = os.path.join(REF_DIR, binary)
    if not os.path.exists(exe):
        pytest.skip("%s not built (reference sources absent when the oracle was built)" % binary)
    p = tmp_path / (g["name"] + ".csp")
    p.write_text(golden_text(g))

    def unlimited_stack():
        resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    r = subprocess.run([exe, "-s"] + golden_flags(g) + [str(p)], cwd=tmp_path, capture_output=True, text=True,
                       preexec_fn=unlimited_stack)
    assert r.returncode == 0, r.stderr
    a = canonical.parse_dot((tmp_path / "solutions.dot").read_text())
    assert canonical.counts(a) == (g["states"], g["edges"])
    assert canonical.canonical_sha256(a) == g["sha256"]
    # the adver prefixes and the variable / constraint counts of the stat line are the reference's own
    assert r.stdout.split("\t")[1:3] == g["stdout"].split("\t")[1:3]
    if "-a" in golden_flags(g) or "-z" in golden_flags(g):
        assert r.stdout.split("0.00\t")[0] == g["stdout"].split("0.00\t")[0]


@pytest.mark.parametrize("key", FAST)
def test_bridge_code_with_the_cpu_oracle_behind_the_abi(key, tmp_path):
    run_bridge("stcsp_ref_bridge_cpu", REFERENCE[key], tmp_path)


# Every run is a fresh process that creates a CUDA context (2-3 s on a cold box), so the default suite takes a representative
# third of the cases -- every model family, every flag, every feature group of the probes; STCSP_FULL_BRIDGE=1 runs them all
# (all green on B200, gpurun_out/pytest_gpu_r02l.log: 81 cases).
BRIDGE_GPU = [k for k in ALL if REFERENCE[k]["edges"] <= 400000]
if not os.environ.get("STCSP_FULL_BRIDGE"):
    _KEEP = ("juggling_b4_f4", "juggling_b5_f6", "juggling_b5_f6_nosym", "juggling_b6_f6_nosym", "partialorder_10", "partialorder_12",
             "digitinvader1", "digitinvader3", "digitinvader5", "digitinvader8", "digitinvader1_a", "digitinvader3_a", "digitinvader3_z",
             "digitinvader7_z", "digitinvader9_a")
    BRIDGE_GPU = [k for k in BRIDGE_GPU if k in _KEEP or (k.startswith("probe_") and any(
        t in k for t in ("adversarial", "until_two", "first_capture", "at2", "fby_expr", "k1", "k3", "unsat", "dead", "quirk", "arr")))]


@pytest.mark.gpu
@pytest.mark.parametrize("key", BRIDGE_GPU)
def test_reference_host_around_the_gpu_search(key, tmp_path):
    run_bridge("stcsp_ref_gpu", REFERENCE[key], tmp_path)
