"""Parity of the CUDA path (through the C ABI, stcsp_gpu_solve) with the reference.

* every golden automaton produced by the reference itself (tests/golden, all 26 shipped models,
  the -a/-z sweep and the feature probes): canonical text / SHA-256, verdict, state and edge counts;
* the oracle (CPU restatement) on the same inputs, array by array after canonical relabelling;
* instances the reference cannot finish (juggling_b7/b8, partialorder_15-18, digitinvader10-12): the canonical SHA-256
  of the independent semantic oracle (tests/golden/semantic_*.json, oracle/make_semantic_goldens.py), which is itself
  pinned against every reference golden (tests/test_semantic_oracle.py) -- states, edges AND labels, not just counts.
Bit-exact: everything here is integer work.
"""
import math

import pytest

from conftest import GOLDENS, golden_flags, golden_text
from stcsp_solver_b200 import binding, instances

import _oracle

pytestmark = pytest.mark.gpu

# (partialorder_19 / _20 -- 30 M and 63 M edges, 12 GB of automaton -- are checked by bench.py's `also` leg, not here)
CASES = sorted(k for k, g in GOLDENS.items() if "sha256" in g and g.get("edges", 0) <= 20_000_000)


def run_gpu(text, flags=(), **opts):
    """Like `bin/stcsp`: the -a / -z fixpoints (and liveness, for models with `until`) run on the device."""
    k = next((int(f[2:]) for f in flags if f.startswith("-k")), 2)
    model = binding.Model(text, k)
    adversarial = ("-a" in flags) | (("-z" in flags) << 1)
    if adversarial:
        opts = dict(opts, adversarial=adversarial)
    automaton = binding.solve(model, binding.default_options(**opts) if opts else None)
    if adversarial:
        assert automaton.post_applied == 1 | (adversarial << 1)
    return model, automaton, binding.Solution(model, automaton, "-a" in flags, "-z" in flags)


POST_CASES = [k for k in CASES if {"-a", "-z"} & set(golden_flags(GOLDENS[k])) or "until" in golden_text(GOLDENS[k])]


@pytest.mark.parametrize("key", POST_CASES)
def test_device_postprocessing_equals_host_postprocessing(key):
    """Liveness and the -a / -z fixpoints as device sweeps (automaton.cu) against the host restatement of reference
    src/graph.cpp:247-418 (postprocess.cpp) on the same automaton: same flags, same surviving edges, same canonical text."""
    g = GOLDENS[key]
    flags = golden_flags(g)
    model, automaton, sol = run_gpu(golden_text(g), flags)
    plain = binding.solve(model)                         # liveness alone on the device (until) or nothing: the rest on the host
    layered = binding.Solution(model, plain, "-a" in flags, "-z" in flags)
    had = plain.post_applied
    plain.c.post_applied = 0                             # nothing taken from the device
    host = binding.Solution(model, plain, "-a" in flags, "-z" in flags)
    assert "until" not in golden_text(g) or had == 1
    assert sol.canonical_text() == host.canonical_text() == layered.canonical_text()
    assert (sol.adver1, sol.adver2) == (host.adver1, host.adver2) == (layered.adver1, layered.adver2)
    assert sol.canonical_sha256() == g["sha256"]


@pytest.mark.parametrize("key", CASES)
def test_matches_reference_golden(key):
    g = GOLDENS[key]
    flags = golden_flags(g)
    model, automaton, sol = run_gpu(golden_text(g), flags)
    assert (sol.n_states, sol.n_edges) == (g["states"], g["edges"])
    if "canonical" in g:
        assert sol.canonical_text() == g["canonical"]
    assert sol.canonical_sha256_streamed() == g["sha256"]
    if g["edges"] <= 200000:
        assert sol.canonical_sha256() == g["sha256"]           # hashlib over the materialised text
    if "-a" in flags:
        assert g["stdout"].startswith("adver1: %d; " % sol.adver1)
    if "-z" in flags:
        assert g["stdout"].startswith("adver2: %d\n" % sol.adver2)
    # the DOT text re-parsed by the independent Python canonicaliser gives the same hash
    if g["edges"] <= 20000:
        from stcsp_solver_b200 import canonical
        assert canonical.canonical_sha256(sol.to_python()) == g["sha256"]


@pytest.mark.parametrize("lookahead", [1, 3])
@pytest.mark.parametrize("key", [k for k in CASES if GOLDENS[k].get("edges", 0) <= 400_000])
def test_lookahead_policy_does_not_change_the_automaton(key, lookahead):
    """Pointwise constraints at time offsets >= 1 only find a dead end one state early (stcsp_options_t::lookahead): always run
    (1) or never run (3; the default drops them by itself when they never fail anything), the automaton is the reference's --
    dead ends are then created as states and removed by the fail rule (reference src/solveralgorithm.cpp:904-910)."""
    g = GOLDENS[key]
    flags = golden_flags(g)
    _, _, sol = run_gpu(golden_text(g), flags, lookahead=lookahead)
    assert (sol.n_states, sol.n_edges) == (g["states"], g["edges"])
    assert sol.canonical_sha256_streamed() == g["sha256"]


SMALL = [k for k in CASES if GOLDENS[k].get("wall_s", 99) <= 1.0]


@pytest.mark.parametrize("key", SMALL)
def test_matches_oracle(key):
    g = GOLDENS[key]
    flags = golden_flags(g)
    model, automaton, sol = run_gpu(golden_text(g), flags)
    k = next((int(f[2:]) for f in flags if f.startswith("-k")), 2)
    omodel = binding.Model(golden_text(g), k)
    oautomaton, _ = _oracle.solve(omodel)
    osol = binding.Solution(omodel, oautomaton, "-a" in flags, "-z" in flags)
    assert sol.canonical_text() == osol.canonical_text()
    assert sol.dot().split("\n", 1)[1] == osol.dot().split("\n", 1)[1]     # line 1 counts failed states: search-order dependent


@pytest.mark.parametrize("limits", [(1, 1), (8, 2), (64, 64), (1 << 20, 1 << 16)])
def test_enumeration_budget_does_not_change_the_automaton(limits):
    """Propagation strength is not a parity target: any budget gives the same automaton."""
    for name in ("juggling_b4_f5_nosym", "digitinvader2", "partialorder_10", "juggling_b5_f5"):
        g = GOLDENS[name]
        _, _, sol = run_gpu(golden_text(g), (), enum_limit_now=limits[0], enum_limit_ahead=limits[1])
        assert sol.canonical_sha256() == g["sha256"], (name, limits)


@pytest.mark.parametrize("name", ["juggling_b5_f6", "juggling_b5_f6_nosym", "digitinvader3", "partialorder_11", "probe_first_capture",
                                  "probe_at2", "probe_until_two"])
def test_stepwise_path_equals_persistent_kernel(name):
    """profile_kernels=1 runs one expand / route / ingest launch per wave (the path the multi-GPU sessions use);
    the default is the persistent search kernel.  Same automaton, same search statistics."""
    g = GOLDENS[name]
    _, _, s0 = run_gpu(golden_text(g))
    _, _, s3 = run_gpu(golden_text(g), (), profile_kernels=1)
    assert s0.canonical_sha256() == s3.canonical_sha256() == g["sha256"]
    # (statistics: with the look-ahead propagators always on -- the automatic policy samples by node index and decides from
    #  running totals, so the two paths need not drop them for the same nodes)
    _, a1, s1 = run_gpu(golden_text(g), (), lookahead=1)
    _, a2, s2 = run_gpu(golden_text(g), (), profile_kernels=1, lookahead=1)
    assert s1.canonical_sha256() == s2.canonical_sha256() == g["sha256"]
    st1, st2 = a1.stats(), a2.stats()
    for key in ("n_states", "n_edges", "n_search_nodes", "n_fails", "n_leaves", "n_waves"):
        assert st1[key] == st2[key], key
    assert st1["n_kernel_launches"] < st2["n_kernel_launches"]


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("name", ["juggling_b5_f6_nosym", "digitinvader4", "partialorder_12", "probe_until_two", "probe_first_capture"])
def test_every_node_mapping_gives_the_same_automaton(name, mode):
    """expand_mode: a warp per node, a CTA per node, four nodes per warp -- forced for every wave."""
    g = GOLDENS[name]
    for extra in (dict(), dict(profile_kernels=1)):
        _, _, sol = run_gpu(golden_text(g), (), expand_mode=mode, **extra)
        assert sol.canonical_sha256() == g["sha256"], (name, mode, extra)


@pytest.mark.parametrize("wide", [1, 64, 4096, -1])
@pytest.mark.parametrize("name", ["juggling_b5_f6_nosym", "digitinvader4", "partialorder_13", "probe_until_two", "probe_first_capture"])
def test_wide_waves_leaving_the_persistent_kernel(name, wide):
    """wide_wave_nodes: waves wider than this run as stand-alone launches between two runs of the search kernel;
    the automaton and the search statistics do not depend on where the switch happens."""
    g = GOLDENS[name]
    _, _, sol = run_gpu(golden_text(g), (), wide_wave_nodes=wide)
    assert sol.canonical_sha256() == g["sha256"], (name, wide)
    # (statistics with the look-ahead propagators always on: the automatic policy decides from running totals)
    _, base, _ = run_gpu(golden_text(g), (), lookahead=1)
    _, automaton, sol = run_gpu(golden_text(g), (), wide_wave_nodes=wide, lookahead=1)
    assert sol.canonical_sha256() == g["sha256"], (name, wide)
    st0, st1 = base.stats(), automaton.stats()
    for key in ("n_states", "n_edges", "n_search_nodes", "n_fails", "n_leaves", "n_dominance", "n_waves"):
        assert st0[key] == st1[key], (key, wide)


LONG_CHAIN = """var C : [0, 63];
var D : [0, 7];
first C == 0;
first D == 0;
next C == if C eq 63 then 0 else (C + 1);
next D == if C eq 63 then (if D eq 7 then 0 else (D + 1)) else D;
"""


def test_time_limit_slices_the_search_without_changing_it():
    """With a time limit the search kernel comes back to the host every 256 waves (the deadline is checked there);
    a 512-state cycle, one state per wave or two, needs several such slices and must give the same automaton."""
    model = binding.Model(LONG_CHAIN)
    binding.solve(model)                # first solve of a model: pools grow, the kernel comes back for that too
    a0 = binding.solve(model)
    a1 = binding.solve(model, binding.default_options(time_limit_s=1000))
    s0, s1 = binding.Solution(model, a0), binding.Solution(model, a1)
    assert (s0.n_states, s0.n_edges) == (513, 513)      # the start state and the 512-cycle (the reference prints 513 too)
    assert s0.canonical_text() == s1.canonical_text()
    assert a0.stats()["n_waves"] > 256
    assert a1.stats()["n_kernel_launches"] > a0.stats()["n_kernel_launches"]
    import _oracle
    oracle_automaton, _ = _oracle.solve(model, 20.0)
    assert binding.Solution(model, oracle_automaton).canonical_text() == s0.canonical_text()


def test_session_api_single_rank_matches_solve():
    """create / expand / [resolve] / ingest / finish / assemble / trim by hand == stcsp_gpu_solve."""
    g = GOLDENS["probe_first_capture"]
    model = binding.Model(g["model"])
    s = binding.Session(model, None, 0, 1)
    frontier, waves = 1, 0
    while frontier > 0:
        n_leaves, n_pending = s.expand()
        if n_pending:
            s.resolve(s.pending(n_pending))
        frontier = s.ingest(None, 0)
        waves += 1
    part = s.finish()
    s.close()
    merged = binding.assemble([binding.part_to_arrays(part)], trim=True)
    sol = binding.Solution(model, merged)
    assert sol.canonical_sha256() == g["sha256"]
    assert waves > 2


@pytest.mark.parametrize("balls,height", [(3, 5), (5, 7), (6, 7), (7, 7), (8, 8)])
def test_juggling_nosym_closed_form(balls, height):
    """juggling_b{B}_f{F}_nosym has 1 + F!/(F-B)! states, none of them failed; with B == F every
    non-root state has exactly two out-edges; and the whole automaton equals the semantic oracle's (labels included)."""
    model, automaton, sol = run_gpu(instances.juggling(balls, height, sym=False))
    n = 1 + math.factorial(height) // math.factorial(height - balls)
    assert sol.n_states == n
    assert sol.root_valid
    if balls == height:
        assert sol.n_edges == 2 * (n - 1)
    key = "semantic_juggling_b%d_f%d_nosym" % (balls, height)
    want = GOLDENS[key] if key in GOLDENS else None
    if want is None:
        r = _oracle.semantic(model)
        want = {"sha256": r["sha256"], "states": r["n_states"], "edges": r["n_edges"]}
    assert (sol.n_states, sol.n_edges) == (want["states"], want["edges"])
    assert sol.canonical_sha256_streamed() == want["sha256"]


def test_generated_benchmark_instances_are_pinned():
    """BASELINE.json config 5: every synthetic instance the benchmark reports has a canonical-hash golden in CASES."""
    for name in ("juggling_b8_f8_nosym", "partialorder_16", "partialorder_18"):
        assert "semantic_" + name in CASES


def test_trim_is_idempotent_and_edges_grouped():
    model, automaton, sol = run_gpu(GOLDENS["probe_dead_branch"]["model"])
    import ctypes as C
    import numpy as np
    before = (automaton.c.n_edges, automaton.edge_src.copy())
    binding.lib().stcsp_automaton_trim(C.byref(automaton.c))
    assert automaton.c.n_edges == before[0]
    assert np.all(np.diff(before[1]) >= 0)


def test_release_caches_then_solve_again():
    g = GOLDENS["juggling_b4_f5_nosym"]
    _, _, s1 = run_gpu(golden_text(g))
    binding.release_caches()
    _, _, s2 = run_gpu(golden_text(g))
    assert s1.canonical_sha256() == s2.canonical_sha256() == g["sha256"]


def test_warmup_then_results_come_from_the_pinned_arena():
    """stcsp_gpu_warmup: one-time set-up (context, modules, arenas, pinned host memory for results); solves after it give the
    same automata, including results larger than what is left of the arena (those are pinned the ordinary way)."""
    binding.warmup(-1, 8 << 20)
    binding.warmup(-1, 8 << 20)           # idempotent
    for name in ("partialorder_11", "partialorder_13", "digitinvader6", "juggling_b5_f6_nosym"):
        binding.release_caches()
        _, _, sol = run_gpu(instances.by_name(name))
        assert sol.canonical_sha256() == GOLDENS[name]["sha256"], name


def test_frontier_limit_reports_capacity():
    model = binding.Model(instances.by_name("partialorder_10"))
    with pytest.raises(binding.StcspError) as e:
        binding.solve(model, binding.default_options(max_frontier_nodes=16))
    assert e.value.status == binding.ERR_CAPACITY
    # and the library is still usable afterwards
    _, _, sol = run_gpu(instances.by_name("partialorder_10"))
    assert sol.canonical_sha256() == GOLDENS["partialorder_10"]["sha256"]


def test_no_device_option_errors_loudly():
    model = binding.Model(instances.by_name("juggling_b4_f4"))
    with pytest.raises(binding.StcspError) as e:
        binding.solve(model, binding.default_options(device=99))
    assert e.value.status == binding.ERR_CUDA


CLI_CASES = [("juggling_b4_f4", ""), ("juggling_b5_f6_nosym", ""), ("digitinvader3", "-a"), ("digitinvader3", "-z"),
             ("probe_adversarial_win", "-z"), ("probe_adversarial", "-a"), ("probe_k1", "-k1"), ("probe_until_two", ""),
             ("partialorder_12", "")]


@pytest.mark.parametrize("name,flag", CLI_CASES)
def test_cli_on_gpu_writes_the_reference_output(name, flag, tmp_path):
    """bin/stcsp (the C++ host: front end, flags, C-ABI solve, post-processing, DOT writer) as a user runs it:
    `stcsp -s [flag] file.csp` -> the reference's stat line on stdout, solutions.dot in the working directory; the DOT
    re-parsed by the independent Python canonicaliser gives the reference's golden hash."""
    import os
    import subprocess
    from conftest import ROOT
    from stcsp_solver_b200 import canonical
    key = name + ("_" + flag.strip("-") if flag else "")
    g = GOLDENS[key]
    p = tmp_path / (name + ".csp")
    p.write_text(golden_text(g))
    argv = [os.path.join(ROOT, "bin", "stcsp"), "-s"] + ([flag] if flag else []) + ["--sha256", str(p)]
    r = subprocess.run(argv, cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stat = r.stdout.strip().split("\n")[-1].split("\t")
    assert len(stat) == 8                                                   # src/solveralgorithm.cpp:1000-1001
    assert (int(stat[1]), int(stat[2])) == (g["stat"]["vars"], g["stat"]["cons"])
    if flag == "-a":
        assert r.stdout.startswith(g["stdout"].split(";")[0] + "; ")
    if flag == "-z":
        assert r.stdout.startswith(g["stdout"].split("\n")[0] + "\n")
    dot = (tmp_path / "solutions.dot").read_text()
    a = canonical.parse_dot(dot)
    assert canonical.canonical_sha256(a) == g["sha256"]
    assert canonical.counts(a) == (g["states"], g["edges"])
    assert ("canonical sha256 " + g["sha256"]) in r.stderr                  # the library's streamed hash agrees
    assert dot.split("\n")[1] == g["header_vars"] and dot.split("\n")[2].rstrip() == g["header_sig"].rstrip()
