"""ctypes access to oracle/liboracle.so (TEST INFRASTRUCTURE: the CPU restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os

from stcsp_solver_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")


class OracleStats(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_fails", C.c_int64), ("num_dominance", C.c_int64),
                ("gac_calls", C.c_int64), ("validates", C.c_int64), ("node_visits", C.c_int64),
                ("revisions", C.c_int64), ("leaves", C.c_int64), ("splits", C.c_int64), ("max_depth", C.c_int64),
                ("constraint_sets", C.c_int64), ("solve_s", C.c_double), ("timed_out", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(ORACLE_PATH)
        L.stcsp_oracle_solve.argtypes = [C.POINTER(binding.Problem), C.c_double, C.POINTER(binding.AutomatonC),
                                         C.POINTER(OracleStats)]
        L.stcsp_oracle_automaton_free.argtypes = [C.POINTER(binding.AutomatonC)]
        L.stcsp_oracle_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def solve(model: binding.Model, time_limit_s: float = 0.0):
    """Run the oracle.  Returns (Automaton or None on timeout, stats dict)."""
    out = binding.AutomatonC()
    st = OracleStats()
    rc = lib().stcsp_oracle_solve(model.problem, float(time_limit_s), C.byref(out), C.byref(st))
    stats = {k: getattr(st, k) for k, _ in OracleStats._fields_}
    if rc == binding.ERR_TIMEOUT:
        return None, stats
    if rc != 0:
        raise RuntimeError("oracle: %s" % lib().stcsp_oracle_last_error().decode())
    return binding.Automaton(out, lib().stcsp_oracle_automaton_free), stats


def sample(model: binding.Model, seconds: float):
    """Bounded sample of the oracle's search: run for `seconds`, report the counters (no automaton)."""
    st = OracleStats()
    rc = lib().stcsp_oracle_solve(model.problem, float(seconds), None, C.byref(st))
    stats = {k: getattr(st, k) for k, _ in OracleStats._fields_}
    if rc not in (0, binding.ERR_TIMEOUT):
        raise RuntimeError("oracle: %s" % lib().stcsp_oracle_last_error().decode())
    return stats


# ---- the independent semantic oracle (oracle/semantic_oracle.cpp): BFS over signatures from the definition ----
SEMANTIC_PATH = os.path.join(ROOT, "oracle", "libsemantic.so")


class SemanticResult(C.Structure):
    _fields_ = [("n_states", C.c_int64), ("n_edges", C.c_int64), ("n_table_states", C.c_int64),
                ("n_raw_edges", C.c_int64), ("n_point_nodes", C.c_int64), ("root_valid", C.c_int32),
                ("adver1", C.c_int32), ("adver2", C.c_int32), ("n_constraint_sets", C.c_int32),
                ("seconds", C.c_double), ("sha256", C.c_char * 72)]


_sem = None


def semantic_lib():
    global _sem
    if _sem is None:
        L = C.CDLL(SEMANTIC_PATH)
        L.stcsp_semantic_solve.argtypes = [C.POINTER(binding.Problem), C.c_int, C.c_int, C.POINTER(SemanticResult),
                                           C.POINTER(C.c_void_p)]
        L.stcsp_semantic_last_error.restype = C.c_char_p
        L.stcsp_semantic_free_text.argtypes = [C.c_void_p]
        _sem = L
    return _sem


def semantic(model: binding.Model, adversarial1: bool = False, adversarial2: bool = False, want_text: bool = False):
    """Canonical automaton from the definition.  Returns a dict (states, edges, sha256, text or None, ...)."""
    res = SemanticResult()
    text = C.c_void_p()
    rc = semantic_lib().stcsp_semantic_solve(model.problem, int(adversarial1), int(adversarial2), C.byref(res),
                                             C.byref(text) if want_text else None)
    if rc != 0:
        raise RuntimeError("semantic oracle: %s" % semantic_lib().stcsp_semantic_last_error().decode())
    out = {k: getattr(res, k) for k, _ in SemanticResult._fields_}
    out["sha256"] = res.sha256.decode()
    out["text"] = None
    if want_text:
        out["text"] = C.string_at(text).decode()
        semantic_lib().stcsp_semantic_free_text(text)
    return out
