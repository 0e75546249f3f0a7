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
