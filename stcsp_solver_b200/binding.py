"""ctypes mirror of include/stcsp_b200.h and include/stcsp_host.h.

This is the reference-facing call path of the package: text -> ``Model`` (front end + normaliser,
host) -> ``solve`` (CUDA, through the C ABI only) -> ``Automaton`` -> ``postprocess`` (host).
There is no CPU solver here: ``solve`` raises when the library reports ``STCSP_ERR_CUDA``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import canonical

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstcsp_b200.so")

STCSP_OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_CAPACITY, ERR_TIMEOUT, ERR_PARSE = range(7)
_STATUS = ["ok", "invalid", "unsupported", "cuda", "capacity", "timeout", "parse"]


class Tok(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32)]


class Problem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("prefix_k", C.c_int32), ("n_vars", C.c_int32),
        ("var_lb", C.POINTER(C.c_int32)), ("var_ub", C.POINTER(C.c_int32)),
        ("var_names", C.POINTER(C.c_char_p)),
        ("n_arrays", C.c_int32), ("arr_offsets", C.POINTER(C.c_int32)), ("arr_values", C.POINTER(C.c_int32)),
        ("n_constraints", C.c_int32), ("con_offsets", C.POINTER(C.c_int32)), ("con_tokens", C.POINTER(Tok)),
    ]


class Options(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("use_current_device", C.c_int32), ("time_limit_s", C.c_int32),
        ("verbosity", C.c_int32), ("enum_limit_now", C.c_int64), ("enum_limit_ahead", C.c_int64),
        ("max_frontier_nodes", C.c_int64), ("max_states", C.c_int64), ("max_edges", C.c_int64),
        ("expand_mode", C.c_int32), ("profile_kernels", C.c_int32), ("no_trim", C.c_int32),
        ("lookahead", C.c_int32), ("wide_wave_nodes", C.c_int32), ("single_branch", C.c_int32),
        ("shard_mode", C.c_int32), ("adversarial", C.c_int32),
    ]


class ExchangeStats(C.Structure):
    """stcsp_exchange_stats_t"""
    _fields_ = [("sharded", C.c_int32), ("pad", C.c_int32), ("waves", C.c_int64), ("exchanges", C.c_int64),
                ("records", C.c_int64), ("bytes_pulled", C.c_int64), ("exchange_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "pad"}


class AutomatonC(C.Structure):
    _fields_ = [
        ("n_vars", C.c_int32), ("n_sig_vars", C.c_int32), ("n_until", C.c_int32), ("n_until_vars", C.c_int32),
        ("sig_len", C.c_int32), ("sig_vars", C.POINTER(C.c_int32)), ("root_final", C.c_int32),
        ("n_constraint_sets", C.c_int32), ("n_states", C.c_int64),
        ("state_sig", C.POINTER(C.c_int32)), ("state_cset", C.POINTER(C.c_int32)),
        ("state_failed", C.POINTER(C.c_uint8)), ("n_edges", C.c_int64),
        ("edge_src", C.POINTER(C.c_int32)), ("edge_dst", C.POINTER(C.c_int32)), ("edge_label", C.POINTER(C.c_int32)),
        ("n_search_nodes", C.c_int64), ("n_fails", C.c_int64), ("n_leaves", C.c_int64), ("n_dominance", C.c_int64),
        ("n_waves", C.c_int64), ("n_tuples", C.c_int64), ("n_revisions", C.c_int64),
        ("n_kernel_launches", C.c_int64),
        ("solve_ms", C.c_double), ("wall_ms", C.c_double), ("expand_ms", C.c_double),
        ("n_expand_launches", C.c_int64),
        ("algorithmic_bytes", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("state_final", C.POINTER(C.c_uint8)), ("state_valid", C.POINTER(C.c_uint8)), ("edge_alive", C.POINTER(C.c_uint8)),
        ("post_applied", C.c_int32), ("adver1", C.c_int32), ("adver2", C.c_int32), ("pad0", C.c_int32),
        ("impl", C.c_void_p),
    ]


class SolutionC(C.Structure):
    _fields_ = [
        ("root_valid", C.c_int32), ("adver1", C.c_int32), ("adver2", C.c_int32),
        ("n_states", C.c_int64), ("n_edges", C.c_int64), ("n_table_states", C.c_int64),
        ("n_vars", C.c_int32), ("sig_len", C.c_int32),
        ("state_cset", C.POINTER(C.c_int32)), ("state_final", C.POINTER(C.c_uint8)),
        ("state_sig", C.POINTER(C.c_int32)),
        ("edge_src", C.POINTER(C.c_int32)), ("edge_dst", C.POINTER(C.c_int32)), ("edge_label", C.POINTER(C.c_int32)),
        ("impl", C.c_void_p),
    ]


class StcspError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("stcsp %s error: %s" % (_STATUS[status] if 0 <= status < len(_STATUS) else status, message))
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Load libstcsp_b200.so (built in-tree by ``__graft_entry__.build()`` / ``make``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python __graft_entry__.py build` (needs nvcc); "
                              "there is no pure-Python or CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.stcsp_last_error.restype = C.c_char_p
        L.stcsp_model_parse_text.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.stcsp_model_parse_file.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.stcsp_model_free.argtypes = [C.c_void_p]
        L.stcsp_model_problem.argtypes = [C.c_void_p]
        L.stcsp_model_problem.restype = C.POINTER(Problem)
        L.stcsp_model_dump.argtypes = [C.c_void_p]
        L.stcsp_model_dump.restype = C.c_void_p
        L.stcsp_string_free.argtypes = [C.c_void_p]
        L.stcsp_gpu_solve.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(AutomatonC)]
        L.stcsp_automaton_free.argtypes = [C.POINTER(AutomatonC)]
        L.stcsp_automaton_trim.argtypes = [C.POINTER(AutomatonC)]
        L.stcsp_postprocess.argtypes = [C.POINTER(Problem), C.POINTER(AutomatonC), C.c_int32, C.c_int32,
                                        C.POINTER(SolutionC)]
        L.stcsp_solution_free.argtypes = [C.POINTER(SolutionC)]
        L.stcsp_solution_dot.argtypes = [C.POINTER(Problem), C.POINTER(SolutionC)]
        L.stcsp_solution_dot.restype = C.c_void_p
        L.stcsp_solution_canonical.argtypes = [C.POINTER(Problem), C.POINTER(SolutionC)]
        L.stcsp_solution_canonical.restype = C.c_void_p
        L.stcsp_solution_write_dot.argtypes = [C.POINTER(Problem), C.POINTER(SolutionC), C.c_char_p]
        L.stcsp_solution_write_canonical.argtypes = [C.POINTER(Problem), C.POINTER(SolutionC), C.c_char_p]
        L.stcsp_solution_canonical_sha256.argtypes = [C.POINTER(Problem), C.POINTER(SolutionC), C.c_char_p]
        L.stcsp_gpu_device_count.restype = C.c_int
        L.stcsp_gpu_release_caches.restype = None
        L.stcsp_gpu_warmup.argtypes = [C.c_int32, C.c_int64]
        L.stcsp_gpu_warmup.restype = C.c_int
        L.stcsp_session_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.c_int32, C.c_int32,
                                           C.POINTER(C.c_void_p)]
        L.stcsp_session_destroy.argtypes = [C.c_void_p]
        L.stcsp_session_record_words.argtypes = [C.c_void_p]
        L.stcsp_session_record_words.restype = C.c_int32
        L.stcsp_session_request_words.argtypes = [C.c_void_p]
        L.stcsp_session_request_words.restype = C.c_int32
        L.stcsp_session_key_words.argtypes = [C.c_void_p]
        L.stcsp_session_key_words.restype = C.c_int32
        L.stcsp_session_expand.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.stcsp_session_pending.argtypes = [C.c_void_p, C.c_void_p]
        L.stcsp_session_resolve.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.stcsp_session_outbox.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.stcsp_session_ingest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.stcsp_session_finish.argtypes = [C.c_void_p, C.POINTER(AutomatonC)]
        L.stcsp_session_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.stcsp_session_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.stcsp_session_finish_merged.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32,
                                                  C.POINTER(AutomatonC)]
        L.stcsp_automaton_assemble.argtypes = [C.POINTER(AutomatonC), C.c_int32, C.POINTER(AutomatonC)]
        L.stcsp_group_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.stcsp_group_destroy.argtypes = [C.c_void_p]
        L.stcsp_group_share_bytes.restype = C.c_int64
        L.stcsp_group_share.argtypes = [C.c_void_p, C.c_void_p]
        L.stcsp_group_attach.argtypes = [C.c_void_p, C.c_void_p]
        L.stcsp_group_solve.argtypes = [C.c_void_p, C.POINTER(Problem), C.POINTER(Options), C.POINTER(AutomatonC),
                                        C.POINTER(ExchangeStats)]
        L.stcsp_gpu_solve_multi.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.c_int32, C.POINTER(C.c_int32),
                                            C.POINTER(AutomatonC), C.POINTER(ExchangeStats)]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != STCSP_OK:
        raise StcspError(rc, (lib().stcsp_last_error() or b"").decode())


def _take_string(ptr) -> str:
    s = C.string_at(ptr).decode()
    lib().stcsp_string_free(ptr)
    return s


class Model:
    """A parsed and normalised .csp model (host side; reference solverParse, src/solver.cpp:138-159)."""

    def __init__(self, text: str, prefix_k: int = 2):
        self._h = C.c_void_p()
        _check(lib().stcsp_model_parse_text(text.encode(), prefix_k, C.byref(self._h)))
        self.problem = lib().stcsp_model_problem(self._h)

    @classmethod
    def from_file(cls, path: str, prefix_k: int = 2) -> "Model":
        with open(path) as f:
            return cls(f.read(), prefix_k)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().stcsp_model_free(self._h)
                self._h = None
        except Exception:       # interpreter shutdown: module globals are already gone
            pass

    @property
    def var_names(self) -> List[str]:
        p = self.problem.contents
        return [p.var_names[i].decode() for i in range(p.n_vars)]

    @property
    def n_vars(self) -> int:
        return self.problem.contents.n_vars

    @property
    def n_constraints(self) -> int:
        return self.problem.contents.n_constraints

    def dump(self) -> str:
        return _take_string(lib().stcsp_model_dump(self._h))


def _np(ptr, n, dtype, copy=True):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(n,))
    return a.astype(dtype, copy=True) if copy else a


class Automaton:
    """Owner of one ``stcsp_automaton_t`` plus numpy copies of its arrays."""

    def __init__(self, c_struct: AutomatonC, free_fn, copy: bool = False):
        """`copy=False`: the numpy arrays are views into the library-owned automaton (valid while this object lives)."""
        self.c = c_struct
        self._free = free_fn
        a = c_struct
        self.n_vars, self.sig_len = a.n_vars, a.sig_len
        self.n_states, self.n_edges = a.n_states, a.n_edges
        self.sig_vars = _np(a.sig_vars, a.n_sig_vars, np.int32, copy)
        self.state_sig = _np(a.state_sig, a.n_states * a.sig_len, np.int32, copy).reshape(a.n_states, a.sig_len)
        self.state_cset = _np(a.state_cset, a.n_states, np.int32, copy)
        self.state_failed = _np(a.state_failed, a.n_states, np.uint8, copy)
        self.edge_src = _np(a.edge_src, a.n_edges, np.int32, copy)
        self.edge_dst = _np(a.edge_dst, a.n_edges, np.int32, copy)
        self.edge_label = _np(a.edge_label, a.n_edges * a.n_vars, np.int32, copy).reshape(a.n_edges, a.n_vars)
        # post-processing done on the device (None when it did not run)
        self.post_applied = a.post_applied
        self.state_final = _np(a.state_final, a.n_states, np.uint8, copy) if a.state_final else None
        self.state_valid = _np(a.state_valid, a.n_states, np.uint8, copy) if a.state_valid else None
        self.edge_alive = _np(a.edge_alive, a.n_edges, np.uint8, copy) if a.edge_alive else None

    def stats(self) -> dict:
        a = self.c
        return {k: getattr(a, k) for k in (
            "n_states", "n_edges", "n_constraint_sets", "n_search_nodes", "n_fails", "n_leaves", "n_dominance",
            "n_waves", "n_tuples", "n_revisions", "n_kernel_launches", "n_expand_launches", "solve_ms", "wall_ms",
            "expand_ms", "algorithmic_bytes",
            "h2d_bytes", "d2h_bytes")}

    def __del__(self):
        try:
            if getattr(self, "_free", None):
                self._free(C.byref(self.c))
                self._free = None
        except Exception:
            pass


class Solution:
    """Post-processed automaton (what the reference prints with -s)."""

    def __init__(self, model: Model, automaton: Automaton, adversarial1: bool = False, adversarial2: bool = False):
        self.model = model
        self.c = SolutionC()
        _check(lib().stcsp_postprocess(model.problem, C.byref(automaton.c), int(adversarial1), int(adversarial2),
                                       C.byref(self.c)))
        self.root_valid = bool(self.c.root_valid)
        self.adver1, self.adver2 = self.c.adver1, self.c.adver2
        self.n_states, self.n_edges = self.c.n_states, self.c.n_edges

    def dot(self) -> str:
        return _take_string(lib().stcsp_solution_dot(self.model.problem, C.byref(self.c)))

    def canonical_text(self) -> str:
        return _take_string(lib().stcsp_solution_canonical(self.model.problem, C.byref(self.c)))

    def canonical_sha256(self) -> str:
        import hashlib
        return hashlib.sha256(self.canonical_text().encode()).hexdigest()

    def canonical_sha256_streamed(self) -> str:
        """SHA-256 of the canonical text computed line by line in the library (no text is materialised)."""
        buf = C.create_string_buffer(65)
        _check(lib().stcsp_solution_canonical_sha256(self.model.problem, C.byref(self.c), buf))
        return buf.value.decode()

    def write_dot(self, path: str) -> None:
        _check(lib().stcsp_solution_write_dot(self.model.problem, C.byref(self.c), path.encode()))

    def to_python(self) -> canonical.Automaton:
        """Re-parse the DOT text with the independent Python canonicaliser."""
        return canonical.parse_dot(self.dot())

    def __del__(self):
        try:
            if getattr(self, "c", None) is not None and self.c.impl:
                lib().stcsp_solution_free(C.byref(self.c))
        except Exception:
            pass


def default_options(**kw) -> Options:
    o = Options()
    o.device = -1
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def solve(model: Model, options: Optional[Options] = None) -> Automaton:
    """Run the search on the GPU (stcsp_gpu_solve).  Raises StcspError(ERR_CUDA) without a device."""
    out = AutomatonC()
    opts = options if options is not None else default_options()
    _check(lib().stcsp_gpu_solve(model.problem, C.byref(opts), C.byref(out)))
    return Automaton(out, lib().stcsp_automaton_free)


def warmup(device: int = -1, pinned_bytes: int = 0) -> None:
    """One-time set-up of a process that will solve more than once (stcsp_gpu_warmup): context, kernel modules, arenas,
    and `pinned_bytes` of pinned host memory for results."""
    _check(lib().stcsp_gpu_warmup(int(device), int(pinned_bytes)))


def release_caches() -> None:
    """Give the library's cached device / pinned memory and resident models back (stcsp_gpu_release_caches)."""
    lib().stcsp_gpu_release_caches()


def solve_text(text: str, flags: Sequence[str] = (), options: Optional[Options] = None):
    """Convenience mirror of the reference CLI: returns (Model, Automaton, Solution)."""
    k = 2
    for f in flags:
        if f.startswith("-k"):
            k = int(f[2:])
    model = Model(text, k)
    automaton = solve(model, options)
    return model, automaton, Solution(model, automaton, "-a" in flags, "-z" in flags)


class Session:
    """Step-wise search on one GPU (one rank of a multi-GPU solve); see include/stcsp_b200.h."""

    def __init__(self, model: Model, options: Optional[Options], rank: int, world_size: int):
        self.model = model
        self._h = C.c_void_p()
        opts = options if options is not None else default_options()
        _check(lib().stcsp_session_create(model.problem, C.byref(opts), rank, world_size, C.byref(self._h)))
        self.rank, self.world_size = rank, world_size
        self.record_words = lib().stcsp_session_record_words(self._h)
        self.request_words = lib().stcsp_session_request_words(self._h)
        self.key_words = lib().stcsp_session_key_words(self._h)

    def close(self):
        try:
            if getattr(self, "_h", None):
                lib().stcsp_session_destroy(self._h)
                self._h = None
        except Exception:
            pass

    __del__ = close

    def expand(self):
        n_leaves, n_pending = C.c_int64(), C.c_int64()
        _check(lib().stcsp_session_expand(self._h, C.byref(n_leaves), C.byref(n_pending)))
        return n_leaves.value, n_pending.value

    def pending(self, n_pending: int) -> np.ndarray:
        out = np.zeros((n_pending, self.request_words), dtype=np.int32)
        if n_pending:
            _check(lib().stcsp_session_pending(self._h, out.ctypes.data))
        return out

    def resolve(self, requests: np.ndarray) -> None:
        requests = np.ascontiguousarray(requests, dtype=np.int32)
        _check(lib().stcsp_session_resolve(self._h, requests.ctypes.data, requests.shape[0]))

    def outbox(self, device_ptr: int, capacity: int) -> np.ndarray:
        counts = (C.c_int64 * self.world_size)()
        _check(lib().stcsp_session_outbox(self._h, device_ptr, capacity, counts))
        return np.array(list(counts), dtype=np.int64)

    def ingest(self, device_ptr: Optional[int], n_records: int) -> int:
        nxt = C.c_int64()
        _check(lib().stcsp_session_ingest(self._h, device_ptr, n_records, C.byref(nxt)))
        return nxt.value

    def finish(self) -> Automaton:
        out = AutomatonC()
        _check(lib().stcsp_session_finish(self._h, C.byref(out)))
        return Automaton(out, lib().stcsp_automaton_free)

    def counts(self):
        """(n_states, n_edges, stats[10]) of this rank's part."""
        ns, ne = C.c_int64(), C.c_int64()
        stats = (C.c_int64 * 10)()
        _check(lib().stcsp_session_counts(self._h, C.byref(ns), C.byref(ne), stats))
        return ns.value, ne.value, np.array(list(stats), dtype=np.int64)

    def export(self, keys_ptr: int, src_ptr: int, dst_ptr: int, label_ptr: int) -> None:
        _check(lib().stcsp_session_export(self._h, keys_ptr, src_ptr, dst_ptr, label_ptr))

    def finish_merged(self, n_states, n_edges, keys_ptr, src_ptr, dst_ptr, label_ptr, extra_stats, trim=True) -> Automaton:
        w = len(n_states)
        ns = (C.c_int64 * w)(*[int(x) for x in n_states])
        ne = (C.c_int64 * w)(*[int(x) for x in n_edges])
        ex = (C.c_int64 * 10)(*[int(x) for x in extra_stats])
        out = AutomatonC()
        _check(lib().stcsp_session_finish_merged(self._h, w, ns, ne, keys_ptr, src_ptr, dst_ptr, label_ptr, ex, int(trim),
                                                 C.byref(out)))
        return Automaton(out, lib().stcsp_automaton_free)


_PART_STATS = ("n_constraint_sets", "n_search_nodes", "n_fails", "n_leaves", "n_dominance", "n_waves", "n_tuples",
               "n_revisions", "n_kernel_launches", "n_expand_launches", "algorithmic_bytes", "h2d_bytes", "d2h_bytes")
_PART_TIMES = ("solve_ms", "wall_ms", "expand_ms")


def part_to_arrays(a: Automaton) -> dict:
    """Plain-numpy form of one rank's part (what travels between ranks)."""
    c = a.c
    head = np.array([c.n_vars, c.n_sig_vars, c.n_until, c.n_until_vars, c.sig_len, c.root_final, c.n_states, c.n_edges]
                    + [getattr(c, k) for k in _PART_STATS], dtype=np.int64)
    times = np.array([getattr(c, k) for k in _PART_TIMES], dtype=np.float64)
    # copies: the dict outlives the automaton whose memory the arrays of `a` are views of
    return {"head": head, "times": times, "sig_vars": np.array(a.sig_vars), "state_sig": np.array(a.state_sig).reshape(-1),
            "state_cset": np.array(a.state_cset), "edge_src": np.array(a.edge_src), "edge_dst": np.array(a.edge_dst),
            "edge_label": np.array(a.edge_label).reshape(-1)}


def assemble(parts: Sequence[dict], trim: bool = True) -> Automaton:
    """stcsp_automaton_assemble (+ trim) over per-rank parts given as numpy arrays (rank order)."""
    arr = (AutomatonC * len(parts))()
    keep = []
    for i, p in enumerate(parts):
        h = [int(x) for x in p["head"]]
        c = arr[i]
        c.n_vars, c.n_sig_vars, c.n_until, c.n_until_vars, c.sig_len, c.root_final, c.n_states, c.n_edges = h[:8]
        for k, v in zip(_PART_STATS, h[8:]):
            setattr(c, k, v)
        for k, v in zip(_PART_TIMES, p["times"]):
            setattr(c, k, float(v))
        for name in ("sig_vars", "state_sig", "state_cset", "edge_src", "edge_dst", "edge_label"):
            a = np.ascontiguousarray(p[name], dtype=np.int32)
            keep.append(a)
            setattr(c, name, a.ctypes.data_as(C.POINTER(C.c_int32)))
        c.state_failed = None
    out = AutomatonC()
    _check(lib().stcsp_automaton_assemble(arr, len(parts), C.byref(out)))
    if trim:
        _check(lib().stcsp_automaton_trim(C.byref(out)))
    return Automaton(out, lib().stcsp_automaton_free)


class Group:
    """Several GPUs on one automaton (stcsp_group_*): one Group per rank.  Form it once, solve many times.

    share() -> bytes for the peers; attach(list of every rank's bytes, rank order); solve(model) is collective and
    returns (Automaton or None, exchange statistics) -- the automaton on rank 0 only."""

    def __init__(self, rank: int, world_size: int, device: int = -1):
        self.rank, self.world = rank, world_size
        self._h = C.c_void_p()
        _check(lib().stcsp_group_create(rank, world_size, device, C.byref(self._h)))

    def share(self) -> bytes:
        buf = C.create_string_buffer(lib().stcsp_group_share_bytes())
        _check(lib().stcsp_group_share(self._h, buf))
        return buf.raw

    def attach(self, blobs: Sequence[bytes]) -> None:
        joined = b"".join(blobs)
        _check(lib().stcsp_group_attach(self._h, C.create_string_buffer(joined, len(joined))))

    def solve(self, model: Model, options: Optional[Options] = None):
        out, xs = AutomatonC(), ExchangeStats()
        opts = options if options is not None else default_options()
        _check(lib().stcsp_group_solve(self._h, model.problem, C.byref(opts), C.byref(out), C.byref(xs)))
        stats = xs.as_dict()
        if self.rank != 0:
            return None, stats
        a = Automaton(out, lib().stcsp_automaton_free)
        a.exchange_stats = stats
        return a, stats

    def close(self):
        if getattr(self, "_h", None):
            lib().stcsp_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_multi(model: Model, n_gpus: int, options: Optional[Options] = None, devices: Optional[Sequence[int]] = None):
    """One process, one host thread per GPU (stcsp_gpu_solve_multi).  Returns (Automaton, exchange statistics)."""
    out, xs = AutomatonC(), ExchangeStats()
    opts = options if options is not None else default_options()
    dev = (C.c_int32 * n_gpus)(*devices) if devices is not None else None
    _check(lib().stcsp_gpu_solve_multi(model.problem, C.byref(opts), n_gpus, dev, C.byref(out), C.byref(xs)))
    a = Automaton(out, lib().stcsp_automaton_free)
    a.exchange_stats = xs.as_dict()
    return a, a.exchange_stats
