#!/usr/bin/env python
"""Benchmark of the search-and-propagate path (BASELINE.json: search nodes/s + solve time on
juggling_b6_f6_nosym, 1/2/4/8 B200, beside the CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--instance NAME] [--impl b200|reference]

One STEP = one complete solve of the instance (parse/normalise excluded, like the reference's
own solveTime).  Prints ONE JSON line on rank 0.

Unit of work.  The reference's and this library's search trees differ (binary splits and bounds
propagation there, value branching and budgeted domain propagation here), so raw node counts are
not comparable.  `value` therefore counts REFERENCE search nodes -- the number of
generalisedArcConsistent calls the reference needs for this instance (pinned
by tests/test_oracle.py) -- per second of time-to-automaton: value = ref_nodes / solve_time.  The
ratio of two arms is then exactly the ratio of their solve times.  The library's own node rate is
reported as `own_search_nodes_per_s`.

  value     device time (CUDA events on the library's stream around the whole wave loop), model tables
            already in HBM (steady state: the compiled model of the previous solve is resident)
  e2e       wall time of the C-ABI call stcsp_gpu_solve with HOST buffers in and out: H2D of the
            model, the search, D2H of states and edges into host memory
  parity    the automaton of the LAST timed step, post-processed and canonicalised, against the golden
            SHA-256 (reference output, or the pinned semantic oracle for generated instances); no value
            is printed when it differs
  roofline  search_kernel (the dominant kernel): SURVEY.md 8(d) algorithmic bytes per launch / CUDA-event
            duration of the launches, against MEASURED_PEAKS.json; `int_ops` = thread instructions/s (ncu
            count of profiles/counters.json / live duration) against 148 SMs x 128 lanes x clock
  cold      the same call in a FRESH process: first solve of the model (nothing cached), and the
            command-line tool bin/stcsp against oracle/_ref/stcsp_ref, wall clock both
  cpu_baseline  ONE complete solve by the reference itself (oracle/_ref/stcsp_ref on the .csp text, its
            own solveTime and the wall clock), one host core -- the reference is single-threaded; the
            bounded sample of the oracle port is kept as a second figure

--impl reference: every step is one COMPLETE solve by oracle/_ref/stcsp_ref; the steps run as
concurrent processes on the box's cores (the only way the single-threaded reference can use them),
value = the aggregate of those concurrent solves (the CPU box's best throughput), solve_time_s and
single_core_value = from the FASTEST single solve seen (solo run included).

N > 1 GPUs: the headline instance is too small to shard, so `value` is the aggregate of N independent
replicas (one solve per GPU and step, "scaling": "weak"); `sharded` holds the instances that do shard
(partialorder_16/18/20, juggling_b8_f8_nosym: forced sharding over the N GPUs against one GPU).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import resource
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "stcsp_ref")
CLI_BIN = os.path.join(ROOT, "bin", "stcsp")

# Reference work per instance: (generalisedArcConsistent calls, validate() calls), SURVEY.md Appendix G.2,
# measured with the counter build of the unmodified reference; the oracle port reproduces them exactly.
REF_WORK = {
    "juggling_b4_f4": (5, 4950), "juggling_b4_f4_nosym": (133, 128237), "juggling_b4_f5": (327, 741752),
    "juggling_b4_f5_nosym": (805, 1970087), "juggling_b4_f6": (1379, 8356554), "juggling_b4_f6_nosym": (3111, 20469178),
    "juggling_b5_f5": (6, 47727), "juggling_b5_f5_nosym": (677, 5690538), "juggling_b5_f6": (1939, 56328336),
    "juggling_b5_f6_nosym": (4971, 142514383), "juggling_b6_f6": (7, 973373), "juggling_b6_f6_nosym": (4119, 723505885),
    "partialorder_10": (55636, 9155081), "partialorder_11": (126930, 23092746), "partialorder_12": (277708, 55337427),
    "partialorder_13": (612810, 133335094), "partialorder_14": (1322436, 312127024),
    "digitinvader1": (129, 115517), "digitinvader2": (845, 1204444), "digitinvader3": (3535, 8497696),
    "digitinvader4": (11349, 46612054), "digitinvader5": (30503, 209855925), "digitinvader6": (72085, 803218388),
    "digitinvader7": (154455, 2710535457), "digitinvader8": (306323, 8233605443), "digitinvader9": (570589, 22660973789),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def golden_for(name):
    """Golden record of an instance: reference output if the reference can finish it, else the pinned semantic oracle's."""
    for fn in (name + ".json", "semantic_" + name + ".json"):
        p = os.path.join(ROOT, "tests", "golden", fn)
        if os.path.exists(p):
            with open(p) as f:
                g = json.load(f)
            g["golden_file"] = "tests/golden/" + fn
            return g
    return None


def parity_of(binding, model, automaton, name):
    """Canonical SHA-256 of the post-processed automaton (streamed, no text is built) against the golden."""
    g = golden_for(name)
    sol = binding.Solution(model, automaton)
    sha = sol.canonical_sha256_streamed()
    out = {"sha256": sha, "states": int(sol.n_states), "edges": int(sol.n_edges), "golden": None, "sha256_ok": None}
    if g is not None and "sha256" in g:
        out["golden"] = g["golden_file"] + (" (semantic oracle)" if g.get("source") == "semantic_oracle" else " (reference)")
        out["sha256_ok"] = bool(sha == g["sha256"] and sol.n_states == g["states"] and sol.n_edges == g["edges"])
    return out


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons while the timed region runs (NVML, 2 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        self.nv = self.handle = None
        try:                                # NVML is initialised here, not in the thread: the timed region is short
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.handle = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        except Exception as e:              # NVML missing: report it instead of inventing numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def run(self):
        nv, h = self.nv, self.handle
        if nv is None:
            return
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        try:
            while True:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                if self.stop_flag:
                    break
                time.sleep(0.002)
        except Exception as e:
            self.reasons.add("nvml_error:%s" % type(e).__name__)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arms
def _unlimited_stack():
    # the reference recurses once per search node along a path of the automaton (src/stcsp.y:184-191 lifts the limit too)
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))


def run_stcsp_ref(text, name, time_limit_s=3600):
    """One complete solve by the reference binary.  Returns its own solveTime (CPU s, times()) and the wall clock."""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, name + ".csp")
        with open(path, "w") as f:
            f.write(text)
        t0 = time.perf_counter()
        p = subprocess.run([REF_BIN, "-m%d" % time_limit_s, path], cwd=d, capture_output=True, text=True,
                           preexec_fn=_unlimited_stack)
        wall = time.perf_counter() - t0
    stat = [ln for ln in p.stdout.split("\n") if ln.count("\t") == 7]
    if p.returncode != 0 or not stat:
        return None
    fld = stat[-1].split("\t")
    return {"solve_s": float(fld[6]), "process_s": float(fld[7]), "wall_s": wall, "states_incl_failed": int(fld[4]),
            "dominance": int(fld[3]), "fails": int(fld[5])}


def port_sample(model, name, seconds):
    """Bounded sample of the oracle PORT (oracle/stcsp_oracle.cpp) on one host core, projected to a full solve from the
    constraint evaluations done (evaluations/s is stationary, nodes/s is not).  A projection, labelled as such."""
    import _oracle
    st = _oracle.sample(model, seconds)
    nodes, validates = REF_WORK.get(name, (None, None))
    if not st["timed_out"]:
        return {"kind": "port", "solve_time_s": st["solve_s"], "projected": False, "sample": "complete solve (%.2f s)" % st["solve_s"],
                "nodes": st["gac_calls"]}
    if not validates:
        return None
    solve_s = st["solve_s"] * validates / max(st["validates"], 1)
    return {"kind": "port", "solve_time_s": solve_s, "projected": True, "nodes": nodes,
            "sample": "first %.1f s of the DFS = %d of %d constraint evaluations; PROJECTED to %.1f s"
                      % (st["solve_s"], st["validates"], validates, solve_s)}


def reference_feasible(name):
    """The reference finishes this instance in bounded time here (golden wall_s from the same binary)."""
    g = golden_for(name)
    return os.path.exists(REF_BIN) and g is not None and g.get("source") != "semantic_oracle" and g.get("wall_s", 1e9) <= 200


def run_reference_arm(args, name, text):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref_nodes = REF_WORK.get(name, (None, None))[0]
    cores = os.cpu_count() or 1
    unit = "reference search nodes/s"
    base = {"impl": "reference", "metric": "search_nodes_per_s", "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": name, "unit_of_work": "reference generalisedArcConsistent calls (%s per solve)" % ref_nodes}}
    if reference_feasible(name) and ref_nodes:
        # the real reference: one solo solve first (nothing else running), then W + K complete solves as concurrent
        # single-threaded processes, at most one per core
        solo = run_stcsp_ref(text, name)
        if solo is None:
            raise SystemExit("stcsp_ref failed on %s" % name)
        conc = max(1, min(cores - 1 if cores > 1 else 1, args.steps + args.warmup))
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(conc) as ex:
            runs = list(ex.map(lambda _: run_stcsp_ref(text, name), range(args.steps + args.warmup)))
        region = time.perf_counter() - t0
        runs = [r for r in runs if r is not None]
        timed = runs[args.warmup:] if len(runs) > args.warmup else runs
        walls = sorted(r["wall_s"] for r in timed)
        best = min([solo["wall_s"]] + walls)
        # The metric is a throughput.  The reference is single-threaded, so the only way it can use the box's cores is one
        # solve per core: value = what all those concurrent solves deliver together (the best the CPU box can do), and
        # solve_time_s = the fastest single solve (the latency a user of the reference sees).
        value = ref_nodes * len(runs) / region
        line = dict(base)
        line.update({
            "value": value, "ms_per_step": region / max(len(runs), 1) * 1e3, "solve_time_s": best,
            "single_core_value": ref_nodes / best,
            "cpu_baseline": {"value": value, "unit": unit, "cores": conc, "kind": "reference", "cpu_model": cpu_model(),
                             "host_cores": cores,
                             "sample": "oracle/_ref/stcsp_ref (the unmodified reference, single-threaded) on the .csp text: 1 solo + %d "
                                       "complete solves, %d at a time on %d cores; value = aggregate of the concurrent solves, "
                                       "single_core_value = from the FASTEST single solve" % (len(runs), conc, cores),
                             "single_core_value": ref_nodes / best,
                             "solo": solo, "concurrent_wall_s": {"min": walls[0], "median": walls[len(walls) // 2], "max": walls[-1]},
                             "own_solveTime_s": {"min": min(r["solve_s"] for r in timed), "max": max(r["solve_s"] for r in timed)},
                             "aggregate_nodes_per_s_all_cores": ref_nodes * len(runs) / region},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        print(json.dumps(line), flush=True)
        return
    # instances the reference cannot finish in minutes: bounded samples of the port, PROJECTED (say so)
    from stcsp_solver_b200 import binding
    model = binding.Model(text)
    per_step = max(2.0, min(10.0, 150.0 / max(args.steps + args.warmup, 1)))
    outs = [port_sample(model, name, per_step) for _ in range(args.steps + args.warmup)][args.warmup:]
    outs = [o for o in outs if o]
    if not outs:
        print(json.dumps({"impl": "reference", "unavailable": "no reference work count for %s" % name}), flush=True)
        return
    best = min(o["solve_time_s"] for o in outs)
    nodes = outs[0]["nodes"]
    line = dict(base)
    line.update({"value": nodes / best, "ms_per_step": best * 1e3, "solve_time_s": best,
                 "cpu_baseline": {"value": nodes / best, "unit": unit, "cores": 1, "kind": "port", "cpu_model": cpu_model(),
                                  "sample": outs[0]["sample"]},
                 "e2e": {"value": nodes / best, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- cold path
COLD_SNIPPET = r"""
import json, sys, time
sys.path.insert(0, %(root)r)
from stcsp_solver_b200 import binding, instances
import ctypes
name = %(name)r
model = binding.Model(instances.by_name(name))
t0 = time.perf_counter()
binding.lib().stcsp_gpu_device_count()
try:
    ctypes.CDLL("libcudart.so").cudaFree(0)          # context creation, reported separately
except OSError:
    pass
ctx = time.perf_counter() - t0
out = {"context_ms": ctx * 1e3, "solves": []}
if %(api)r:        # the library's own one-time set-up call: context, modules, arenas, pinned host memory for results
    t0 = time.perf_counter()
    binding.warmup(-1, 512 << 20)
    out["warmup_call_ms"] = (time.perf_counter() - t0) * 1e3
if %(warm)r:       # the PROCESS is warm (streams, arenas, kernel modules: one-time set-up of the library), the MODEL is cold
    other = binding.Model(instances.by_name(%(warm)r))
    binding.solve(other)
for i in range(3):
    t0 = time.perf_counter()
    a = binding.solve(model, binding.default_options(verbosity=1 if i == 0 else 0))      # the first solve reports its phases on stderr
    w = time.perf_counter() - t0
    out["solves"].append({"e2e_ms": w * 1e3, "call_ms": a.c.wall_ms, "device_ms": a.c.solve_ms, "launches": a.c.n_kernel_launches})
print(json.dumps(out))
"""


def cold_numbers(name, text, ref_wall_s):
    """First solve of the model in a fresh process, and the command-line tool against the reference's, wall clock."""
    out = {}
    try:
        p = subprocess.run([sys.executable, "-c", COLD_SNIPPET % {"root": ROOT, "name": name, "warm": "", "api": False}], capture_output=True,
                           text=True, timeout=600)
        j = json.loads(p.stdout.strip().split("\n")[-1])
        phases = [ln for ln in p.stderr.split("\n") if ln.startswith("[stcsp] wall:")]
        out = {"context_ms": j["context_ms"], "first_solve_phases": phases[0][len("[stcsp] wall: "):] if phases else None, "first_solve_e2e_ms": j["solves"][0]["e2e_ms"],
               "first_solve_device_ms": j["solves"][0]["device_ms"], "first_solve_launches": j["solves"][0]["launches"],
               "second_solve_e2e_ms": j["solves"][1]["e2e_ms"], "third_solve_e2e_ms": j["solves"][2]["e2e_ms"],
               "first_over_third": j["solves"][0]["e2e_ms"] / max(j["solves"][2]["e2e_ms"], 1e-9)}
        # the same with the library's one-time set-up out of the way: another (small) model is solved first
        warm = "juggling_b4_f5_nosym" if name != "juggling_b4_f5_nosym" else "juggling_b4_f4"
        p = subprocess.run([sys.executable, "-c", COLD_SNIPPET % {"root": ROOT, "name": name, "warm": warm, "api": False}], capture_output=True,
                           text=True, timeout=600)
        j = json.loads(p.stdout.strip().split("\n")[-1])
        lines = [ln for ln in p.stderr.split("\n") if ln.startswith("[stcsp r0] init:") or ln.startswith("[stcsp r0] upload_model:")]
        out["warm_process_first_solve_e2e_ms"] = j["solves"][0]["e2e_ms"]
        out["warm_process_first_solve_phases"] = "; ".join(ln[len("[stcsp r0] "):] for ln in lines) or None
        out["warm_process_first_over_third"] = j["solves"][0]["e2e_ms"] / max(j["solves"][2]["e2e_ms"], 1e-9)
        # ... and after stcsp_gpu_warmup(), the call a long-lived process makes once
        p = subprocess.run([sys.executable, "-c", COLD_SNIPPET % {"root": ROOT, "name": name, "warm": "", "api": True}], capture_output=True,
                           text=True, timeout=600)
        j = json.loads(p.stdout.strip().split("\n")[-1])
        # (wall clock of the C call itself, stcsp_automaton_t::wall_ms: the first Python-side wrapping of a result costs 3-4 ms
        #  of numpy / ctypes set-up that the other two figures have behind them or are dominated by)
        out["after_warmup_first_solve_e2e_ms"] = j["solves"][0]["call_ms"]
        out["after_warmup_first_over_third"] = j["solves"][0]["call_ms"] / max(j["solves"][2]["call_ms"], 1e-9)
        out["after_warmup_first_solve_python_ms"] = j["solves"][0]["e2e_ms"]
        out["warmup_call_ms"] = j.get("warmup_call_ms")
        out["what"] = ("first_solve_*: fresh process, the first call into the library (includes its one-time set-up: streams, device and "
                       "pinned arenas, kernel modules); warm_process_*: fresh process, one solve of ANOTHER model first, then the first "
                       "solve of this model (compile, upload, relation tables, pools: what a new model costs); after_warmup_*: fresh process, "
                       "stcsp_gpu_warmup(device, 512 MiB pinned) first (warmup_call_ms, the one-time set-up as an explicit call), then the first "
                       "solve of this model")
    except Exception as e:          # noqa: BLE001 -- a benchmark side figure must not kill the headline
        out = {"error": "%s: %s" % (type(e).__name__, e)}
    if os.path.exists(CLI_BIN):
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, name + ".csp")
            with open(path, "w") as f:
                f.write(text)
            t0 = time.perf_counter()
            p = subprocess.run([CLI_BIN, "-s", path], cwd=d, capture_output=True, text=True)
            out["cli_wall_ms"] = (time.perf_counter() - t0) * 1e3
            out["cli_rc"] = p.returncode
            out["cli_stat_line"] = p.stdout.strip().split("\n")[-1] if p.stdout.strip() else ""
        out["cli_what"] = "bin/stcsp -s file.csp: process start, CUDA context, parse, cold solve, post-processing, solutions.dot"
        if ref_wall_s:
            out["ref_cli_wall_s"] = ref_wall_s
            out["cli_speedup_wall"] = ref_wall_s / (out["cli_wall_ms"] / 1e3)
    return out


def l2_flush(torch, scratch):
    scratch.add_(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed solves (default: 200 for the B200 arm -- a timed region long enough for the clock sampler; "
                         "8 complete reference solves for the reference arm)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instance", default="juggling_b6_f6_nosym")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="length of the oracle-port sample (second CPU figure)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cold", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra workloads reported under `also`")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.steps is None:
        args.steps = 200 if args.impl == "b200" else 8

    from stcsp_solver_b200 import binding, instances
    name = args.instance
    text = instances.by_name(name)
    if args.impl == "reference":
        return run_reference_arm(args, name, text)

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from stcsp_solver_b200 import distributed

    model = binding.Model(text)
    V, K = model.n_vars, model.problem.contents.prefix_k
    ref_nodes = REF_WORK.get(name, (None, None))[0]
    scratch = torch.zeros(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def one_solve(profile=False):
        opts = binding.default_options(device=local, profile_kernels=1 if profile else 0)
        t0 = time.perf_counter()
        # N > 1: the headline instance does not shard (8 waves of at most 1 200 nodes), so the ranks solve
        # independent REPLICAS of it -- one complete solve per GPU and step, no collective on the data path -- and `value`
        # is the aggregate.  Instances that do shard are measured in the `sharded` leg below.
        a = binding.solve(model, opts)
        return a, time.perf_counter() - t0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        one_solve()
        l2_flush(torch, scratch)
    sampler.start()
    t_wait = time.perf_counter()
    while sampler.nv is not None and not sampler.sm and time.perf_counter() - t_wait < 0.5:
        time.sleep(0.0005)                  # the first sample is in before the timed region starts
    barrier()
    dev_ms, wall_s, last, kern = [], [], None, []
    region0 = time.perf_counter()
    for _ in range(args.steps):
        l2_flush(torch, scratch)
        torch.cuda.synchronize()
        a, w = one_solve()
        wall_s.append(w)
        if a is not None:
            dev_ms.append(a.c.solve_ms)
            kern.append((a.c.expand_ms, a.c.n_expand_launches))
            last = a
    barrier()
    region_s = time.perf_counter() - region0
    clocks = sampler.result()

    # max over ranks of the per-step times
    t = torch.tensor([sum(wall_s), sum(dev_ms) if dev_ms else 0.0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_total, dev_total = float(t[0]), float(t[1])

    # parity of what was just timed: EVERY rank checks its own replica
    parity = parity_of(binding, model, last, name)
    if dist is not None:
        okt = torch.tensor([1 if parity["sha256_ok"] is not False else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        parity["all_ranks_ok"] = bool(int(okt[0]) == 1)
        if not parity["all_ranks_ok"]:
            parity["sha256_ok"] = False

    # dominant kernel of the timed steps: the persistent search kernel, timed by CUDA events around every launch
    # inside the library (expand_ms / n_expand_launches of each step); one extra step-wise pass times expand alone
    prof = None
    if world == 1:
        l2_flush(torch, scratch)
        a, _ = one_solve(profile=True)
        prof = (a.c.expand_ms, a.c.n_expand_launches, a.c.n_search_nodes, a.c.solve_ms)

    # the other configurations of BASELINE.json (context for the headline, not timed above): best of 3 after one
    # untimed solve, each checked against its golden
    also = []
    if world == 1 and not args.no_also:
        peak_gbs = measured_peak()[0]
        binding.warmup(local, 1 << 30)      # pinned host arena for the results below (first_solve_e2e_ms: a new model in a warm process)
        for other in ("juggling_b4_f4", "digitinvader9", "partialorder_14", "juggling_b8_f8_nosym", "partialorder_16",
                      "partialorder_18", "partialorder_20"):
            if other == name:
                continue
            try:
                om = binding.Model(instances.by_name(other))
                t0 = time.perf_counter()
                oa = binding.solve(om)
                first_ms = (time.perf_counter() - t0) * 1e3
                dev, wall = [], []
                reps = 1 if other == "partialorder_20" else 3
                for _ in range(reps):
                    del oa
                    l2_flush(torch, scratch)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    oa = binding.solve(om)
                    wall.append(time.perf_counter() - t0)
                    dev.append(oa.c.solve_ms)
                ost = oa.stats()
                rec = {"workload": other, "device_ms": min(dev), "e2e_ms": min(wall) * 1e3, "first_solve_e2e_ms": first_ms,
                       "states": ost["n_states"], "edges": ost["n_edges"], "search_nodes": ost["n_search_nodes"],
                       "waves": ost["n_waves"], "d2h_bytes": ost["d2h_bytes"],
                       "algorithmic_gbs": ost["algorithmic_bytes"] / (min(dev) / 1e3) / 1e9,
                       "hbm_frac": ost["algorithmic_bytes"] / (min(dev) / 1e3) / 1e9 / peak_gbs}
                t0 = time.perf_counter()
                rec["parity"] = parity_of(binding, om, oa, other)
                rec["parity"]["host_check_s"] = time.perf_counter() - t0
                del oa
                also.append(rec)
            except binding.StcspError as e:
                also.append({"workload": other, "error": str(e)})
            binding.release_caches()

    # N > 1: the sharded search itself (BASELINE.json config 5), sharding forced, against the same instance on one GPU.
    # Every rank takes part in the group solves; rank 0 alone runs the single-GPU comparison while the others wait.
    sharded = []
    if world > 1 and not args.no_also:
        for other in ("juggling_b8_f8_nosym", "partialorder_16", "partialorder_18", "partialorder_20"):
            om = binding.Model(instances.by_name(other))
            rec = {"workload": other, "n_gpus": world}
            try:
                best = None
                for i in range(3):
                    barrier()
                    t0 = time.perf_counter()
                    oa = distributed.solve_distributed(om, binding.default_options(device=local), adaptive=False)
                    wall = (time.perf_counter() - t0) * 1e3
                    t2 = torch.tensor([wall], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
                    if rank == 0 and i > 0 and (best is None or oa.c.solve_ms < best[0]):
                        best = (oa.c.solve_ms, float(t2[0]), dict(oa.exchange_stats), oa.stats())
                    if i < 2:
                        del oa
                if rank == 0:
                    rec.update({"device_ms": best[0], "e2e_ms": best[1], "exchange": best[2], "states": best[3]["n_states"],
                                "edges": best[3]["n_edges"], "search_nodes": best[3]["n_search_nodes"],
                                "nvlink_bytes_per_solve_rank0": best[2]["bytes_pulled"]})
                    rec["parity"] = parity_of(binding, om, oa, other)
                del oa
                binding.release_caches()
                if rank == 0:       # the same instance on one GPU, same process, same box
                    binding.solve(om, binding.default_options(device=local))
                    one = []
                    for _ in range(2):
                        t0 = time.perf_counter()
                        o1 = binding.solve(om, binding.default_options(device=local))
                        one.append((o1.c.solve_ms, (time.perf_counter() - t0) * 1e3))
                        del o1
                    rec["one_gpu_device_ms"] = min(x[0] for x in one)
                    rec["one_gpu_e2e_ms"] = min(x[1] for x in one)
                    rec["speedup_device"] = rec["one_gpu_device_ms"] / rec["device_ms"]
                    rec["speedup_e2e"] = rec["one_gpu_e2e_ms"] / rec["e2e_ms"]
                    rec["efficiency_device"] = rec["speedup_device"] / world
                    binding.release_caches()
            except binding.StcspError as e:
                rec["error"] = str(e)
            barrier()
            if rank == 0:
                sharded.append(rec)

    if rank == 0:
        st = last.stats()
        steps = args.steps
        ms_dev = dev_total / steps      # world > 1: the merged automaton carries the max over ranks of the device times
        work = ref_nodes if ref_nodes else st["n_search_nodes"]
        unit = "reference search nodes/s" if ref_nodes else "search nodes/s"
        value = world * work / (ms_dev / 1e3)           # `world` replicas per step (one per GPU), slowest rank's time
        e2e = world * work / (wall_total / steps)
        peak, peak_src = measured_peak()
        roof = None
        k_ms = sum(x[0] for x in kern) / steps      # N > 1: rank 0's launches (the adaptive driver keeps small instances there)
        k_launches = sum(x[1] for x in kern) / steps
        if k_ms > 0:
            bytes_per_node = 2 * V * K * 8          # one domain block read + one written, (lb, ub) int32 pairs
            achieved = st["algorithmic_bytes"] / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
            counters = {}
            cp = os.path.join(ROOT, "profiles", "counters.json")
            if os.path.exists(cp):
                with open(cp) as f:
                    counters = json.load(f).get(name, {})
            sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
            int_ops = None
            if counters.get("thread_inst_executed"):
                rate = counters["thread_inst_executed"] / (k_ms / max(k_launches, 1) / 1e3)
                int_peak = 148 * 128 * sm_mhz * 1e6
                int_ops = {"thread_inst_per_launch": counters["thread_inst_executed"], "achieved_per_s": rate,
                           "peak_per_s": int_peak, "frac": rate / int_peak,
                           "source": "ncu smsp__thread_inst_executed.sum of one launch (%s) / live launch duration; peak = 148 SMs x 128 "
                                     "lanes x %d MHz" % (counters.get("source", "profiles/"), sm_mhz)}
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": counters.get("dram_bytes"), "peak_source": peak_src,
                    "kernel": "search_kernel (persistent wave loop: expand + route + ingest)",
                    "algorithmic_bytes_per_launch": st["algorithmic_bytes"] / max(k_launches, 1),
                    "bytes_per_unit": bytes_per_node, "units_per_launch": st["n_search_nodes"] / max(k_launches, 1),
                    "launches_per_step": k_launches, "avg_launch_us": k_ms / max(k_launches, 1) * 1e3,
                    "share_of_step": k_ms / ms_dev if ms_dev else None, "int_ops": int_ops,
                    "expand_only": None if prof is None else {
                        "ms_per_step": prof[0], "launches": prof[1], "step_ms_stepwise": prof[3],
                        "achieved_gbs": prof[2] * bytes_per_node / (prof[0] / 1e3) / 1e9 if prof[0] > 0 else 0.0},
                    "note": "integer-issue and latency bound: %d tuple evaluations and %d propagator runs per %d-byte node"
                            % (st["n_tuples"] // max(st["n_search_nodes"], 1),
                               st["n_revisions"] // max(st["n_search_nodes"], 1), bytes_per_node)}
        cpu = None
        ref_wall = None
        if world == 1 and not args.no_cpu_baseline:
            second = port_sample(model, name, args.cpu_seconds)
            if reference_feasible(name) and ref_nodes:
                r = run_stcsp_ref(text, name)
                if r is not None:
                    ref_wall = r["wall_s"]
                    cpu = {"value": ref_nodes / r["wall_s"], "unit": unit, "cores": 1, "kind": "reference",
                           "cpu_model": cpu_model(), "host_cores": os.cpu_count(),
                           "sample": "one complete solve of %s by oracle/_ref/stcsp_ref (the unmodified reference, single-threaded): "
                                     "wall %.2f s, its own solveTime %.2f s" % (name, r["wall_s"], r["solve_s"]),
                           "solve_time_s": r["wall_s"], "own_solveTime_s": r["solve_s"], "port_sample": second}
            if cpu is None and second is not None:
                cpu = {"value": second["nodes"] / second["solve_time_s"], "unit": unit, "cores": 1, "kind": "port",
                       "cpu_model": cpu_model(), "host_cores": os.cpu_count(), "sample": second["sample"],
                       "solve_time_s": second["solve_time_s"]}
        cold = None
        if world == 1 and not args.no_cold:
            binding.release_caches()
            cold = cold_numbers(name, text, ref_wall)
        line = {
            "metric": "search_nodes_per_s", "value": value, "unit": unit, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": name, "unit_of_work": "reference generalisedArcConsistent calls (%s per solve)" % work,
                       "l2": "flushed between steps (192 MiB write)", "timing": "CUDA events on the library stream around the whole search, per step" + ("" if world == 1 else ", max over ranks"),
                       "vars": V, "prefix_k": K, "state": "steady (compiled model resident from the previous solve; see `cold`)",
                       "replicas": world,
                       "parallelism": "1 GPU" if world == 1 else (
                           "%d independent replicas, one complete solve per GPU and step, no data-path collective (the headline "
                           "instance is too small to shard: 8 waves of at most 1 200 search nodes); value = aggregate over the GPUs, "
                           "solve_time_s = one solve on the slowest rank.  Instances that shard (states owned by signature hash, "
                           "device-side exchange over NVLink) are in `sharded`: strong scaling against one GPU" % world)},
            "solve_time_s": ms_dev / 1e3,
            "own_search_nodes_per_s": world * st["n_search_nodes"] / (ms_dev / 1e3),
            "states_per_s": world * st["n_states"] / (ms_dev / 1e3), "edges_per_s": world * st["n_edges"] / (ms_dev / 1e3),
            "automaton": {"states": st["n_states"], "edges": st["n_edges"], "search_nodes": st["n_search_nodes"],
                          "waves": st["n_waves"], "tuples": st["n_tuples"]},
            "parity": parity,
            "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
                    "what": "wall clock of stcsp_gpu_solve(problem in host memory) -> automaton in pinned host memory.  Steady "
                            "state: the compiled model (bytecode, relation tables) is resident in HBM, found by a hash of the host "
                            "problem recomputed on every call, so the bytes that go down per solve are the kernel's launch "
                            "parameters (model descriptor + start values; the root state is created on the device); the "
                            "automaton comes up through mapped pinned memory, written by the search kernel itself.  `cold` "
                            "has the first solve, which uploads and builds everything.",
                    "ms_per_step": wall_total / steps * 1e3},
            "gpu_launches": st["n_kernel_launches"] * steps * world,
            "clocks": clocks, "timed_region_s": region_s,
        }
        if also:
            line["also"] = also
        if sharded:
            line["sharded"] = sharded
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        if cold:
            line["cold"] = cold
        if parity is not None and parity["sha256_ok"] is False:
            # a fast solve of the wrong automaton is not a result
            print(json.dumps({"error": "parity failure: the timed automaton differs from the golden", "parity": parity,
                              "config": line["config"]}), flush=True)
            if dist is not None:
                dist.barrier()
                dist.destroy_process_group()
            raise SystemExit(1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
