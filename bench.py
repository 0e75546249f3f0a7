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

  value  device time (CUDA events on the library's stream around the whole wave loop), model tables
         already in HBM
  e2e    wall time of the C-ABI call stcsp_gpu_solve with HOST buffers in and out: device
         allocation, H2D of the model tables, the search, D2H of states and edges, host assembly + trim
  roofline      expand_kernel (the dominant kernel): SURVEY.md 8(d) algorithmic bytes per search node
                x nodes per launch / CUDA-event duration of the launches, against MEASURED_PEAKS.json
  cpu_baseline  the CPU restatement of the reference (oracle/, kind "port") on a bounded sample of the
                same instance, one host thread (the reference is single-threaded)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

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


def cpu_sample(model, name, seconds):
    """Bounded sample of the reference's search on one host core (oracle port).

    The search is far from stationary in nodes/s (the root's propagation dominates), but it is
    stationary in constraint evaluations/s, and the total evaluation count of the instance is a known
    constant; the full solve time is projected from the evaluations done in `seconds`."""
    import _oracle
    st = _oracle.sample(model, seconds)
    nodes, validates = REF_WORK.get(name, (None, None))
    if not st["timed_out"]:                      # the whole instance fitted in the sample
        solve_s = st["solve_s"]
        nodes = st["gac_calls"]
        what = "complete solve (%.2f s)" % solve_s
    elif validates:
        solve_s = st["solve_s"] * validates / max(st["validates"], 1)
        what = ("first %.1f s of the DFS = %d of %d constraint evaluations; solve time projected to %.1f s"
                % (st["solve_s"], st["validates"], validates, solve_s))
    else:
        return None, None, "no reference work count for %s" % name
    return nodes / solve_s, solve_s, what


def l2_flush(torch, scratch):
    scratch.add_(1)


def run_reference_arm(args, name, text):
    from stcsp_solver_b200 import binding
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = binding.Model(text)
    per_step = max(2.0, min(10.0, 150.0 / max(args.steps + args.warmup, 1)))
    for _ in range(args.warmup):
        cpu_sample(model, name, min(per_step, 2.0))
    vals, sols, what = [], [], ""
    t0 = time.time()
    for _ in range(args.steps):
        v, s, what = cpu_sample(model, name, per_step)
        vals.append(v)
        sols.append(s)
    wall = time.time() - t0
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "search_nodes_per_s", "value": value, "unit": "reference search nodes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": name, "unit_of_work": "reference generalisedArcConsistent calls (%s)" % (REF_WORK.get(name, ("?",))[0],)},
        "solve_time_s": sum(sols) / len(sols),
        "cpu_baseline": {"value": value, "unit": "reference search nodes/s", "cores": 1, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "reference search nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed solves (default: 200 for the B200 arm -- a timed region long enough for the clock sampler; "
                         "50 bounded samples for the reference arm)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instance", default="juggling_b6_f6_nosym")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the extra workloads reported under `also`")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.steps is None:
        args.steps = 200 if args.impl == "b200" else 50

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
        if world > 1:
            a = distributed.solve_distributed(model, opts)
        else:
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

    # dominant kernel of the timed steps: the persistent search kernel, timed by CUDA events around every launch
    # inside the library (expand_ms / n_expand_launches of each step); one extra step-wise pass times expand alone
    prof = None
    if world == 1:
        l2_flush(torch, scratch)
        a, _ = one_solve(profile=True)
        prof = (a.c.expand_ms, a.c.n_expand_launches, a.c.n_search_nodes, a.c.solve_ms)

    # the other single-GPU configurations of BASELINE.json, a few steps each (context for the headline, not timed above)
    also = []
    if world == 1 and not args.no_also:
        for other in ("juggling_b4_f4", "digitinvader9", "partialorder_14"):
            if other == name:
                continue
            om = binding.Model(instances.by_name(other))
            binding.solve(om)
            dev, wall = [], []
            for _ in range(3):
                l2_flush(torch, scratch)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                oa = binding.solve(om)
                wall.append(time.perf_counter() - t0)
                dev.append(oa.c.solve_ms)
            ost = oa.stats()
            also.append({"workload": other, "device_ms": min(dev), "e2e_ms": min(wall) * 1e3, "states": ost["n_states"],
                         "edges": ost["n_edges"], "search_nodes": ost["n_search_nodes"],
                         "algorithmic_gbs": ost["algorithmic_bytes"] / (min(dev) / 1e3) / 1e9})

    if rank == 0:
        st = last.stats()
        steps = args.steps
        ms_dev = dev_total / steps      # world > 1: the merged automaton carries the max over ranks of the device times
        work = ref_nodes if ref_nodes else st["n_search_nodes"]
        unit = "reference search nodes/s" if ref_nodes else "search nodes/s"
        value = work / (ms_dev / 1e3)
        e2e = work / (wall_total / steps)
        peak, peak_src = measured_peak()
        roof = None
        k_ms = sum(x[0] for x in kern) / steps      # N > 1: rank 0's launches (the adaptive driver keeps small instances there)
        k_launches = sum(x[1] for x in kern) / steps
        if k_ms > 0:
            bytes_per_node = 2 * V * K * 8          # one domain block read + one written, (lb, ub) int32 pairs
            achieved = st["algorithmic_bytes"] / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                with open(tp) as f:
                    traffic = json.load(f).get(name)
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "kernel": "search_kernel (persistent wave loop: expand + route + ingest)",
                    "algorithmic_bytes_per_launch": st["algorithmic_bytes"] / max(k_launches, 1),
                    "bytes_per_unit": bytes_per_node, "units_per_launch": st["n_search_nodes"] / max(k_launches, 1),
                    "launches_per_step": k_launches, "avg_launch_us": k_ms / max(k_launches, 1) * 1e3,
                    "share_of_step": k_ms / ms_dev if ms_dev else None,
                    "expand_only": None if prof is None else {
                        "ms_per_step": prof[0], "launches": prof[1], "step_ms_stepwise": prof[3],
                        "achieved_gbs": prof[2] * bytes_per_node / (prof[0] / 1e3) / 1e9 if prof[0] > 0 else 0.0},
                    "note": "integer-issue and latency bound: %d tuple evaluations and %d propagator runs per %d-byte node"
                            % (st["n_tuples"] // max(st["n_search_nodes"], 1),
                               st["n_revisions"] // max(st["n_search_nodes"], 1), bytes_per_node)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, s, what = cpu_sample(model, name, args.cpu_seconds)
            cpu = {"value": v, "unit": unit, "cores": 1, "kind": "port", "sample": what, "solve_time_s": s}
        line = {
            "metric": "search_nodes_per_s", "value": value, "unit": unit, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": name, "unit_of_work": "reference generalisedArcConsistent calls (%s per solve)" % work,
                       "l2": "flushed between steps (192 MiB write)", "timing": "CUDA events on the library stream around the whole search, per step" + ("" if world == 1 else ", max over ranks"),
                       "vars": V, "prefix_k": K,
                       "parallelism": "states sharded by signature hash x%d" % world + (
                           " (adaptive driver: every wave of this instance fits one GPU, so rank 0 solved it alone)"
                           if getattr(last, "exchange_stats", {}).get("single_gpu") else "")},
            "solve_time_s": ms_dev / 1e3,
            "own_search_nodes_per_s": st["n_search_nodes"] / (ms_dev / 1e3),
            "states_per_s": st["n_states"] / (ms_dev / 1e3), "edges_per_s": st["n_edges"] / (ms_dev / 1e3),
            "automaton": {"states": st["n_states"], "edges": st["n_edges"], "search_nodes": st["n_search_nodes"],
                          "waves": st["n_waves"], "tuples": st["n_tuples"]},
            "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
                    "ms_per_step": wall_total / steps * 1e3},
            "gpu_launches": st["n_kernel_launches"] * steps,
            "clocks": clocks, "timed_region_s": region_s,
        }
        if also:
            line["also"] = also
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
