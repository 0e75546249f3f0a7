"""torchrun --nproc-per-node N tools/multi_check.py [names...]: the group path with one PROCESS per GPU (CUDA IPC, NVLink):
every named instance sharded over the ranks, canonical hash against the golden, device / wall time, bytes over NVLink.
Environment: SHARD=0 lets the adaptive policy decide; REPS=n timed repetitions."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stcsp_solver_b200 import binding, distributed, instances  # noqa: E402


def golden(name):
    for fn in (name + ".json", "semantic_" + name + ".json"):
        p = os.path.join(ROOT, "tests", "golden", fn)
        if os.path.exists(p):
            return json.load(open(p))
    return None


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    names = sys.argv[1:] or ["partialorder_11", "probe_first_capture", "partialorder_14", "juggling_b6_f6_nosym"]
    shard = os.environ.get("SHARD", "1") == "1"
    reps = int(os.environ.get("REPS", "3"))
    for name in names:
        g = golden(name)
        text = g["model"] if g and "model" in g else instances.by_name(name)
        model = binding.Model(text)
        best = None
        for i in range(reps + 1):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a = distributed.solve_distributed(model, adaptive=not shard)
            wall = (time.perf_counter() - t0) * 1e3
            if rank == 0 and i > 0 and (best is None or a.c.solve_ms < best[0]):
                best = (a.c.solve_ms, wall, a.exchange_stats)
            last = a
        if rank == 0:
            sol = binding.Solution(model, last)
            ok = None
            if g and "sha256" in g and (sol.n_edges <= int(os.environ.get("HASH_MAX_EDGES", "5000000")) or g["edges"] != sol.n_edges):
                ok = sol.canonical_sha256_streamed() == g["sha256"]
            elif g:
                ok = "counts only: %s" % ((g["states"], g["edges"]) == (int(sol.n_states), int(sol.n_edges)))
            print(json.dumps({"instance": name, "world": world, "sharded": shard, "device_ms": best[0], "e2e_ms": best[1],
                              "states": int(sol.n_states), "edges": int(sol.n_edges), "sha256_ok": ok, "exchange": best[2]}), flush=True)
        del last
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
