"""Run under torchrun: multi-GPU solve of several instances, checked against the reference goldens on rank 0."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from stcsp_solver_b200 import binding, distributed, instances

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
names = sys.argv[1:] or ["juggling_b4_f4", "juggling_b4_f5_nosym", "juggling_b6_f6_nosym", "digitinvader3", "partialorder_12",
                         "partialorder_14"]
bad = 0
for name in names:
    path = os.path.join(ROOT, "tests", "golden", name + ".json")
    g = json.load(open(path)) if os.path.exists(path) else None
    text = g["model"] if g and "model" in g else instances.by_name(name)
    model = binding.Model(text)
    a = None
    for rep in range(4):
        a = None                # release the previous automaton (its pinned blocks go back to the cache)
        dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
        a = distributed.solve_distributed(model, adaptive=os.environ.get("ADAPTIVE", "1") == "1")
        torch.cuda.synchronize(); dt = time.time() - t0
    if dist.get_rank() == 0:
        sol = binding.Solution(model, a)
        ok = g is None or sol.canonical_sha256() == g["sha256"]
        bad += not ok
        st = a.stats()
        print("%-24s world %d %s states %d edges %d nodes %d waves %d dev_ms %.2f wall_ms %.1f sent %d" % (
            name, dist.get_world_size(), "OK " if ok else "BAD", sol.n_states, sol.n_edges, st["n_search_nodes"], st["n_waves"],
            st["solve_ms"], dt * 1e3, a.exchange_stats["records_sent"]), flush=True)
if dist.get_rank() == 0:
    print("bad:", bad)
dist.barrier()
dist.destroy_process_group()
