"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: the collectives of one solve and the
assembly of per-rank parts.  The sessions themselves need CUDA and are covered by -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDENS, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _split_parts(automaton, world):
    """Re-shard a complete single-rank automaton into `world` parts the way the sessions would hold it:
    global id = local * world + rank, edges live with the owner of their destination."""
    n = automaton.n_states
    owner = np.array([0] + [(s * 7 + 3) % world for s in range(1, n)])      # any assignment with root on rank 0
    local = np.zeros(n, dtype=np.int64)
    cnt = [0] * world
    for s in range(n):
        local[s] = cnt[owner[s]]
        cnt[owner[s]] += 1
    gid = local * world + owner
    parts = []
    c = automaton.c
    for r in range(world):
        mine = np.where(owner == r)[0]
        order = mine[np.argsort(local[mine])]
        emask = owner[automaton.edge_dst] == r
        head = np.array([c.n_vars, c.n_sig_vars, c.n_until, c.n_until_vars, c.sig_len, c.root_final, len(order), int(emask.sum())]
                        + [1] + [0] * 12, dtype=np.int64)
        parts.append({"head": head, "times": np.zeros(3), "sig_vars": automaton.sig_vars,
                      "state_sig": automaton.state_sig[order].reshape(-1).astype(np.int32),
                      "state_cset": automaton.state_cset[order].astype(np.int32),
                      "edge_src": gid[automaton.edge_src[emask]].astype(np.int32),
                      "edge_dst": gid[automaton.edge_dst[emask]].astype(np.int32),
                      "edge_label": automaton.edge_label[emask].reshape(-1).astype(np.int32)})
    return parts


def _worker(rank, world, port, name, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _oracle
        from stcsp_solver_b200 import binding, distributed
        ex = distributed.WaveExchange()
        # 1. counts + payload all-to-all: rank r sends (q + 1) records to rank q, tagged with (r, q)
        send_counts = np.array([q + 1 for q in range(world)], dtype=np.int64)
        outbox = torch.cat([torch.full((q + 1, 5), 10 * rank + q, dtype=torch.int32) for q in range(world)])
        recv_counts = ex.exchange_counts(send_counts)
        assert list(recv_counts) == [rank + 1] * world
        inbox = ex.exchange_records(outbox, send_counts, recv_counts)
        want = torch.cat([torch.full((rank + 1, 5), 10 * r + rank, dtype=torch.int32) for r in range(world)])
        assert torch.equal(inbox, want)
        # 2. termination sum and the sorted union of resolve requests (same on every rank)
        assert ex.total(rank + 1) == world * (world + 1) // 2
        rows = np.array([[0, 5 - rank, 1], [0, 9, 9]], dtype=np.int32) if rank else np.zeros((0, 3), dtype=np.int32)
        uni = ex.union_rows(rows)
        assert [tuple(r) for r in uni.tolist()] == sorted({tuple(r) for k in range(1, world) for r in [[0, 5 - k, 1], [0, 9, 9]]})
        # 3. parts gathered to rank 0 and assembled: same canonical automaton as the unsharded one
        g = GOLDENS[name]
        model = binding.Model(g["model"]) if "model" in g else None
        if model is None:
            from stcsp_solver_b200 import instances
            model = binding.Model(instances.by_name(name))
        full, _ = _oracle.solve(model)                  # test infrastructure stands in for the GPU sessions
        parts = _split_parts(full, world)
        gathered = ex.gather_arrays(parts[rank])
        if rank == 0:
            merged = binding.assemble(gathered, trim=True)
            sol = binding.Solution(model, merged)
            q.put((sol.canonical_sha256(), sol.n_states, sol.n_edges))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["juggling_b4_f5_nosym", "probe_until", "probe_dead_branch"])
def test_exchange_and_assemble_world2(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sha, states, edges = q.get(timeout=5)
    g = GOLDENS[name]
    assert (sha, states, edges) == (g["sha256"], g["states"], g["edges"])
