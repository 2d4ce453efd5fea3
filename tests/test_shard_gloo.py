"""world_size-2 test of the coalition sharding + the single all-gather, on CPU with gloo.

The CUDA engine cannot run here, so a stand-in evaluator with the engine's call signature (test
infrastructure) computes a deterministic function of the coalition bits; the sharding, padding,
gather and re-assembly logic under test is the product's (shard.sharded_eval)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakeEngine:
    n_query = 2

    def __call__(self, act, n_coalitions, s0=0, n_s=None):
        n_s = n_coalitions - s0 if n_s is None else n_s
        rows = []
        for s in range(s0, s0 + n_s):
            bits = (act[:, s // 32] >> (s % 32)) & 1
            k = float(bits.sum())
            rows.append([k, k * 0.5 + s])
        return torch.tensor(rows, dtype=torch.float32).reshape(n_s, 2)


def _worker(rank, world, port, n_coalitions, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bikg_graph_explainability_public_b200.shard import sharded_eval

    g = torch.Generator().manual_seed(0)
    act = torch.randint(0, 2 ** 31 - 1, (50, -(-n_coalitions // 32)), generator=g, dtype=torch.int64).to(torch.int32)
    y = sharded_eval(FakeEngine(), act, n_coalitions)
    ref = FakeEngine()(act, n_coalitions)
    out[rank] = bool(torch.equal(y, ref))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_eval_world2():
    for n_coalitions in (1002, 31, 96):  # ragged last tile, fewer tiles than ranks, exact multiple
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), n_coalitions, out), nprocs=2, join=True)
        assert out[0] and out[1], n_coalitions


def _rng_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bikg_graph_explainability_public_b200.explainer import check_ranks_agree, sync_rng_across_ranks

    torch.manual_seed(100 + rank)  # the usual "seed + rank" pattern: ranks would draw different coalition masks
    sync_rng_across_ranks()         # rank 0's generator becomes authoritative (times > 1 does not re-seed, explainer.py:342)
    draws = torch.randint(0, 2 ** 31 - 1, (40, 3), dtype=torch.int64).to(torch.int32)
    gathered = [torch.empty_like(draws) for _ in range(world)]
    dist.all_gather(gathered, draws)
    same_stream = all(torch.equal(g, gathered[0]) for g in gathered)
    check_ranks_agree(draws)        # identical coalition matrices: passes
    act = draws.clone()
    if rank == 1:
        act[7, 1] ^= 4              # one coalition bit differs on one rank
    try:
        check_ranks_agree(act)
        raised = False
    except RuntimeError:
        raised = True
    out[rank] = (same_stream, raised)
    dist.destroy_process_group()


def test_ranks_replay_one_coalition_stream_world2():
    """Explainer.run under one process per GPU: every rank generates the coalition masks itself, so the CPU generators must
    agree (rank 0's state is broadcast) and a divergence must raise instead of pairing predictions with wrong rows."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rng_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] == (True, True) and out[1] == (True, True)
