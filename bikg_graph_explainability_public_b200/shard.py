"""Coalition data parallelism: one process per GPU, graph and weights replicated, each rank
evaluates a contiguous range of 32-coalition words and a single all-gather collects the query
predictions before the (replicated, sequential) surrogate fit.  SURVEY.md 8e.

The reference has no distributed code; this is the only collective of the path.  Without an
initialised ``torch.distributed`` process group everything runs on the local device.
"""
import torch
import torch.distributed as dist


def word_range(n_words, rank, world):
    """Contiguous, balanced split of the coalition words (tiles) over ranks."""
    base, extra = divmod(n_words, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_eval(engine, act, n_coalitions, group=None):
    """y (n_coalitions, n_query) on every rank.  ``engine(act, n_coalitions, s0, n_s)`` evaluates a range."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return engine(act, n_coalitions)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_words = -(-n_coalitions // 32)
    lo, hi = word_range(n_words, rank, world)
    s0, s1 = lo * 32, min(hi * 32, n_coalitions)
    per_rank = (-(-n_words // world)) * 32  # padded rows per rank so that all_gather is regular
    local = None
    if s1 > s0:
        local = engine(act, n_coalitions, s0, s1 - s0)
    nq = engine.n_query
    dev = act.device
    buf = torch.zeros((per_rank, nq), dtype=torch.float32, device=dev)
    if local is not None:
        buf[: s1 - s0] = local
    out = torch.empty((world * per_rank, nq), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)  # the single collective (NCCL over NVLink on GPUs)
    parts = []
    for r in range(world):
        rlo, rhi = word_range(n_words, r, world)
        cnt = max(min(rhi * 32, n_coalitions) - rlo * 32, 0)
        parts.append(out[r * per_rank: r * per_rank + cnt])
    return torch.cat(parts, dim=0)
