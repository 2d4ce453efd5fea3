"""Community-coalition mask generation (device) behind the reference's ``Mask`` interface.

Mirror of ``pathway_explanations.masks.Mask`` (reference ``masks.py:10-397``): same constructor,
same ``mask_generator() -> (loader, pathway_rows)`` contract and the same RNG stream, but the
coalition bits are produced by CUDA kernels that replay torch's CPU MT19937 stream bit-exactly
and are stored packed node-major (one bit per node and coalition) instead of as a (rows, N)
bool matrix.  The returned ``CoalitionSet`` iterates like the reference's DataLoader (batches of
``rows // epochs`` bool rows) for callers that want the materialised view.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .engine import require_cuda
from .rng import DeviceStream


def row_plan(lengths, total):
    """Rows per community (reference ``masks.py:116-125``) in the reference's float32 arithmetic.

    ``len(p) / torch.sum(len_pathways)`` is ``int / int64-tensor`` = ``reciprocal(tensor) * int`` in
    torch, i.e. fl32(fl32(1/sum) * len) -- not fl32(len/sum)."""
    recip = np.float32(1.0) / np.float32(sum(int(x) for x in lengths))
    plan = []
    for ln in lengths:
        frac = np.float32(recip * np.float32(int(ln)))
        size = math.ceil(float(np.float32(frac * np.float32(total))))
        size_int = math.ceil(float(np.float32(frac * np.float32(size))))
        if size_int < 3:
            size_int, size = 1, 2
        plan.append((size, size_int))
    return plan


class CoalitionSet:
    """Packed coalition matrix: ``act[v, w]`` bit ``b`` = node ``v`` active in coalition ``32 w + b``."""

    def __init__(self, act, popcount, n_coalitions, n_elements, batch_size, pathway_rows=None, expander=None):
        self.act, self.popcount = act, popcount
        self.n_coalitions, self.n_elements, self.batch_size = n_coalitions, n_elements, batch_size
        self.pathway_rows = pathway_rows
        self._expander = expander
        self._dense = None

    @property
    def words(self):
        return int(self.act.shape[1])

    def dense(self):
        """(rows, N) bool tensor -- the reference's mask matrix (materialised on demand)."""
        if self._dense is None:
            self._dense = self._expander()
        return self._dense

    @property
    def dataset(self):  # DataLoader compatibility (reference tests read loader.dataset)
        return self.dense()

    def __len__(self):
        return -(-self.n_coalitions // self.batch_size)

    def __iter__(self):
        # iter(DataLoader) draws one int64 base seed from the global CPU generator (wlm.py:210)
        torch.empty((), dtype=torch.int64).random_()
        d = self.dense()
        for s in range(0, self.n_coalitions, self.batch_size):
            yield d[s:s + self.batch_size]


def _i32(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def generate_coalitions(n_elements, communities, params, device=None):
    """Device port of ``Mask.mask_generator`` for node problems.  Consumes torch's global CPU
    generator exactly like the reference and leaves it advanced by the same number of draws."""
    lib = _lib.load()
    dev = require_cuda() if device is None else device
    n_perturbs, epochs = Mask.assertions_mask_generator(params)
    total, epochs = int(n_perturbs * epochs), int(epochs)
    n = int(n_elements)
    stream = DeviceStream(dev)
    st = _lib.stream_ptr()

    if communities is None:  # Shapley mode (masks.py:231-260, 362-365)
        snap0 = stream.snapshot()
        n_draws = total * n + max(total - 1, 0)
        draws = stream.draw(n_draws)
        ind = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.xpgnn_randperm(draws.data_ptr() + 4 * total * n, total, ind.data_ptr(), st))
        stream.hand_back()
        w = -(-total // 32)
        act = torch.zeros((n, max(w, 1)), dtype=torch.int32, device=dev)
        pop = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.xpgnn_shapley_expand(draws.data_ptr(), ind.data_ptr(), total, n, None, act.data_ptr(), w,
                                            pop.data_ptr(), st))

        def expand():
            # the draws (one int32 per mask bit) are not kept alive: the dense view replays them from the saved state
            d = stream.replay(snap0, n_draws)
            m = torch.empty((total, n), dtype=torch.uint8, device=dev)
            _lib.check(lib.xpgnn_shapley_expand(d.data_ptr(), ind.data_ptr(), total, n, m.data_ptr(), None, w,
                                                None, _lib.stream_ptr()))
            return m.bool()

        del draws

        if total < epochs:
            raise ValueError("batch_size should be a positive integer value, but got batch_size=0")
        return CoalitionSet(act, pop, total, n, total // epochs, None, expand)

    c = len(communities)
    lens = [len(p) for p in communities]
    order = torch.argsort(torch.tensor(lens), descending=True).tolist()  # masks.py:314 (unstable, CPU)
    plan = row_plan(lens, total)
    sizes, sizes_int, visited, cumulative = [], [], [], 0
    for cid in order:  # masks.py:322-348
        communities[cid].sort()  # masks.py:323 mutates the caller's lists
        sizes.append(plan[cid][0])
        sizes_int.append(plan[cid][1])
        visited.append(cid)
        if cumulative > total and n > 4000:
            break
        cumulative += plan[cid][0]
    n_pos, rows = len(visited), int(sum(sizes))
    row_start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    com_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    flat = (np.concatenate([np.sort(np.asarray(p, dtype=np.int64)) for p in communities]) if c else np.zeros(0, dtype=np.int64))
    if flat.size and (flat.min() < 0 or flat.max() >= n):
        raise IndexError("community member outside [0, %d)" % n)
    com_of = np.repeat(np.arange(c, dtype=np.int32), lens)
    slot_of = (np.arange(flat.size) - np.repeat(com_ptr[:-1], lens)).astype(np.int32)
    o = np.argsort(flat, kind="stable")
    node_ptr = np.concatenate([[0], np.cumsum(np.bincount(flat, minlength=n))]).astype(np.int32)
    keep = [_i32(com_ptr, dev), _i32(flat, dev), _i32(node_ptr, dev), _i32(com_of[o], dev), _i32(slot_of[o], dev),
            _i32(visited, dev), _i32(sizes, dev), _i32(sizes_int, dev), _i32(row_start, dev)]
    mp = _lib.MaskPlan(n, c, n_pos, rows, *[t.data_ptr() for t in keep])

    truncate = n > 4000 and rows > total  # masks.py:367-380
    vis_len = np.ascontiguousarray([lens[cid] for cid in visited], dtype=np.int32)
    hs, hi = np.ascontiguousarray(sizes, dtype=np.int32), np.ascontiguousarray(sizes_int, dtype=np.int32)
    max_draws = int(lib.xpgnn_mask_max_draws(hs.ctypes.data, hi.ctypes.data, vis_len.ctypes.data, n_pos, c,
                                             0 if truncate else 1))
    snap = stream.snapshot()
    draws = stream.draw(max_draws)
    offsets = torch.empty(max(4 * n_pos, 1), dtype=torch.int64, device=dev)
    consumed = torch.zeros(1, dtype=torch.int64, device=dev)
    ind = torch.empty(max(rows, 1), dtype=torch.int32, device=dev)
    _lib.check(lib.xpgnn_mask_resolve(C.byref(mp), draws.data_ptr(), offsets.data_ptr(), consumed.data_ptr(),
                                      0 if truncate else 1, ind.data_ptr(), st))
    used = int(consumed.item())
    if used != max_draws:  # fewer dead-mask repairs than the bound: rewind and consume exactly `used`
        stream.restore(snap)
        stream.draw(used)
    stream.hand_back()
    n_out = rows
    if truncate:
        psz = torch.from_numpy(np.repeat(vis_len, sizes).astype(np.int32))
        ind = torch.argsort(psz, descending=True)[:total].to(torch.int32).to(dev)  # masks.py:379-380
        n_out = int(ind.numel())
    w = -(-n_out // 32)
    act = torch.zeros((n, max(w, 1)), dtype=torch.int32, device=dev)
    pop = torch.zeros(max(n_out, 1), dtype=torch.int32, device=dev)
    prow = torch.empty(max(n_out, 1), dtype=torch.int32, device=dev)
    _lib.check(lib.xpgnn_mask_expand(C.byref(mp), draws.data_ptr(), offsets.data_ptr(), ind.data_ptr(), n_out, None,
                                     act.data_ptr(), w, prow.data_ptr(), pop.data_ptr(), st))

    def expand():
        d = stream.replay(snap, used)  # see the Shapley branch: draws are replayed, not kept
        m = torch.empty((n_out, n), dtype=torch.uint8, device=dev)
        _lib.check(lib.xpgnn_mask_expand(C.byref(mp), d.data_ptr(), offsets.data_ptr(), ind.data_ptr(), n_out,
                                         m.data_ptr(), None, w, None, None, _lib.stream_ptr()))
        return m.bool()

    del draws

    expand._keep = keep  # the plan's device arrays must outlive the closure
    if n_out < epochs:
        raise ValueError("batch_size should be a positive integer value, but got batch_size=0")
    return CoalitionSet(act, pop[:n_out], n_out, n, n_out // epochs, prow[:n_out], expand)


class Mask:
    """Same constructor and ``mask_generator`` contract as the reference class (``masks.py:10-35``)."""

    def __init__(self, feat, edge_index, pathways, params, problem):
        self.feat, self.edge_index = feat, edge_index
        self.pathways, self.params, self.problem = pathways, params, problem

    @staticmethod
    def assertions_mask_generator(params):
        n_perturbs = params["interpret_samples"]
        epochs = params["epochs"]
        assert isinstance(n_perturbs, int) or isinstance(
            n_perturbs, float
        ), "Number of perturbations in batch is not numeric"
        assert isinstance(epochs, int) or isinstance(epochs, float), "Number of epochs in batch is not numeric"
        return abs(n_perturbs), abs(epochs)

    def mask_generator(self):
        if "edge" in self.problem:
            raise NotImplementedError("edge-level coalitions: the reference reads self.edge_size, which is never set (masks.py:294)")
        feat = self.feat
        if isinstance(feat, dict):
            n = sum(int(t.shape[0]) for t in feat.values())
        else:
            n = int(feat.shape[0])
        cs = generate_coalitions(n, self.pathways, self.params)
        return cs, cs.pathway_rows
