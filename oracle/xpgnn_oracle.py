"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (numpy for integer/bit work, CPU torch for fp32/fp64 arithmetic so
that BLAS/summation semantics equal the reference's) of the perturbation hot path
of ``pathway_explanations``.  Every function
cites the reference ``file:line`` it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.

Parity status: this port is pinned against the *reference itself*
run in the build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz`` and
``tests/test_oracle_vs_reference.py``), against the reference tests' golden vectors
for ``build_edge_mask`` / ``pathway_mask2node_mask`` / ``aggregate`` / SHAP kernel /
``weighted_mse_loss`` / ``regularizer`` / k-hop node sets (``tests/test_oracle_goldens.py``),
and against torch's CPU generator for the RNG stream.  The GNN layer arithmetic
(third-party PyG 2.0.4, absent from the reference tree) is restated from its
published algorithm in ``oracle/pyg_standin`` and is *unpinned* by reference tests.
"""
import itertools
import math

import numpy as np
import torch

from .mt19937 import MT19937

# --------------------------------------------------------------------------
# a1  set_seed                                                explainer.py:14-22


def seeded_stream(seed):
    """CPU stream origin used by ``Explainer.run(times=1)``: ``torch.manual_seed(seed + 2)``."""
    return MT19937(int(seed) + 2)


# --------------------------------------------------------------------------
# a2  k-hop computational graph                  data.py:281-361 + PyG k_hop_subgraph


def k_hop_subgraph(edge_index, query, num_hops):
    """(L+1)-hop in-neighbourhood, induced edges in original order, relabelled.

    ``num_hops`` is the value *after* the reference's ``n_hops += 1`` (data.py:328).
    Returns subset (sorted int64), sub_edge_index (2,E_sub) int64, query rank, edge_mask.
    The empty-edge fallback of data.py:337-339 (one self loop) is applied here too.
    """
    ei = np.asarray(edge_index, dtype=np.int64)
    src, dst = ei[0], ei[1]
    n = int(ei.max()) + 1 if ei.size else 0
    frontier = np.zeros(n, dtype=bool)
    seen = np.zeros(n, dtype=bool)
    frontier[query] = True
    seen[query] = True
    for _ in range(num_hops):
        nxt = np.zeros(n, dtype=bool)
        nxt[src[frontier[dst]]] = True
        seen |= nxt
        frontier = nxt  # PyG re-expands the whole last layer, not only unseen nodes
    subset = np.flatnonzero(seen).astype(np.int64)
    edge_mask = seen[src] & seen[dst]
    relabel = np.full(n, -1, dtype=np.int64)
    relabel[subset] = np.arange(subset.size)
    sub_ei = relabel[ei[:, edge_mask]]
    sub_ind = int(relabel[query])
    if edge_mask.sum() == 0:
        sub_ei = np.array([[sub_ind], [sub_ind]], dtype=np.int64)
    return subset, sub_ei, sub_ind, edge_mask


# --------------------------------------------------------------------------
# a3  community bookkeeping                                   pathways.py:33-136


def communities_in_subgraph(communities, community_names, sub_names):
    """pathways.py:33-102: keep communities with >=1 member among ``sub_names``."""
    names = np.array(sub_names, dtype=str)
    kept, kept_names, kept_pos = [], [], []
    for i, (com, cname) in enumerate(zip(communities, community_names)):
        common = np.intersect1d(np.array(com, dtype=str), names)
        if len(common) > 0:
            kept.append(common.tolist())
            kept_names.append(cname)
            kept_pos.append(i)
    return kept, kept_names, kept_pos


def names_to_indices(communities, sub_names):
    """pathways.py:104-136: indices in lexicographic *name* order (intersect1d)."""
    if isinstance(communities[0][0], int):
        return communities
    names = np.array(sub_names, dtype=str)
    out = []
    for com in communities:
        _, ind, _ = np.intersect1d(names, np.array(com, dtype=str), return_indices=True)
        out.append(ind.tolist())
    return out


# --------------------------------------------------------------------------
# a4  coalition masks                                 masks.py:80-397, pathways.py:234-385


def row_plan(lengths, total):
    """masks.py:116-125 in the reference's float32 arithmetic.

    Returns per community (size, size_internal).
    """
    lengths = [int(x) for x in lengths]
    # ``len(p) / torch.sum(...)`` is int / int64-tensor == Tensor.__rtruediv__ == reciprocal() * int:
    # fl32(fl32(1 / sum) * len), NOT fl32(len / sum) (differs e.g. for 10/12 * 834 -> 696 vs 695).
    recip = np.float32(1.0) / np.float32(sum(lengths))
    plan = []
    for ln in lengths:
        frac = np.float32(recip * np.float32(ln))
        size = math.ceil(float(np.float32(frac * np.float32(total))))
        size_int = math.ceil(float(np.float32(frac * np.float32(size))))
        if size_int < 3:
            size_int, size = 1, 2
        plan.append((size, size_int))
    return plan


def community_order(lengths):
    """masks.py:314: ``torch.argsort(descending=True)`` -- NOT stable; call torch itself."""
    return torch.argsort(torch.tensor([int(x) for x in lengths]), descending=True).tolist()


def mask_generator(n_elements, communities, params, mt):
    """masks.py:262-397.  ``communities``: list of index lists (sorted in place like the reference).

    Returns (mask[rows, n] bool after the row shuffle, pathway_rows int32, batch_size).
    ``mt`` is advanced exactly like torch's CPU generator.
    """
    total = int(abs(params["interpret_samples"]) * abs(params["epochs"]))
    epochs = int(abs(params["epochs"]))
    if communities is None:  # Shapley mode, masks.py:231-260,362-365
        mask = mt.randint_bool(total, n_elements)
        ind = mt.randperm(mask.shape[0])
        mask = mask[ind]
        return mask, None, mask.shape[0] // epochs

    C = len(communities)
    lens = [len(c) for c in communities]
    order = community_order(lens)
    plan = row_plan(lens, total)
    blocks, rows_of, sizes_of = [], [], []
    cumulative = 0
    for pos, cid in enumerate(order):
        com = communities[cid]
        com.sort()  # masks.py:323 (mutates the caller's list)
        flat = np.array(list(itertools.chain.from_iterable(communities)), dtype=np.int64)
        size, size_int = plan[cid]
        internal = mt.randint_bool(size, len(com))  # RNG#1 masks.py:130
        block = np.zeros((size, n_elements), dtype=bool)
        # external community bits, antithetic halves          pathways.py:234-283
        half = (size - size_int) // 2
        ext = mt.randint_bool(half, C)  # RNG#2
        ext = np.vstack([ext, ~ext])
        if (size - size_int) % 2 != 0:
            ext = np.vstack([ext, mt.randint_bool(1, C)])  # RNG#3
        ext[:, pos] = False  # masks.py:178 -- sorted *position* on the unsorted list
        if C - 1 > 0 and ext.sum() == 0:  # masks.py:181 -> pathways.py:285-334
            perm = mt.randperm(C)  # RNG#4
            perm = perm[perm != pos]
            if ext.shape[0] > len(perm) and len(perm) > 0:
                q = ext.shape[0] // len(perm)
                perm = np.concatenate([perm] * (q + 1))
            perm = perm[: ext.shape[0]]
            ext[np.arange(ext.shape[0]), perm] = True
        # community -> node expansion                         pathways.py:336-385, masks.py:186-192
        node_on = np.repeat(ext, lens, axis=1)  # (rows_ext, sum lens)
        r, c = np.nonzero(node_on)
        block[r + size_int, flat[c]] = True
        block[:, com] = internal  # masks.py:338
        blocks.append(block)
        rows_of.append(np.full(size, cid, dtype=np.int32))
        sizes_of.append(np.full(size, len(com), dtype=np.int32))
        if cumulative > total and n_elements > 4000:  # masks.py:344-346
            break
        cumulative += size
    mask = np.vstack(blocks)
    pathway_rows = np.concatenate(rows_of)
    pathway_sizes = np.concatenate(sizes_of)
    if n_elements > 4000 and mask.shape[0] > total:  # masks.py:367-380
        ind = torch.argsort(torch.from_numpy(pathway_sizes), descending=True)[:total].numpy()
    else:
        ind = mt.randperm(mask.shape[0])  # RNG#5 masks.py:385
    mask = mask[ind]
    pathway_rows = pathway_rows[ind]
    return mask, pathway_rows, mask.shape[0] // epochs  # masks.py:225


# --------------------------------------------------------------------------
# a5  perturbation                                              data.py:390-648


def build_edge_mask(edge_index, mask):
    """data.py:390-451: block-diagonal edge list (int32) and keep flags."""
    ei = np.asarray(edge_index).astype(np.int32)
    b, n = mask.shape
    ones = np.flatnonzero(mask.reshape(-1))
    edges = np.hstack([ei + i * n for i in range(b)])
    keep = np.isin(edges[0].astype(np.int64), ones) & np.isin(edges[1].astype(np.int64), ones)
    return keep, edges


def perturbator(feat, edge_index, mask, edge_type=None):
    """data.py:591-648 (node problems): features replicated, edges filtered."""
    keep, edges = build_edge_mask(edge_index, mask)
    concat = torch.vstack([feat] * mask.shape[0]).float()
    pet = None
    if edge_type is not None:
        pet = np.hstack([np.asarray(edge_type).astype(np.int32)] * mask.shape[0])[keep]
    return concat, edges[:, keep].astype(np.int64), pet


# --------------------------------------------------------------------------
# a8  SHAP kernel                                              kernels.py:22-174

_BINOM_CACHE = {}


def _binom_table(ref):
    from scipy.special import binom

    if ref not in _BINOM_CACHE:
        _BINOM_CACHE[ref] = binom(ref, np.arange(ref)).astype(np.float64)
    return _BINOM_CACHE[ref]


def shap_kernel(mask):
    """kernels.py:115-174.  ``mask`` (B,N) bool -> float64 weights (B,)."""
    from scipy.special import binom

    mask = np.asarray(mask)
    k = mask.sum(axis=1).astype(np.int64)
    total = mask.shape[1] - 1
    with np.errstate(all="ignore"):
        if total > 1000:
            ref = 1000
            kern = np.zeros(total)
            while kern.sum() == 0 and ref > 0:
                choose = (_binom_table(ref) + 1e-10) * total / 1000
                # torch: (int64 * 1000 / int) -> float32 true division, then .long()
                idx = (torch.from_numpy(k) * 1000 / total).long().numpy()
                idx = np.clip(idx, 0, len(choose) - 1)
                # python_scalar / tensor is Tensor.__rtruediv__ == reciprocal(tensor) * scalar (two roundings)
                kern = (1.0 / (choose[idx] * k.astype(np.float64) * (total - k).astype(np.float64))) * float(total)
                s = kern.sum()
                if s > 0 and ref > 0:
                    break
                ref = int(0.9 * ref)
        else:
            choose = binom(total + 1, k).astype(np.float64)
            kern = (1.0 / (choose * (total + 1 - k) * k)) * total  # __rtruediv__: reciprocal * scalar
    kern = np.nan_to_num(kern, nan=0.0, posinf=0.0, neginf=0.0)
    return kern


# --------------------------------------------------------------------------
# a6/a7  black-box inference on the block-diagonal batch       model.py:62-328, wlm.py:284-438


def _forward_arity(arch):
    import inspect

    return len(inspect.getfullargspec(arch.forward).args)  # model.py:104 (getargspec there)


def _unique_types(t):
    return torch.unique(t)


def kernel_output(mask, feat, edge_index, arch, q, node_type=None, edge_type=None,
                  node_type_names=None, edge_type_names=None, padded_dims=None):
    """wlm.py:284-438 for node problems.  Returns (kernel f64 (B,), y)."""
    mask_np = np.asarray(mask)
    b, n = mask_np.shape
    concat, pei, pet = perturbator(feat, edge_index, mask_np, edge_type)
    pei_t = torch.from_numpy(pei)
    hetero = node_type is not None and edge_type is not None and node_type_names is not None \
        and edge_type_names is not None
    n_types = 1
    if node_type_names is not None:
        n_types = len(_unique_types(node_type))
    with torch.no_grad():
        if n_types < 2:
            if hetero:  # single node type: wlm.py:369-389 -> data.py:149-232
                ctype = torch.hstack([node_type] * b)
                x_dict = {}
                for i, nm in enumerate(node_type_names):
                    x = concat[torch.where(ctype == i)[0]]
                    if padded_dims is not None and padded_dims[i] > 0:
                        x = x[:, : -padded_dims[i]]
                    x_dict[nm] = x
                pet_t = torch.from_numpy(pet)
                ei_dict = {nm: pei_t[:, torch.where(pet_t == i)[0]].long()
                           for i, nm in enumerate(edge_type_names)}
                out = arch(x_dict, ei_dict)
            elif node_type is not None and edge_type is not None and _forward_arity(arch) == 5:
                # model.py:110-112: homogenised hetero graph, the model takes the replicated type vectors as well
                out = arch(concat, pei_t, torch.hstack([node_type] * b), torch.from_numpy(pet))
            else:
                out = arch(concat, pei_t)
            if isinstance(out, dict):  # model.py:255-292
                out = torch.vstack(list(out.values()))
            y = out[torch.arange(q, out.shape[0], n)]  # model.py:294-328
        else:  # multi node type: model.py:118-253 (per-coalition loop)
            y = torch.tensor(_predict_hetero(concat, pei_t, node_type, torch.from_numpy(pet),
                                             node_type_names, edge_type_names, b, n, q,
                                             padded_dims, arch))
    return shap_kernel(mask_np), y


def _predict_hetero(feat, edge_index, node_type, pet, node_type_names, edge_type_names, b, n,
                    q, padded_dims, arch):
    """model.py:118-253.  Deviation (documented): ``node_pointers`` is computed *before* the
    loop; the reference computes it inside ``if perturb == 0`` after the zero-edge ``continue``
    (model.py:213-222) and raises UnboundLocalError when the first coalition has no edge."""
    ntype = torch.hstack([node_type] * b)
    pointers = [int(torch.where(node_type == t)[0][0]) for t in _unique_types(node_type)]
    outs = []
    src = edge_index[0]
    for p in range(b):
        lo, hi = p * n, (p + 1) * n
        sel = torch.where((src >= lo) & (src < hi))[0]
        if sel.numel() == 0:
            outs.append(0.0)
            continue
        x = feat[lo:hi]
        t = ntype[lo:hi]
        ei = edge_index[:, sel] - lo
        et = pet[sel]
        x_dict = {}
        for i, nm in enumerate(node_type_names):
            xi = x[torch.where(t == i)[0]]
            if padded_dims is not None and padded_dims[i] > 0:
                xi = xi[:, : -padded_dims[i]]
            x_dict[nm] = xi
        ei_dict = {}
        for i, nm in enumerate(edge_type_names):
            m = ei[:, torch.where(et == i)[0]].long().clone()
            m[0] -= pointers[node_type_names.index(nm[0])]
            m[1] -= pointers[node_type_names.index(nm[-1])]
            ei_dict[nm] = m
        out = arch(x_dict, ei_dict)
        outs.append(float(out[q, 0]))
    return outs


# --------------------------------------------------------------------------
# a9  weighted linear surrogate                                  wlm.py:17-278, 441-520


def weighted_mse_loss(pred, target, weight):
    """wlm.py:491-520 (keeps the (B,1)-vs-(B,) broadcast of wlm.py:517)."""
    diff = (pred.flatten() - target) ** 2
    return torch.mean(weight * diff) / weight.sum()


def regularizer(w, factor):
    """wlm.py:101-129."""
    a = torch.abs(w.view(-1))
    return factor * (a.sum() / a.shape[0])


def train_wlm(batches, w0, params):
    """wlm.py:132-278 with torch autograd + Adam exactly as the reference.

    ``batches``: iterable of (mask (B,N) bool ndarray, kernel f64 (B,), y tensor).
    Returns final weights (N,) fp32 and the loss list.
    """
    w = torch.nn.Parameter(torch.from_numpy(np.asarray(w0, dtype=np.float32)).clone().view(1, -1))
    opt = torch.optim.Adam([w], lr=abs(params["lr"]), weight_decay=1e-2)  # wlm.py:478
    losses = []
    for mask, kern, y in batches:
        opt.zero_grad()
        x = torch.from_numpy(np.asarray(mask)).float()
        pred = torch.nn.functional.linear(x, w)
        loss = weighted_mse_loss(pred, y, torch.from_numpy(np.asarray(kern, dtype=np.float64)))
        loss = loss + regularizer(w, params["l1_lambda"])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return w.detach()[0].clone(), losses


def train_wlm_closed_form(batches, w0, params):
    """Closed form of one reference step (SURVEY.md section 7 'Closed-form fit step'); used to
    cross-check ``train_wlm`` and as the spec of the device kernel."""
    w = np.asarray(w0, dtype=np.float32).copy()
    n = w.size
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    lr, lam = abs(params["lr"]), params["l1_lambda"]
    for t, (mask, kern, y) in enumerate(batches, start=1):
        x = np.asarray(mask, dtype=np.float32)
        b = x.shape[0]
        y = np.asarray(y, dtype=np.float32)
        p = x @ w
        kern = np.asarray(kern, dtype=np.float64)
        ksum = kern.sum()
        if y.ndim == 2:  # (B,1): every p_j is pulled towards the batch mean of y
            coef = (2.0 * (kern / ksum) / (b * b)).astype(np.float32)
            r = b * p - y.sum(dtype=np.float32)
        else:
            coef = (2.0 * (kern / ksum) / b).astype(np.float32)
            r = p - y
        dp = coef * r
        g = x.T @ dp + np.float32(lam / n) * np.sign(w) + np.float32(1e-2) * w
        m = np.float32(0.9) * m + np.float32(0.1) * g
        v = np.float32(0.999) * v + np.float32(0.001) * g * g
        bc1, bc2 = 1 - 0.9 ** t, 1 - 0.999 ** t
        denom = np.sqrt(v) / np.float32(math.sqrt(bc2)) + np.float32(1e-8)
        w = w - np.float32(lr / bc1) * (m / denom)
    return w


# --------------------------------------------------------------------------
# a10  aggregation                                 explainer.py:288-314, pathways.py:387-429


def weight_stacking(weights):
    stack = torch.vstack([torch.as_tensor(w) for w in weights])
    return torch.mean(stack, 0), torch.std(stack, 0, unbiased=False)


def aggregate(config_val, community_inds):
    cv = torch.as_tensor(config_val)
    return [torch.mean(cv[torch.as_tensor(ci, dtype=torch.long)]).item() for ci in community_inds]


# --------------------------------------------------------------------------
# a11  hetero -> homo flattening                              data.py:39-147, 695-878


def flatten_hetero(feat, edge_index):
    """data.py:95-147.  dict graph -> padded features, offset edges, float type vectors."""
    ntypes = list(feat.keys())
    etypes = list(edge_index.keys())
    fmax = max(t.shape[1] for t in feat.values())
    padded, pads, nptr, ptr = [], [], [], 0
    for t in feat.values():
        d = fmax - t.shape[1]
        pads.append(d)
        nptr.append(ptr)
        ptr += t.shape[0]
        padded.append(torch.nn.functional.pad(t, (0, d)) if d > 0 else t)
    x = torch.vstack(padded)
    node_types = torch.hstack([torch.zeros(p.shape[0]) + i for i, p in enumerate(padded)])
    eis, ets, eptr, ptr = [], [], [], 0
    for i, (rel, ei) in enumerate(edge_index.items()):
        eptr.append(ptr)
        add = torch.tensor([[nptr[ntypes.index(rel[0])]], [nptr[ntypes.index(rel[-1])]]])
        eis.append(ei + add)
        ets.append(torch.zeros(ei.shape[-1]) + i)
        ptr += ei.shape[-1]
    return ntypes, etypes, x, torch.hstack(eis), node_types, torch.hstack(ets), nptr, eptr, pads


# --------------------------------------------------------------------------
# L4  orchestration                                              explainer.py:316-546


def explain(feat, edge_index, arch, params, names, pathways=None, pathway_names=None,
            element_type=None, problem="node_prediction", element=None, times=1, mt=None,
            node_types=None, edge_types=None):
    """Port of ``Explainer.run`` for node problems.  Returns a dict of every intermediate.

    ``mt``: an MT19937 positioned where torch's CPU generator would be; defaults to the
    ``set_seed`` origin when ``times == 1`` (explainer.py:342-343).
    """
    import pandas as pd

    from torch_geometric.utils.subgraph import get_num_hops  # stand-in, see oracle/__init__.py

    problem = problem.lower().strip()
    if times == 1 or mt is None:
        mt = seeded_stream(params["seed"])
    ntn = etn = ntypes = etypes = nptr = pads = None
    if isinstance(feat, dict):
        ntn, etn, feat, edge_index, ntypes, etypes, nptr, _eptr, pads = flatten_hetero(feat, edge_index)
    if isinstance(names, dict):
        names = list(itertools.chain.from_iterable(names.values()))
    ptypes = None
    if pathways is not None and isinstance(pathways, dict):
        keys = list(pathways.keys())
        first = pathways[keys[0]][0][0]
        if isinstance(first, (int, float)) and problem == "node":  # pathways.py:204-213 (exact match!)
            for key, ptr in zip(keys, nptr):
                for i in range(len(pathways[key])):
                    pathways[key][i] = (np.array(pathways[key][i]) + ptr).tolist()
        flat_p, flat_n = [], []
        for key in keys:
            flat_p.extend(pathways[key])
            flat_n.extend(pathway_names[key])
        pathways, pathway_names = flat_p, flat_n
    if pathways is not None and pathway_names is None:
        pathway_names = list(range(len(pathways)))

    if ntypes is None and node_types is not None:  # explainer.py:365-371: caller-provided type vectors
        ntypes = node_types.clone()
    if etypes is None and edge_types is not None:
        etypes = edge_types.clone()
    assert "edge" not in problem, "the reference itself raises on edge problems of homogeneous graphs (data.py:359)"
    hops = get_num_hops(arch)
    if etn is not None:
        hops //= len(etn)
    ind = int(np.where(np.array(names, dtype=str) == element)[0][0])
    ei_np = edge_index.numpy() if isinstance(edge_index, torch.Tensor) else np.asarray(edge_index)
    if "graph" in problem:  # explainer.py:427-447: no computational-graph cut, every community kept as given
        subset, sub_ei, sub_ind = np.arange(feat.shape[0]), ei_np.astype(np.int64), ind
        edge_mask = np.ones(ei_np.shape[1], dtype=bool)
    else:
        subset, sub_ei, sub_ind, edge_mask = k_hop_subgraph(ei_np, ind, hops + 1)
    sub_feat = feat[torch.from_numpy(subset)]
    sub_names = np.array(names, dtype=str)[subset].tolist()
    sub_nt = ntypes[torch.from_numpy(subset)] if ntypes is not None else None
    sub_et = etypes[torch.from_numpy(np.flatnonzero(edge_mask))] if etypes is not None else None
    sub_p = sub_pn = sub_pi = None
    if pathways is not None:
        if "graph" in problem:
            sub_p, sub_pn = pathways, pathway_names
        else:
            sub_p, sub_pn, _ = communities_in_subgraph(pathways, pathway_names, sub_names)
        sub_pi = names_to_indices(sub_p, sub_names) if isinstance(sub_p[0][0], str) else sub_p  # explainer.py:467-470
    if "graph" in problem:
        pass  # explainer.py:451: no per-type index filtering for graph problems
    elif element_type is not None:  # explainer.py:451-463
        t = ntn.index(element_type)
        filt = np.array(sub_names, dtype=str)[(sub_nt == t).numpy()].tolist()
        sub_ind = int(np.where(np.array(filt, dtype=str) == element)[0][0])
    elif node_types is not None or edge_types is not None:  # explainer.py:280-284: "that element with node type 1"
        filt = np.array(sub_names, dtype=str)[(sub_nt == 1).numpy()].tolist()
        sub_ind = int(np.where(np.array(filt, dtype=str) == element)[0][0])

    n_sub = sub_feat.shape[0]
    runs = []
    for _ in range(times):
        mask, prow, bsz = mask_generator(n_sub, sub_pi, params, mt)
        w0 = mt.linear_init(n_sub)  # explainer.py:497 -> wlm.py:45
        mt.dataloader_iter()  # wlm.py:210
        batches = []
        for s in range(0, mask.shape[0], bsz):
            mb = mask[s : s + bsz]
            kern, y = kernel_output(mb, sub_feat, sub_ei, arch, sub_ind, sub_nt, sub_et, ntn, etn, pads)
            batches.append((mb, kern, y))
        w, losses = train_wlm(batches, w0, params)
        runs.append(dict(mask=mask, pathway_rows=prow, batch_size=bsz, w0=w0, batches=batches,
                         weights=w.numpy(), losses=losses))
    mean, std = weight_stacking([r["weights"] for r in runs])
    cfg = pd.DataFrame({"name": sub_names, "config_value_mean": mean.numpy(),
                        "config_value_std": std.numpy()}).set_index("name")
    cfg = cfg.sort_values(by=["config_value_mean"], ascending=False)
    pdf = None
    if pathways is not None:
        pdf = pd.DataFrame({"name": sub_pn, "score": aggregate(mean, sub_pi)}).set_index("name")
        pdf = pdf.sort_values(by=["score"], ascending=False).dropna()
    return dict(subset=subset, sub_edge_index=sub_ei, sub_ind=sub_ind, edge_mask=edge_mask,
                sub_names=sub_names, sub_pathway_inds=sub_pi, sub_pathway_names=sub_pn,
                runs=runs, mean=mean.numpy(), std=std.numpy(), config_val_df=cfg, pathway_df=pdf)
