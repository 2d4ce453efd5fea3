"""``Explainer(...).run(query_node, repeats)`` -- the drop-in boundary.

Same constructor, assertions, side effects and return values as the reference's
``pathway_explanations.explainer.Explainer`` (``explainer.py:25-546``), with the perturbation hot
path (k-hop cut, coalition masks, perturbed forward, SHAP weights, surrogate fit, community
scores) running in CUDA kernels through ``libxpgnn_b200.so``.  There is no CPU fallback: without
a CUDA device or without the built library, ``run`` raises.
"""
import random

import numpy as np
import torch

from . import _lib
from .data import Data
from .engine import GraphSpec, MaskedForward, require_cuda
from .kernels import shap_weights
from .lowering import lower
from .masks import Mask
from .model import Model
from .pathways import Pathways, all_str
from .shard import sharded_eval
from .wlm import LinearRegression, fit_surrogate


def set_seed(seed=100):
    """explainer.py:14-22 (the CPU generator seeded with ``seed + 2`` is the stream that matters)."""
    random.seed(seed)
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 2)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed + 3)
        torch.cuda.manual_seed_all(seed + 4)
    torch.backends.cudnn.enabled = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def build_engine(feat, edge_index, arch, element_index, node_type=None, edge_type=None, node_type_names=None,
                 edge_type_names=None, padded_dims=None, out_type=None, hop=None, prune=False, precision="fp32",
                 query_flat=None, model=None, queries=None):
    """Lower ``arch`` and bind it to the (flattened) computational graph.

    ``element_index`` follows the reference: the row of the model output that is read
    (``model.py:247,325``), i.e. an index *inside the output node type* for hetero graphs.
    ``model``: an already lowered plan (``lowering.lower_candidates``); ``queries``: flat node ids to read instead of
    the single element (probe check)."""
    dev = require_cuda()
    if model is None:
        model = lower(arch)
    n = int(feat.shape[0])
    if node_type_names is None:
        if model.hetero:
            raise NotImplementedError("a HeteroConv model needs dict inputs (node / edge types)")
        graph = GraphSpec(feat, edge_index, [0, n])
        n_types = 1
        q = int(element_index) if query_flat is None else int(query_flat)
    else:
        nt = node_type.to(torch.int64).cpu()
        if nt.numel() > 1 and bool((nt[1:] < nt[:-1]).any()):
            raise ValueError("node types must be contiguous blocks (flattened hetero graph)")
        counts = torch.bincount(nt, minlength=len(node_type_names)).tolist()
        type_ptr = [0]
        for c in counts:
            type_ptr.append(type_ptr[-1] + int(c))
        graph = GraphSpec(feat, edge_index, type_ptr, list(node_type_names), edge_type.to(torch.int64),
                          [tuple(e) for e in edge_type_names])
        n_types = int((torch.tensor(counts) > 0).sum())
        if query_flat is not None:
            q = int(query_flat)
        else:
            t = out_type if out_type is not None else getattr(arch, "out_type", None) or node_type_names[0]
            q = type_ptr[node_type_names.index(t)] + int(element_index)
    if model.out_dim != 1:
        raise NotImplementedError("the surrogate needs a scalar prediction per node (model output width 1)")
    multi = n_types >= 2
    eng = MaskedForward(graph, model, [q] if queries is None else list(queries), prune=prune and hop is not None, hop=hop,
                        zero_edge_rule=multi and queries is None, precision=precision)
    # (B,1) targets broadcast against (B,) predictions in the reference loss (wlm.py:517); the
    # multi-node-type branch yields (B,) targets (model.py:251) and the plain weighted MSE
    eng.broadcast_y = not multi
    return eng


def _forward_arity(arch):
    import inspect

    return len(inspect.getfullargspec(arch.forward).args)  # model.py:104: 3 = (self, x, edge_index), 5 = + node / edge types


def verify_lowering(arch, feat_dims, node_type_names=None, edge_type_names=None, out_type=None, seed=12345):
    """Pick the lowered plan that IS ``arch``: the reference calls ``arch(feat, edge_index)`` as a black box
    (``model.py:104-112``), the engine runs a lowered plan, so every candidate plan (``lowering.lower_candidates``) is
    evaluated by the engine on a small random probe graph with no node removed and compared with ``arch`` itself on the
    same graph, all output rows.  Returns ``(plan, status)``, status = "verified" or "unverified: ..." when ``arch``
    cannot be called (weight containers of ``nn.py``: their plan is their definition).  Raises ``NotImplementedError``
    when no candidate reproduces ``arch`` -- a model with skip connections, functional ops the lowering does not know, or
    layers applied in an order the plan does not reflect must not be explained with a different function."""
    from .lowering import lower_candidates

    cands = lower_candidates(arch)
    g = torch.Generator().manual_seed(seed)
    par = next(arch.parameters(), None)
    adev = par.device if par is not None else torch.device("cpu")
    hetero = node_type_names is not None
    if not hetero:
        n, e = 48, 220
        x = torch.randn(n, int(feat_dims), generator=g)
        ei = torch.randint(0, n, (2, e), generator=g)
        ei = torch.cat([ei, ei[:, :7], torch.arange(5).repeat(2, 1)], 1)  # duplicate edges and self loops
        args = (x.to(adev), ei.to(adev))
        if _forward_arity(arch) == 5:
            args = args + (torch.zeros(n, device=adev), torch.zeros(ei.shape[1], device=adev))
    else:
        counts = [20 + 3 * i for i in range(len(node_type_names))]
        x_dict = {t: torch.randn(c, int(feat_dims[t]), generator=g) for t, c in zip(node_type_names, counts)}
        ei_dict = {}
        for r in edge_type_names:
            r = tuple(r)
            ns, nd = counts[node_type_names.index(r[0])], counts[node_type_names.index(r[-1])]
            ei_dict[r] = torch.stack([torch.randint(0, ns, (60,), generator=g), torch.randint(0, nd, (60,), generator=g)])
        args = ({t: v.to(adev) for t, v in x_dict.items()}, {r: v.to(adev) for r, v in ei_dict.items()})
    was_training = arch.training
    arch.eval()
    try:
        with torch.no_grad():
            ref = arch(*args)
    except NotImplementedError as ex:
        # nn.py layers hold weights only, and a Module without a forward (torch: 'missing the required "forward" function')
        # cannot be called either: there is no black box to compare with, the plan is the model's definition
        if "weight container" in str(ex) or 'missing the required "forward"' in str(ex):
            return cands[-1], "unverified: arch cannot be called (weight containers / no forward), its plan is its definition"
        raise
    finally:
        arch.train(was_training)
    # rows of the reference output <-> flat probe node ids
    if hetero:
        raw = Data(x_dict, ei_dict)
        ntn, etn, feat, ei, node_types, edge_types, nptr, _eptr, pads = raw.preprocess_hetero_graph()
        if isinstance(ref, dict):  # model.py:255-292: outputs of the node types, stacked in dict order
            rows, flat = [], []
            for t, v in ref.items():
                rows.append(v)
                lo = nptr[ntn.index(t)]
                flat += list(range(lo, lo + v.shape[0]))
            ref = torch.vstack(rows)
        else:
            t = out_type if out_type is not None else getattr(arch, "out_type", None) or ntn[0]
            lo = nptr[ntn.index(t)]
            flat = list(range(lo, lo + ref.shape[0]))
    else:
        ntn = etn = node_types = edge_types = pads = None
        feat, ei = x, ei
        flat = list(range(ref.shape[0]))
    ref = ref.detach().float().cpu().reshape(len(flat), -1)[:, 0].numpy()
    dev = require_cuda()
    act = torch.ones((int(feat.shape[0]), 1), dtype=torch.int32, device=dev)
    worst = []
    for cand in cands:
        try:
            eng = build_engine(feat.to(dev), ei.to(dev), arch, 0, node_types, edge_types, ntn, etn, pads, out_type=out_type,
                               model=cand, queries=flat)
        except NotImplementedError as ex:
            worst.append("%s: %s" % (cand.how, ex))
            continue
        y = eng(act, 1)[0].cpu().numpy()
        err = float(np.max(np.abs(y - ref) / np.maximum(np.abs(ref), 1e-3)))
        if err <= 2e-4:
            return cand, "verified"
        worst.append("%s: max rel. deviation %.3g on the probe graph" % (cand.how, err))
    raise NotImplementedError(
        "arch.forward is not the layer sequence the engine can lower (%s); supported: chains of GCNConv / SAGEConv(mean) / "
        "HeteroConv(sum) with ReLU / Sigmoid, then Linear layers" % "; ".join(worst))


def sync_rng_across_ranks():
    """One process per GPU: every rank replays the coalition stream of the CPU generator, so the generators must agree.
    ``set_seed`` guarantees that only for ``times == 1`` (explainer.py:342); rank 0's state is made authoritative."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    box = [torch.get_rng_state()]
    dist.broadcast_object_list(box, src=0)
    torch.set_rng_state(box[0])


def check_ranks_agree(act):
    """Ranks evaluate slices of ONE coalition matrix; raise if theirs differ (mismatched masks would pair predictions
    with the wrong rows and the replicated fit would silently train on wrong targets)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    a = act.view(-1)
    n = int(a.numel())
    wcol = (torch.arange(int(act.shape[1]), device=act.device, dtype=torch.int64) % 8191) + 1
    h = torch.stack([torch.sum(a, dtype=torch.int64), torch.sum(torch.sum(act, dim=0, dtype=torch.int64) * wcol),
                     torch.sum(act[:: max(n // (1 << 20) // max(int(act.shape[1]), 1), 1)].to(torch.int64) * 31 % 1000003),
                     torch.tensor(n, device=act.device, dtype=torch.int64)])
    lo, hi = h.clone(), h.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        raise RuntimeError("ranks generated different coalition masks (per-rank seeding?): Explainer.run needs the same "
                           "global CPU generator state on every rank")


class Explainer:
    #: engine knobs (extension).  precision "fp32": fp32 storage and accumulation, dense transforms as 3xTF32 tensor-core
    #: products (error ~1e-6 relative, inside the 1e-4 bar); "fp32_exact": exact fp32 FMA transforms (slower, for audits);
    #: "bf16" / "bf16_act": the 2e-2 bar.  verify: compare the
    #: lowered plan with ``arch`` itself on a probe graph before explaining (see ``verify_lowering``).
    engine_options = dict(prune=True, precision="fp32", verify=True)

    def __init__(self, feat, edge_index, arch, params, names, pathways=None, pathway_names=None, element_type=None,
                 problem="node_prediction", node_types=None, edge_types=None):
        self.initial_assertions(feat, edge_index, arch, params, names, pathways, pathway_names, element_type, problem)
        problem = problem.lower().strip()
        self.feat, self.edge_index, self.arch, self.params, self.names = feat, edge_index, arch, params, names
        self.pathways, self.pathway_names, self.element_type = pathways, pathway_names, element_type
        self.problem, self.node_types, self.edge_types = problem, node_types, edge_types
        self.last_stats = {}

    @staticmethod
    def initial_assertions(feat, edge_index, arch, params, names, pathways, pathway_names, element_type, problem):
        """explainer.py:106-189, same messages."""
        if pathways is not None:
            assert isinstance(pathways, list) or isinstance(pathways, dict), "Pathways is not list or dict"
        if pathway_names is not None:
            assert isinstance(pathway_names, list) or isinstance(
                pathway_names, dict
            ), "Pathway names is not list or dict"
            assert len(pathway_names) == len(
                pathways
            ), "Length of list with pathway names and list with pathway indexes do not match"
        assert isinstance(feat, torch.Tensor) or isinstance(feat, dict), "Feature matrix is not torch tensor or dict"
        assert isinstance(edge_index, torch.Tensor) or isinstance(
            edge_index, dict
        ), "Edge index matrix is not torch tensor or dict"
        assert isinstance(names, list) or isinstance(names, dict), "Element names is not list or dict"
        assert isinstance(params, dict), "Hyperparameters given is not dictionary"
        assert isinstance(problem, str), "Problem type given is not string"
        if element_type is not None:
            assert isinstance(element_type, str) or isinstance(
                element_type, tuple
            ), "Element type is not string (node) nor tuple (edge)"
            if "node" in problem:
                assert isinstance(feat, dict), "Feature given is not a dict of node types"
                assert (
                    element_type in list(feat.keys())
                ), "Node type '{}' is not among input node types in heterogeneous graph".format(element_type)
            elif "edge" in problem:
                assert isinstance(edge_index, dict), "Edge index given is not a dict of edge index types"
                assert (
                    element_type in list(edge_index.keys())
                ), "Edge type '{}' is not among input node types in heterogeneous graph".format(element_type)

    @staticmethod
    def extract_index(element, names=None):
        """explainer.py:191-226."""
        if names is None:
            assert isinstance(element, int) or isinstance(
                element, float
            ), "No element names have been given and the node name given is not numeric"
            return int(element)
        assert element in names, "Element name '{}' is not present in the graph".format(element)
        if type(element) is str and all_str(names):
            return names.index(element)  # first match, as np.where(...)[0][0]
        return int(np.where(np.array(names, dtype=str) == element)[0][0])

    def filter_hetero_names(self, names, node_type, edge_type, node_type_names, edge_type_names):
        """explainer.py:228-286."""
        names_array = np.array(names, dtype=str)
        if isinstance(self.element_type, str):
            sel = torch.where(node_type == node_type_names.index(self.element_type))[0]
        elif isinstance(self.element_type, tuple):
            sel = torch.where(edge_type == edge_type_names.index(self.element_type))[0]
        else:
            sel = torch.where(node_type == 1)[0]
        return names_array[sel.cpu().numpy()].tolist()

    @staticmethod
    def weight_stacking(weights):
        """explainer.py:288-314: mean and population std over repeats (device kernel)."""
        lib = _lib.load()
        stack = torch.vstack([w.reshape(1, -1) for w in weights]).to(require_cuda(), torch.float32).contiguous()
        t, n = stack.shape
        mean, std = torch.empty(n, device=stack.device), torch.empty(n, device=stack.device)
        _lib.check(lib.xpgnn_repeat_stats(stack.data_ptr(), t, n, mean.data_ptr(), std.data_ptr(), _lib.stream_ptr()))
        return mean, std

    def run(self, element, times=1):
        """explainer.py:316-546.  Returns (config_val_df, pathway_df)."""
        import gc

        # The host side handles lists of up to millions of names; every generation-2 collection of the cyclic GC walks all of
        # them (measured: up to 0.4 s per explained query at 1 M names).  Nothing here creates reference cycles.
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._run(element, times)
        finally:
            if gc_was_enabled:
                gc.enable()

    def _run(self, element, times):
        dev = require_cuda()
        _lib.load()
        if "edge" in self.problem:
            # the reference itself cannot run them: comp_graph raises UnboundLocalError on homogeneous graphs (data.py:359,
            # ``ind_filter``), Mask reads an attribute that is never set (masks.py:294, ``self.edge_size``)
            raise NotImplementedError("edge problems: the reference path is not runnable (data.py:359, masks.py:294); not built")
        graph_problem = "graph" in self.problem
        if times == 1:
            set_seed(self.params["seed"])
        nvtx = torch.cuda.nvtx
        nvtx.range_push("xpgnn:run:flatten")
        raw = Data(self.feat, self.edge_index)
        if self.pathways is not None:
            raw_pathways = Pathways(self.pathways, self.pathway_names)
        (ntn, etn, self.feat, self.edge_index, node_types, edge_types, nptr, eptr, pads) = raw.preprocess_hetero_graph()
        typed_homo = False
        if node_types is None and self.node_types is not None:  # explainer.py:365-371: caller-provided type vectors of a
            node_types = self.node_types.clone()                 # homogenised graph (5-argument forward, model.py:110-112)
            typed_homo = True
        if edge_types is None and self.edge_types is not None:
            edge_types = self.edge_types.clone()
            typed_homo = True
        self.names, _ = raw.hetero2homo_names(self.names)
        if self.pathways is not None:
            self.pathways, self.pathway_names, ptypes = raw_pathways.hetero2homo(self.problem, nptr, eptr)
            pathway_class = Pathways(self.pathways, self.pathway_names, ptypes)

        nvtx.range_pop()
        nvtx.range_push("xpgnn:run:khop")
        self.feat = self.feat.to(dev)
        self.edge_index = self.edge_index.to(dev)
        data_class = Data(self.feat, self.edge_index)
        relations = len(etn) if etn is not None else 0
        n_hops = Model(self.arch).get_hops(relations)
        ind = self.extract_index(element, self.names)
        if graph_problem:
            # explainer.py:427-447: no computational-graph cut -- the whole graph, every community as given, the prediction
            # read at the element's own row (wlm.py:435-436); every conv layer runs over the whole graph ("full" mode)
            sub_feat, sub_ei, sub_names = self.feat, self.edge_index.to(torch.int64), self.names
            sub_ind = torch.tensor([ind])
            sub_nt = None if node_types is None else node_types.to(dev)
            sub_et = None if edge_types is None else edge_types.to(dev)
            hop = None
        else:
            sub_feat, sub_ei, sub_names, sub_ind, sub_nt, sub_et = data_class.comp_graph(
                ind, n_hops, self.problem, self.names, node_types, edge_types)
            hop = data_class.last_hop
        query_flat = hop_query = int(sub_ind[0])  # hop levels are distances to this node

        nvtx.range_pop()
        nvtx.range_push("xpgnn:run:communities")
        sub_pathway_inds = sub_pathway_names = None
        if self.pathways is not None and graph_problem:  # explainer.py:447-449, 464-470: communities kept as given
            sub_pathway_names = self.pathway_names
            first = self.pathways[0][0]
            sub_pathway_inds = Pathways(self.pathways, sub_pathway_names).names2inds(sub_names) if isinstance(first, str) else self.pathways
            sub_pathway_class = Pathways(sub_pathway_inds, sub_pathway_names)
        elif self.pathways is not None:
            # pathways.py:33-136 (comp_graph + names2inds) in one vectorised pass; see Pathways.resolve_indices
            sub_pathway_inds, sub_pathway_names = pathway_class.resolve_indices(sub_names)
            if not sub_pathway_inds:
                raise IndexError("list index out of range")  # explainer.py:467: no community reaches the computational graph
            sub_pathway_class = Pathways(sub_pathway_inds, sub_pathway_names)
        if not graph_problem and (self.element_type is not None or self.node_types is not None or self.edge_types is not None):
            filtered = self.filter_hetero_names(sub_names, sub_nt, sub_et, ntn, etn)
            sub_ind = torch.tensor([self.extract_index(element, filtered)])
        del self.feat, self.edge_index  # explainer.py:476: the object is single use

        nvtx.range_pop()
        nvtx.range_push("xpgnn:run:engine_build")
        elements = int(sub_feat.shape[0])
        opts = dict(type(self).engine_options)
        opts.update(getattr(self, "options", {}))
        out_type = self.element_type if isinstance(self.element_type, str) else None
        self.arch.eval()
        plan = None
        self.lowering_status = "not checked"
        if opts.get("verify", True):  # the plan must be the function the reference would call (model.py:104-112)
            fd = int(sub_feat.shape[1]) if ntn is None else {t: int(sub_feat.shape[1]) - int(pads[i]) for i, t in enumerate(ntn)}
            plan, self.lowering_status = verify_lowering(self.arch, fd, ntn, etn, out_type)
        if typed_homo:
            # The reference hands the type vectors to arch.forward and then reads row ``sub_ind`` -- the query's index among
            # the type-1 nodes (explainer.py:280-284) -- of the full output (wlm.py:435-436).  The engine lowers models whose
            # plan reproduces arch on the probe graph (i.e. that do not branch on the types) and reads the same row.
            sub_nt = sub_et = None
            query_flat = int(sub_ind[0])
        engine = build_engine(sub_feat, sub_ei, self.arch, int(sub_ind[0]), sub_nt, sub_et, ntn, etn, pads,
                              out_type=out_type, hop=hop if not typed_homo or query_flat == int(hop_query) else None,
                              prune=opts["prune"], precision=opts["precision"], query_flat=query_flat, model=plan)
        # ---- all repeats' coalitions first: the masks, the surrogate init and the DataLoader seed draw of repeat r + 1
        # depend on the RNG stream only (masks -> randperm -> N init draws -> 2 draws, explainer.py:490-523 / wlm.py:210),
        # never on the predictions, so every repeat's rows go through ONE sharded engine call ----
        nvtx.range_pop()
        nvtx.range_push("xpgnn:run:masks")
        sync_rng_across_ranks()
        sets = []
        for _ in range(times):
            coalitions, _rows = Mask(sub_feat, sub_ei, sub_pathway_inds, self.params, self.problem).mask_generator()
            wlrm = LinearRegression(elements)           # kaiming-uniform init: next N draws of the CPU stream
            torch.empty((), dtype=torch.int64).random_()  # iter(DataLoader) base seed (wlm.py:210)
            sets.append((coalitions, wlrm.layer.weight.detach().reshape(-1)))
        if times == 1:
            act_all, n_all = sets[0][0].act, sets[0][0].n_coalitions
        else:  # repeats side by side along the word axis; a repeat's last word is padded with all-off rows
            act_all = torch.cat([c.act for c, _ in sets], dim=1).contiguous()
            n_all = 32 * int(act_all.shape[1])
        check_ranks_agree(act_all)
        nvtx.range_pop()
        nvtx.range_push("xpgnn:run:masked_forward")
        y_all = sharded_eval(engine, act_all, n_all)[:, 0]
        nvtx.range_pop()
        nvtx.range_push("xpgnn:run:fit")
        config_vals, word0 = [], 0
        for coalitions, w0 in sets:
            y = y_all[32 * word0: 32 * word0 + coalitions.n_coalitions]
            word0 += coalitions.words
            kern = shap_weights(coalitions.popcount, elements, coalitions.batch_size)
            w, losses = fit_surrogate(coalitions, y, kern, w0, self.params, broadcast_y=engine.broadcast_y, want_losses=False)
            config_vals.append(w)
            self.last_stats = dict(n_sub=elements, e_sub=int(sub_ei.shape[1]), coalitions=coalitions.n_coalitions,
                                   batch_size=coalitions.batch_size, tile_coalitions=engine.tile_coalitions,
                                   repeats_per_engine_call=times, lowering=self.lowering_status)
            if opts.get("keep_last", False):  # tests / debugging only: pins the coalition draws and predictions in HBM
                self._last = dict(coalitions=coalitions, y=y, kernel=kern, w0=w0, subset_names=sub_names,
                                  sub_edge_index=sub_ei, sub_ind=query_flat)
        nvtx.range_pop()
        nvtx.range_push("xpgnn:run:aggregate")
        mean, std = self.weight_stacking(config_vals)
        config_val_df = Data.config_val_dataframe(mean, std, sub_names)
        pathway_df = None
        if self.pathways is not None:
            pathway_df = sub_pathway_class.aggregate(mean, sub_pathway_inds)
        nvtx.range_pop()
        return config_val_df, pathway_df
