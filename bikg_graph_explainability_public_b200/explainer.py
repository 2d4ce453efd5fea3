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
                 query_flat=None):
    """Lower ``arch`` and bind it to the (flattened) computational graph.

    ``element_index`` follows the reference: the row of the model output that is read
    (``model.py:247,325``), i.e. an index *inside the output node type* for hetero graphs."""
    dev = require_cuda()
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
    eng = MaskedForward(graph, model, [q], prune=prune and hop is not None, hop=hop, zero_edge_rule=multi,
                        precision=precision)
    # (B,1) targets broadcast against (B,) predictions in the reference loss (wlm.py:517); the
    # multi-node-type branch yields (B,) targets (model.py:251) and the plain weighted MSE
    eng.broadcast_y = not multi
    return eng


class Explainer:
    #: engine knobs (extension; defaults give reference-identical results)
    engine_options = dict(prune=True, precision="fp32")

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
        dev = require_cuda()
        _lib.load()
        if "graph" in self.problem or "edge" in self.problem:
            raise NotImplementedError("graph / edge problems are outside the accelerated path (SURVEY.md 8f-4)")
        if times == 1:
            set_seed(self.params["seed"])
        raw = Data(self.feat, self.edge_index)
        if self.pathways is not None:
            raw_pathways = Pathways(self.pathways, self.pathway_names)
        (ntn, etn, self.feat, self.edge_index, node_types, edge_types, nptr, eptr, pads) = raw.preprocess_hetero_graph()
        if node_types is None and self.node_types is not None:
            raise NotImplementedError("custom node_types on a homogeneous graph (5-arg forward) is not lowered")
        self.names, _ = raw.hetero2homo_names(self.names)
        if self.pathways is not None:
            self.pathways, self.pathway_names, ptypes = raw_pathways.hetero2homo(self.problem, nptr, eptr)
            pathway_class = Pathways(self.pathways, self.pathway_names, ptypes)

        self.feat = self.feat.to(dev)
        self.edge_index = self.edge_index.to(dev)
        data_class = Data(self.feat, self.edge_index)
        relations = len(etn) if etn is not None else 0
        n_hops = Model(self.arch).get_hops(relations)
        ind = self.extract_index(element, self.names)
        sub_feat, sub_ei, sub_names, sub_ind, sub_nt, sub_et = data_class.comp_graph(
            ind, n_hops, self.problem, self.names, node_types, edge_types)
        hop = data_class.last_hop
        query_flat = int(sub_ind[0])

        sub_pathway = sub_pathway_names = None
        if self.pathways is not None:
            sub_pathway, sub_pathway_names, _ = pathway_class.comp_graph(sub_names)
        if self.element_type is not None or self.node_types is not None or self.edge_types is not None:
            filtered = self.filter_hetero_names(sub_names, sub_nt, sub_et, ntn, etn)
            sub_ind = torch.tensor([self.extract_index(element, filtered)])
        sub_pathway_inds = None
        if self.pathways is not None:
            sub_pathway_class = Pathways(sub_pathway, sub_pathway_names)
            if isinstance(sub_pathway[0][0], str):
                sub_pathway_inds = sub_pathway_class.names2inds(sub_names, index=pathway_class.last_index, filtered=True)
            elif isinstance(sub_pathway[0][0], int):
                sub_pathway_inds = sub_pathway
        del self.feat, self.edge_index  # explainer.py:476: the object is single use

        elements = int(sub_feat.shape[0])
        opts = dict(type(self).engine_options)
        opts.update(getattr(self, "options", {}))
        engine = build_engine(sub_feat, sub_ei, self.arch, int(sub_ind[0]), sub_nt, sub_et, ntn, etn, pads,
                              out_type=self.element_type if isinstance(self.element_type, str) else None,
                              hop=hop, prune=opts["prune"], precision=opts["precision"], query_flat=query_flat)
        self.arch.eval()
        config_vals = []
        for _ in range(times):
            coalitions, _rows = Mask(sub_feat, sub_ei, sub_pathway_inds, self.params, self.problem).mask_generator()
            wlrm = LinearRegression(elements)           # kaiming-uniform init: next N draws of the CPU stream
            torch.empty((), dtype=torch.int64).random_()  # iter(DataLoader) base seed (wlm.py:210)
            y = sharded_eval(engine, coalitions.act, coalitions.n_coalitions)[:, 0]
            kern = shap_weights(coalitions.popcount, elements, coalitions.batch_size)
            w, losses = fit_surrogate(coalitions, y, kern, wlrm.layer.weight.detach().reshape(-1), self.params,
                                      broadcast_y=engine.broadcast_y, want_losses=False)
            config_vals.append(w)
            self.last_stats = dict(n_sub=elements, e_sub=int(sub_ei.shape[1]), coalitions=coalitions.n_coalitions,
                                   batch_size=coalitions.batch_size, tile_coalitions=engine.tile_coalitions)
            self._last = dict(coalitions=coalitions, y=y, kernel=kern, w0=wlrm.layer.weight.detach().reshape(-1),
                              subset_names=sub_names, sub_edge_index=sub_ei, sub_ind=query_flat)
        mean, std = self.weight_stacking(config_vals)
        config_val_df = Data.config_val_dataframe(mean, std, sub_names)
        pathway_df = None
        if self.pathways is not None:
            pathway_df = sub_pathway_class.aggregate(mean, sub_pathway_inds)
        return config_val_df, pathway_df
