"""Graph data operations behind the reference's ``Data`` interface (``data.py:19-878``).

Only what sits on the perturbation path is provided.  ``comp_graph`` runs the frontier-BFS k-hop
kernel; the hetero <-> homo flattening is tensor plumbing (cat / pad on the device tensors, as in
the reference).  The reference's ``perturbator`` / ``build_edge_mask`` / ``perturb_node`` /
``concat_features`` materialise a block-diagonal batch: the engine never does that, so those
entry points raise ``NotImplementedError`` pointing at ``engine.MaskedForward``.
"""
import itertools
import operator

import numpy as np
import pandas as pd
import torch

from . import _lib
from .engine import require_cuda
from .pathways import all_str


def khop_subgraph(edge_index, n_nodes, query, hops):
    """Device k-hop (``xpgnn_khop_subgraph``).  Returns subset (int64), sub_edge_index (2,E_sub) int64,
    rank of the query, edge_mask (bool, E), hop levels (int8, N; -1 outside), relabel (int32, N)."""
    lib = _lib.load()
    dev = require_cuda()
    ei = edge_index.to(dev, torch.int64).contiguous()
    e = int(ei.shape[1])
    subset = torch.empty(n_nodes, dtype=torch.int64, device=dev)
    relabel = torch.empty(n_nodes, dtype=torch.int32, device=dev)
    hop = torch.empty(n_nodes, dtype=torch.int8, device=dev)
    edge_mask = torch.zeros(max(e, 1), dtype=torch.uint8, device=dev)
    sub_ei = torch.empty((2, max(e, 1)), dtype=torch.int64, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    _lib.check(lib.xpgnn_khop_subgraph(ei.data_ptr(), e, int(n_nodes), int(query), int(hops), subset.data_ptr(),
                                       relabel.data_ptr(), hop.data_ptr(), edge_mask.data_ptr(), sub_ei.data_ptr(),
                                       counts.data_ptr(), _lib.stream_ptr()))
    n_sub, e_sub, n_bad = (int(v) for v in counts.tolist())
    if n_bad:
        raise IndexError("edge_index holds %d edge(s) with an endpoint outside [0, %d)" % (n_bad, n_nodes))
    subset = subset[:n_sub]
    sub_ei = sub_ei[:, :e_sub].contiguous()
    sub_ind = relabel[int(query)].to(torch.int64).reshape(1)
    return subset, sub_ei, sub_ind, edge_mask[:e].bool(), hop, relabel


class Data:
    def __init__(self, feat, edge_index):
        self.feat = feat
        self.edge_index = edge_index

    # ------------------------------------------------------------------ hetero -> homo (data.py:39-147, 695-878)
    def preprocess_hetero_graph(self):
        if isinstance(self.edge_index, dict) and isinstance(self.feat, dict):
            ntypes, etypes = list(self.feat.keys()), list(self.edge_index.keys())
            feat_homo, ei_homo, node_types, edge_types, nptr, eptr, pads = self.hetero2homo()
            return ntypes, etypes, feat_homo, ei_homo, node_types, edge_types, nptr, eptr, pads
        return None, None, self.feat, self.edge_index, None, None, None, None, None

    def hetero2homo(self):
        feat_homo, node_types, pads, nptr = self.concatenate_hetero_features()
        ei_homo, edge_types, eptr = self.concatenate_hetero_edge_indices(nptr)
        return feat_homo, ei_homo, node_types, edge_types, nptr, eptr, pads

    def concatenate_hetero_features(self):
        tensors = list(self.feat.values())
        padded, pads, ptrs = pad_feat_tensors(tensors)
        dev = tensors[0].device
        node_types = torch.hstack([torch.zeros(p.shape[0], device=dev) + i for i, p in enumerate(padded)])
        return torch.vstack(padded), node_types, pads, ptrs

    def concatenate_hetero_edge_indices(self, node_pointers):
        names = list(self.feat.keys())
        mapped, types, ptrs, ptr = [], [], [], 0
        for i, (rel, ei) in enumerate(self.edge_index.items()):
            ptrs.append(ptr)
            add = torch.tensor([[node_pointers[names.index(rel[0])]], [node_pointers[names.index(rel[-1])]]],
                               device=ei.device)
            mapped.append(ei + add)
            types.append(torch.zeros(ei.shape[-1], device=ei.device) + i)
            ptr += ei.shape[-1]
        return torch.hstack(mapped), torch.hstack(types), ptrs

    @staticmethod
    def hetero2homo_names(names):
        if isinstance(names, dict):
            lists = list(names.values())
            homo = list(itertools.chain.from_iterable(lists))
            types = torch.cat([torch.zeros(len(l)) + i for i, l in enumerate(lists)])
            return homo, types
        return names, None

    # ------------------------------------------------------------------ k-hop (data.py:281-361)
    def comp_graph(self, ind, n_hops, problem, names, node_types=None, edge_types=None):
        n_hops += 1  # data.py:328: one more hop than the model covers
        n = int(self.feat.shape[0])
        subset, sub_ei, sub_ind, edge_mask, hop, _ = khop_subgraph(self.edge_index, n, ind, n_hops)
        self.last_hop = hop[subset]
        sub_feat = self.feat.to(subset.device)[subset]
        sub_nt = node_types.to(subset.device)[subset] if node_types is not None else None
        sub_et = edge_types.to(subset.device)[edge_mask] if edge_types is not None else None
        if ("node" in problem) or ("graph" in problem):
            idx = subset.cpu().tolist()
            if all_str(names):
                sub_names = list(operator.itemgetter(*idx)(names)) if len(idx) > 1 else [names[i] for i in idx]
                # == np.array(names, dtype=str)[subset].tolist(), without the N-string array
            else:
                sub_names = np.array(names, dtype=str)[idx].tolist()
        else:
            raise NotImplementedError("edge problems: the reference raises here as well (ind_filter is undefined for "
                                      "homogeneous graphs, data.py:359)")
        return sub_feat, sub_ei, sub_names, sub_ind, sub_nt, sub_et

    def element_size(self, problem):
        return self.edge_index.shape[1] if "edge" in problem else self.feat.shape[0]

    # ------------------------------------------------------------------ materialising entry points
    def perturbator(self, *a, **k):
        raise NotImplementedError("the block-diagonal batch of data.py:591-648 is never materialised; "
                                  "use engine.MaskedForward on the packed coalition bits")

    build_edge_mask = perturb_node = perturb_edge = concat_features = perturbator

    @staticmethod
    def config_val_dataframe(config_val_mean, config_val_std, names):
        df = pd.DataFrame({"name": names,
                           "config_value_mean": config_val_mean.cpu().detach().numpy(),
                           "config_value_std": config_val_std.cpu().detach().numpy()}).set_index("name")
        return df.sort_values(by=["config_value_mean"], ascending=False)


def pad_feat_tensors(feat_tensors):
    """data.py:825-878: zero-pad every node type to the widest feature dimension."""
    fmax = max(t.shape[1] for t in feat_tensors)
    padded, pads, ptrs, ptr = [], [], [], 0
    for t in feat_tensors:
        d = fmax - t.shape[1]
        pads.append(d)
        ptrs.append(ptr)
        ptr += t.shape[0]
        padded.append(torch.nn.functional.pad(t, (0, d)) if d > 0 else t)
    return padded, pads, ptrs
