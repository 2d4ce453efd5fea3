"""ORACLE / TEST INFRASTRUCTURE ONLY.

Black-box model fixtures built from the PyG stand-in layers, shaped like the
reference's own fixtures (``tests/test_utils.py:10-83`` ``GCN_homo``: a ModuleList
``conv`` = [conv, ReLU, ...] and a ModuleList ``fc`` = [Linear, act, ...]) so their
state-dict keys match the reference checkpoints in ``test_data/``.
"""
import torch
from torch import nn

from torch_geometric.nn import GCNConv, HeteroConv, Linear, SAGEConv  # stand-in, see oracle/__init__.py


def _head(dims, final_sigmoid=True):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(Linear(dims[i], dims[i + 1]))
        last = i == len(dims) - 2
        if not last:
            mods.append(nn.ReLU())
        elif final_sigmoid:
            mods.append(nn.Sigmoid())
    return nn.ModuleList(mods)


class HomoGCN(nn.Module):
    """[GCNConv+ReLU]*L then an MLP head; L=1,F=84,hidden 16, head 16-16-32-1 == GCN_homo(84)."""

    def __init__(self, in_dim, conv_dims=(16,), head_dims=(16, 16, 32, 1), final_sigmoid=True, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        mods, d = [], in_dim
        for h in conv_dims:
            mods += [GCNConv(d, h), nn.ReLU()]
            d = h
        self.conv = nn.ModuleList(mods)
        self.fc = _head(head_dims, final_sigmoid)

    def forward(self, x, edge_index):
        for i, c in enumerate(self.conv):
            x = c(x, edge_index) if i % 2 == 0 else c(x)
        for l in self.fc:
            x = l(x)
        return x


class HomoGCN5(HomoGCN):
    """Same stack behind the reference's 5-argument protocol ``forward(x, edge_index, node_types, edge_types)``
    (``model.py:110-112``): a model trained on a homogenised hetero graph that receives the type vectors."""

    def forward(self, x, edge_index, node_types, edge_types):
        assert node_types.shape[0] == x.shape[0] and edge_types.shape[0] == edge_index.shape[1]
        return super().forward(x, edge_index)


class HeteroGCNSingleType(nn.Module):
    """HeteroConv of GCNConv per relation over ONE node type (matches
    ``test_data/gcn_hetero_1hop_lungCancer.pth.tar``)."""

    def __init__(self, in_dim, relations, conv_dims=(16,), head_dims=(16, 16, 32, 1), seed=0):
        super().__init__()
        torch.manual_seed(seed)
        mods, d = [], in_dim
        for h in conv_dims:
            mods += [HeteroConv({r: GCNConv(d, h) for r in relations}, aggr="sum"), nn.ReLU()]
            d = h
        self.conv = nn.ModuleList(mods)
        self.fc = _head(head_dims, True)

    def forward(self, x_dict, edge_index_dict):
        x = x_dict
        for i, c in enumerate(self.conv):
            if i % 2 == 0:
                x = c(x, edge_index_dict)
            else:
                x = {k: c(v) for k, v in x.items()}
        x = x[list(x.keys())[0]]
        for l in self.fc:
            x = l(x)
        return x


class HeteroGCNSingleTypeDict(HeteroGCNSingleType):
    """The dict-returning variant (``model.py:255-292``): ``{node type: output}`` instead of the output tensor."""

    def forward(self, x_dict, edge_index_dict):
        key = list(x_dict.keys())[0]
        return {key: super().forward(x_dict, edge_index_dict)}


class HeteroSAGE(nn.Module):
    """[HeteroConv(SAGEConv mean, sum)+ReLU]*L over several node types, Linear head on
    ``out_type`` (BASELINE.json config 4 shape)."""

    def __init__(self, in_dims, relations, out_type, conv_dims=(16, 16), head_dims=(16, 1), seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.out_type = out_type
        mods, dims = [], dict(in_dims)
        for h in conv_dims:
            mods += [HeteroConv({r: SAGEConv((dims[r[0]], dims[r[-1]]), h) for r in relations},
                                aggr="sum"), nn.ReLU()]
            dims = {k: h for k in dims}
        self.conv = nn.ModuleList(mods)
        self.fc = _head(head_dims, True)

    def forward(self, x_dict, edge_index_dict):
        x = x_dict
        for i, c in enumerate(self.conv):
            if i % 2 == 0:
                x = c(x, edge_index_dict)
            else:
                x = {k: c(v) for k, v in x.items()}
        x = x[self.out_type]
        for l in self.fc:
            x = l(x)
        return x


class HomoSAGE(nn.Module):
    def __init__(self, in_dim, conv_dims=(16, 16), head_dims=(16, 1), seed=0):
        super().__init__()
        torch.manual_seed(seed)
        mods, d = [], in_dim
        for h in conv_dims:
            mods += [SAGEConv(d, h), nn.ReLU()]
            d = h
        self.conv = nn.ModuleList(mods)
        self.fc = _head(head_dims, True)

    def forward(self, x, edge_index):
        for i, c in enumerate(self.conv):
            x = c(x, edge_index) if i % 2 == 0 else c(x)
        for l in self.fc:
            x = l(x)
        return x


def build_model(spec, state=None):
    spec = dict(spec)
    cls = globals()[spec.pop("cls")]
    if "relations" in spec:
        spec["relations"] = [tuple(r) for r in spec["relations"]]
    for k in ("conv_dims", "head_dims"):
        if k in spec:
            spec[k] = tuple(spec[k])
    m = cls(**spec)
    if state is not None:
        m.load_state_dict(state)
    return m.eval()
