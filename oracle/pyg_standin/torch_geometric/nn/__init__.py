"""ORACLE / TEST INFRASTRUCTURE ONLY.

CPU restatement (plain torch ops, fp32) of the PyG 2.0.4 layers that the
reference's fixtures and notebooks instantiate
(``tests/test_utils.py:7,46,62,133-139``; ``examples/toy_example-caseA.ipynb``
cell 9).  State-dict keys follow PyG (``lin.weight``, ``bias``, ``lin_l.*``,
``lin_r.weight``, ``convs.<src>__<rel>__<dst>.*``) so the reference's
``test_data/*.pth.tar`` checkpoints load.
"""
import math

import torch
from torch import nn


class MessagePassing(nn.Module):
    """Marker base class: ``get_num_hops`` counts instances of it."""


class Linear(nn.Module):
    """``torch_geometric.nn.Linear``: y = x W^T + b (kaiming-uniform default)."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None,
                 bias_initializer=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        if weight_initializer == "glorot":
            a = math.sqrt(6.0 / (in_channels + out_channels))
            nn.init.uniform_(self.weight, -a, a)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            if bias_initializer == "zeros":
                nn.init.zeros_(self.bias)
            else:
                bound = 1.0 / math.sqrt(in_channels) if in_channels > 0 else 0.0
                nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return nn.functional.linear(x, self.weight, self.bias)


def _scatter_add(src, index, dim_size):
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def gcn_norm(edge_index, num_nodes, dtype):
    """``gcn_norm`` with ``add_remaining_self_loops`` (fill 1), unit edge weights.

    Existing self-loops are removed and exactly one unit self-loop per node is
    appended; degree = in-degree on targets (incl. the self-loop).
    """
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loop = torch.arange(num_nodes, dtype=row.dtype, device=row.device)
    row = torch.cat([row[keep], loop])
    col = torch.cat([col[keep], loop])
    w = torch.ones(row.numel(), dtype=dtype, device=row.device)
    deg = _scatter_add(w, col, num_nodes)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return row, col, dis[row] * w * dis[col]


class GCNConv(MessagePassing):
    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True, **kwargs):
        super().__init__()
        assert not improved and add_self_loops and normalize
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x, edge_index, edge_weight=None):
        assert edge_weight is None
        n = x.size(0)
        row, col, norm = gcn_norm(edge_index, n, x.dtype)
        z = self.lin(x)
        out = _scatter_add(z[row] * norm.view(-1, 1), col, n)
        if self.bias is not None:
            out = out + self.bias
        return out


class SAGEConv(MessagePassing):
    """SAGEConv(aggr='mean', root_weight=True): lin_l(mean_{u->v} x_u) + lin_r(x_v)."""

    def __init__(self, in_channels, out_channels, normalize=False, root_weight=True, bias=True,
                 **kwargs):
        super().__init__()
        assert not normalize
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels, self.root_weight = in_channels, out_channels, root_weight
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index):
        if isinstance(x, torch.Tensor):
            x = (x, x)
        x_src, x_dst = x
        n_dst = x_dst.size(0)
        row, col = edge_index[0], edge_index[1]
        s = _scatter_add(x_src[row], col, n_dst)
        cnt = _scatter_add(torch.ones(row.numel(), dtype=x_src.dtype), col, n_dst)
        out = self.lin_l(s / cnt.clamp(min=1).view(-1, 1))
        if self.root_weight:
            out = out + self.lin_r(x_dst)
        return out


class GATConv(MessagePassing):  # fixture-only layer (tests/test_utils.py:135); not on the hot path
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("GATConv is outside the restated hot path")


class HeteroConv(nn.Module):
    def __init__(self, convs, aggr="sum"):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(k): m for k, m in convs.items()})
        self.aggr = aggr

    def forward(self, x_dict, edge_index_dict):
        out_dict = {}
        for edge_type, edge_index in edge_index_dict.items():
            src, dst = edge_type[0], edge_type[-1]
            key = "__".join(edge_type)
            if key not in self.convs:
                continue
            conv = self.convs[key]
            if src == dst:
                out = conv(x_dict[src], edge_index)
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index)
            out_dict.setdefault(dst, []).append(out)
        res = {}
        for key, xs in out_dict.items():
            if self.aggr is None:
                res[key] = torch.stack(xs, dim=1)
            else:
                o = getattr(torch, self.aggr)(torch.stack(xs, dim=0), dim=0)
                res[key] = o[0] if isinstance(o, tuple) else o
        return res
