"""Layer containers recognised by the plan lowering.

They carry parameters under the same state-dict keys as the PyG 2.0.4 layers the reference's
fixtures use (``tests/test_utils.py:7,46,62,133-139``): ``lin.weight`` / ``bias`` (GCNConv),
``lin_l.{weight,bias}`` / ``lin_r.weight`` (SAGEConv), ``convs.<src>__<rel>__<dst>.*`` (HeteroConv),
so ``test_data/*.pth.tar`` checkpoints load without PyG installed.  They hold weights only: the
arithmetic runs in the CUDA engine (``engine.MaskedForward``); calling ``forward`` on a layer
directly is not supported (there is no CPU path).  Real PyG layers of the same class names are
lowered the same way (duck-typed on class name and parameter layout).
"""
import math

import torch
from torch import nn


class MessagePassing(nn.Module):
    """Marker base: ``Model.get_hops`` counts instances (reference ``model.py:52``)."""

    def forward(self, *a, **k):
        raise NotImplementedError(
            "layers are weight containers; run the model through Explainer / engine.MaskedForward "
            "(CUDA kernels, no CPU fallback)")


class Linear(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None, bias_initializer=None):
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
            bound = 1.0 / math.sqrt(in_channels) if in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        raise NotImplementedError("weight container; see module docstring")


class GCNConv(MessagePassing):
    def __init__(self, in_channels, out_channels, bias=True, **kwargs):
        super().__init__()
        for k, v in (("improved", False), ("add_self_loops", True), ("normalize", True)):
            if kwargs.get(k, v) != v:
                raise NotImplementedError("GCNConv(%s=%r) is not supported by the engine" % (k, kwargs[k]))
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None


class SAGEConv(MessagePassing):
    def __init__(self, in_channels, out_channels, aggr="mean", normalize=False, root_weight=True, bias=True, **kw):
        super().__init__()
        if aggr != "mean" or normalize:
            raise NotImplementedError("only SAGEConv(aggr='mean', normalize=False) is supported")
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels, self.root_weight = tuple(in_channels), out_channels, root_weight
        self.lin_l = Linear(in_channels[0], out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels[1], out_channels, bias=False)


class HeteroConv(nn.Module):
    def __init__(self, convs, aggr="sum"):
        super().__init__()
        if aggr != "sum":
            raise NotImplementedError("only HeteroConv(aggr='sum') is supported")
        self.convs = nn.ModuleDict({"__".join(k): m for k, m in convs.items()})
        self.aggr = aggr

    def forward(self, *a, **k):
        raise NotImplementedError("weight container; see module docstring")
