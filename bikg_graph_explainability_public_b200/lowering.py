"""Lowers a black-box ``arch`` (reference protocol: SURVEY.md 8b "arch protocol") to an engine plan.

The reference calls ``arch(feat, edge_index)`` on a materialised block-diagonal graph
(``model.py:104-112``).  The engine instead needs the *structure* of the model, so the module
tree is walked in registration order and matched against
``[conv (+ activation)]* [Linear (+ activation)]*`` where conv is GCNConv, SAGEConv(mean) or a
HeteroConv(sum) of those.  Anything else raises ``NotImplementedError`` -- never a CPU fallback.
Layers are duck-typed on class name + parameter layout, so real PyG modules, this package's
``nn`` containers and the oracle's CPU stand-ins all lower identically.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
from torch import nn

_CONTAINERS = (nn.ModuleList, nn.Sequential, nn.ModuleDict)
_ACTS = {"ReLU": "relu", "Sigmoid": "sigmoid", "Identity": None}


@dataclass
class LoweredRelation:
    kind: str                         # "gcn" | "sage"
    key: Optional[Tuple[str, str, str]]  # (src, rel, dst) or None for homogeneous
    w_nbr: torch.Tensor               # [out, in_src]
    b_nbr: Optional[torch.Tensor]
    w_root: Optional[torch.Tensor]    # [out, in_dst] (SAGE root weight)


@dataclass
class LoweredConv:
    relations: List[LoweredRelation]
    out_dim: int
    act: Optional[str] = None
    hetero: bool = False


@dataclass
class LoweredModel:
    convs: List[LoweredConv] = field(default_factory=list)
    head: List[Tuple[torch.Tensor, Optional[torch.Tensor], Optional[str]]] = field(default_factory=list)
    n_message_passing: int = 0        # what PyG ``get_num_hops`` would count (model.py:52)

    @property
    def hetero(self):
        return any(c.hetero for c in self.convs)

    @property
    def out_dim(self):
        return self.head[-1][0].shape[0] if self.head else self.convs[-1].out_dim


def _name(m):
    return type(m).__name__


def _lower_conv(m, key=None):
    n = _name(m)
    if n == "GCNConv":
        for attr, want in (("improved", False), ("add_self_loops", True), ("normalize", True)):
            if getattr(m, attr, want) != want:
                raise NotImplementedError("GCNConv with %s=%r cannot be lowered" % (attr, getattr(m, attr)))
        return LoweredRelation("gcn", key, m.lin.weight.detach(), None if m.bias is None else m.bias.detach(), None)
    if n == "SAGEConv":
        if getattr(m, "aggr", "mean") not in ("mean", None) and not isinstance(getattr(m, "aggr", "mean"), str):
            raise NotImplementedError("SAGEConv aggregation %r cannot be lowered" % (m.aggr,))
        if getattr(m, "aggr", "mean") != "mean" or getattr(m, "normalize", False):
            raise NotImplementedError("only SAGEConv(aggr='mean', normalize=False) can be lowered")
        w_root = m.lin_r.weight.detach() if getattr(m, "root_weight", True) and hasattr(m, "lin_r") else None
        b = None if getattr(m.lin_l, "bias", None) is None else m.lin_l.bias.detach()
        return LoweredRelation("sage", key, m.lin_l.weight.detach(), b, w_root)
    raise NotImplementedError("message passing layer %s cannot be lowered to the CUDA engine" % n)


def lower(arch: nn.Module) -> LoweredModel:
    model = LoweredModel()
    consumed = set()
    seq = []
    for mod in arch.modules():
        if mod is arch or id(mod) in consumed or isinstance(mod, _CONTAINERS):
            continue
        n = _name(mod)
        if n == "HeteroConv":
            if getattr(mod, "aggr", "sum") != "sum":
                raise NotImplementedError("only HeteroConv(aggr='sum') can be lowered")
            rels = []
            for key, conv in mod.convs.items():
                rels.append(_lower_conv(conv, tuple(key.split("__"))))
                for sub in conv.modules():
                    consumed.add(id(sub))
                model.n_message_passing += 1
            consumed.add(id(mod.convs))
            seq.append(("conv", LoweredConv(rels, rels[0].w_nbr.shape[0], hetero=True)))
        elif n in ("GCNConv", "SAGEConv"):
            rel = _lower_conv(mod)
            for sub in mod.modules():
                consumed.add(id(sub))
            model.n_message_passing += 1
            seq.append(("conv", LoweredConv([rel], rel.w_nbr.shape[0])))
        elif n == "Linear":
            seq.append(("linear", (mod.weight.detach(), None if mod.bias is None else mod.bias.detach())))
        elif n in _ACTS:
            seq.append(("act", _ACTS[n]))
        elif n == "Dropout":
            continue  # arch.eval() (wlm.py:204): identity
        else:
            raise NotImplementedError(
                "module %s cannot be lowered to the CUDA engine (supported: GCNConv, SAGEConv(mean), "
                "HeteroConv(sum), Linear, ReLU, Sigmoid)" % n)
    i, stage = 0, "conv"
    while i < len(seq):
        kind, payload = seq[i]
        act = None
        if i + 1 < len(seq) and seq[i + 1][0] == "act":
            act = seq[i + 1][1]
        if kind == "conv":
            if stage != "conv":
                raise NotImplementedError("a message passing layer after the MLP head cannot be lowered")
            payload.act = act
            model.convs.append(payload)
        elif kind == "linear":
            stage = "head"
            model.head.append((payload[0], payload[1], act))
        elif kind == "act":
            raise NotImplementedError("two consecutive activations / leading activation cannot be lowered")
        i += 2 if (i + 1 < len(seq) and seq[i + 1][0] == "act") else 1
    if not model.convs:
        raise NotImplementedError("arch holds no message passing layer")
    return model
