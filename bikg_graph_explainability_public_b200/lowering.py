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


_CONV_NAMES = ("GCNConv", "SAGEConv", "HeteroConv")
_LEAF_NAMES = _CONV_NAMES + ("Linear", "Dropout") + tuple(_ACTS)


def _conv_entry(mod, model, consumed=None):
    """One ('conv', LoweredConv) entry for a GCNConv / SAGEConv / HeteroConv module."""
    if _name(mod) == "HeteroConv":
        if getattr(mod, "aggr", "sum") != "sum":
            raise NotImplementedError("only HeteroConv(aggr='sum') can be lowered")
        rels = []
        for key, conv in mod.convs.items():
            rels.append(_lower_conv(conv, tuple(key.split("__"))))
            if consumed is not None:
                for sub in conv.modules():
                    consumed.add(id(sub))
            model.n_message_passing += 1
        if consumed is not None:
            consumed.add(id(mod.convs))
        return ("conv", LoweredConv(rels, rels[0].w_nbr.shape[0], hetero=True))
    rel = _lower_conv(mod)
    if consumed is not None:
        for sub in mod.modules():
            consumed.add(id(sub))
    model.n_message_passing += 1
    return ("conv", LoweredConv([rel], rel.w_nbr.shape[0]))


def _sequence_from_modules(arch, model):
    """Registration order of ``arch.modules()``: right for models that register layers and activations in execution
    order (the reference's fixtures, ``tests/test_utils.py:10-182``); says nothing about what ``forward`` does."""
    consumed = set()
    seq = []
    for mod in arch.modules():
        if mod is arch or id(mod) in consumed or isinstance(mod, _CONTAINERS):
            continue
        n = _name(mod)
        if n in _CONV_NAMES:
            seq.append(_conv_entry(mod, model, consumed))
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
    return seq


def _sequence_from_trace(arch, model):
    """Execution order of ``arch.forward`` from a ``torch.fx`` symbolic trace (message passing layers, Linear and
    activation modules are leaves): sees functionally applied activations (``x.relu()``, ``F.relu``, ``torch.sigmoid``),
    dropout calls and layers that run in another order than they were registered.  Anything that is not a plain chain
    ``x -> conv(x, edge_index) -> act -> ... -> linear -> ...`` (skip connections, a layer used twice with different
    inputs is fine, sums of branches are not) raises ``NotImplementedError``."""
    import operator

    import torch.fx as fx
    import torch.nn.functional as F

    class _Tracer(fx.Tracer):
        def is_leaf_module(self, m, qualname):
            return _name(m) in _LEAF_NAMES or super().is_leaf_module(m, qualname)

    graph = _Tracer().trace(arch)
    mods = dict(arch.named_modules())
    relu_fns = {torch.relu, F.relu, torch.nn.functional.relu, torch.relu_}
    sig_fns = {torch.sigmoid, F.sigmoid, torch.nn.functional.sigmoid}
    drop_fns = {F.dropout, torch.dropout, torch.nn.functional.dropout}
    placeholders = [nd for nd in graph.nodes if nd.op == "placeholder"]
    if len(placeholders) < 2:
        raise NotImplementedError("forward takes fewer than two inputs")
    cur, edge = placeholders[0], placeholders[1]
    seq = []

    def chain(node):
        if not node.args or node.args[0] is not cur:
            raise NotImplementedError("forward is not a chain of layers (node %s)" % node.format_node())

    for node in graph.nodes:
        if node.op == "placeholder":
            continue
        if node.op == "call_module":
            m = mods[node.target]
            n = _name(m)
            chain(node)
            if n in _CONV_NAMES:
                second = node.args[1] if len(node.args) > 1 else node.kwargs.get("edge_index", node.kwargs.get("edge_index_dict"))
                if second is not edge:
                    raise NotImplementedError("message passing layer not applied to the input edge index")
                seq.append(_conv_entry(m, model))
            elif n == "Linear":
                seq.append(("linear", (m.weight.detach(), None if m.bias is None else m.bias.detach())))
            elif n in _ACTS:
                if _ACTS[n] is not None:
                    seq.append(("act", _ACTS[n]))
            elif n != "Dropout":
                raise NotImplementedError("module %s cannot be lowered" % n)
            cur = node
        elif node.op in ("call_function", "call_method"):
            t = node.target
            if t in relu_fns or t == "relu" or t == "relu_":
                chain(node)
                seq.append(("act", "relu"))
            elif t in sig_fns or t == "sigmoid":
                chain(node)
                seq.append(("act", "sigmoid"))
            elif t in drop_fns:
                chain(node)
                training = node.kwargs.get("training", node.args[2] if len(node.args) > 2 else True)
                if t is not torch.dropout and training not in (False,) and arch.training:
                    raise NotImplementedError("dropout active in forward")
            elif t in ("contiguous", "float", "clone") or t is operator.pos:
                chain(node)
            else:
                raise NotImplementedError("operation %s in forward cannot be lowered" % (t,))
            cur = node
        elif node.op == "output":
            if node.args[0] is not cur:
                raise NotImplementedError("forward does not return the last layer's output")
        else:
            raise NotImplementedError("fx node %s cannot be lowered" % node.op)
    return seq


def _assemble(seq, model):
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


def lower_candidates(arch: nn.Module):
    """Lowered plans of ``arch``, most trustworthy first: from the traced ``forward`` (when it can be traced), then from
    the module registration order.  ``explainer.verify_lowering`` picks the first that reproduces ``arch`` on a probe
    graph; ``lower`` alone (weight containers whose ``forward`` cannot run) returns the first."""
    out, errors = [], []
    was_training = arch.training
    arch.eval()
    try:
        for how, fn in (("trace", _sequence_from_trace), ("modules", _sequence_from_modules)):
            model = LoweredModel()
            try:
                m = _assemble(fn(arch, model), model)
                m.how = how
                out.append(m)
            except NotImplementedError as ex:
                errors.append(ex)
            except Exception as ex:  # fx cannot trace data-dependent control flow, dict comprehensions over proxies, ...
                if how != "trace":
                    raise
                errors.append(ex)
    finally:
        arch.train(was_training)
    if not out:
        raise errors[-1] if isinstance(errors[-1], NotImplementedError) else NotImplementedError(str(errors[-1]))
    return out


def lower(arch: nn.Module) -> LoweredModel:
    return lower_candidates(arch)[0]
