"""Host side of the coalition-batched masked message-passing engine.

Owns the device-resident computational graph (per-relation CSR by destination), the lowered
model weights and the ``xpgnn_plan_t`` handed to ``xpgnn_forward`` (``include/xpgnn_b200.h``).
Replaces the reference's ``Data.perturbator`` + ``Model.infer`` /
``Model.predict_hetero_output`` + ``extract_node_edge_output`` chain
(``data.py:591-648``, ``model.py:62-328``, ``wlm.py:350-436``): one call evaluates the query
prediction under every coalition of a packed bit matrix.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import _lib
from .lowering import LoweredModel

_ACT = {None: _lib.ACT_NONE, "relu": _lib.ACT_RELU, "sigmoid": _lib.ACT_SIGMOID}
_KIND = {"gcn": _lib.CONV_GCN, "sage": _lib.CONV_SAGE_MEAN}


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.XpgnnError("a CUDA device is required: the perturbation path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


@dataclass
class GraphSpec:
    """Flattened computational graph (device tensors)."""
    x: torch.Tensor                      # (N, F) float32
    edge_index: torch.Tensor             # (2, E) int64
    type_ptr: List[int]                  # node id range of each node type: [0, .., N]
    node_type_names: Optional[List[str]] = None
    edge_type: Optional[torch.Tensor] = None          # (E,) integer relation id per edge
    edge_type_names: Optional[List[Tuple[str, str, str]]] = None

    @property
    def n_nodes(self):
        return int(self.x.shape[0])


def build_csr(src, dst, n_nodes, drop_self_loops):
    """Device CSR by destination (``xpgnn_build_csr``).  Returns rowptr int32 [N+1], col int32 [E']."""
    lib = _lib.load()
    e = int(src.numel())
    src = src.contiguous().to(torch.int64)
    dst = dst.contiguous().to(torch.int64)
    rowptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=src.device)
    col = torch.empty(max(e, 1), dtype=torch.int32, device=src.device)
    kept = torch.zeros(1, dtype=torch.int64, device=src.device)
    _lib.check(lib.xpgnn_build_csr(_lib.dptr(src), _lib.dptr(dst), e, n_nodes, int(drop_self_loops),
                                   _lib.dptr(rowptr), _lib.dptr(col), _lib.dptr(kept), _lib.stream_ptr()))
    return rowptr, col


class MaskedForward:
    """y[s, q] = arch(graph perturbed by coalition s)[query q]  for a packed coalition matrix."""

    def __init__(self, graph: GraphSpec, model: LoweredModel, queries, prune=False, hop=None,
                 zero_edge_rule=False, precision="fp32", out_col=0, tile_coalitions=None,
                 workspace_fraction=0.8):
        self.lib = _lib.load()
        dev = require_cuda()
        self.graph, self.model = graph, model
        n = graph.n_nodes
        self._keep = []  # tensors / ctypes objects referenced by raw pointers
        x = graph.x.to(dev, torch.float32).contiguous()
        self._keep.append(x)
        ei = graph.edge_index.to(dev)
        tnames = graph.node_type_names
        self.edges_per_layer = []

        def rng(tname):
            if tnames is None or tname is None:
                return 0, n
            t = tnames.index(tname)
            return graph.type_ptr[t], graph.type_ptr[t + 1]

        csr_cache = {}

        def csr_for(key, kind):
            ck = (key, kind == "gcn")
            if ck not in csr_cache:
                if key is None:
                    src, dst = ei[0], ei[1]
                else:
                    if graph.edge_type_names is None or key not in graph.edge_type_names:
                        raise NotImplementedError("model relation %r has no edge type in the graph" % (key,))
                    sel = graph.edge_type.to(dev) == graph.edge_type_names.index(key)
                    src, dst = ei[0][sel], ei[1][sel]
                rowptr, col = build_csr(src, dst, n, drop_self_loops=(kind == "gcn"))
                csr_cache[ck] = (rowptr, col, int(rowptr[-1].item()))
            return csr_cache[ck]

        f_in = int(x.shape[1])
        layers = (_lib.Layer * len(model.convs))()
        h_in = f_in
        for li, conv in enumerate(model.convs):
            rels = (_lib.Relation * len(conv.relations))()
            n_edges = 0
            for ri, r in enumerate(conv.relations):
                src_t, dst_t = (r.key[0], r.key[-1]) if r.key is not None else (None, None)
                if r.kind == "gcn" and src_t != dst_t:
                    raise NotImplementedError("GCNConv over a bipartite relation is not defined (PyG neither)")
                rowptr, col, e_r = csr_for(r.key, r.kind)
                n_edges += e_r
                w_nbr = self._weight(r.w_nbr, h_in, dev)
                w_root = None if r.w_root is None else self._weight(r.w_root, h_in, dev)
                b = None if r.b_nbr is None else r.b_nbr.to(dev, torch.float32).contiguous()
                self._keep += [rowptr, col, w_nbr, w_root, b]
                R = rels[ri]
                R.conv_kind = _KIND[r.kind]
                R.src_lo, R.src_hi = rng(src_t)
                R.dst_lo, R.dst_hi = rng(dst_t)
                R.n_edges = e_r
                R.rowptr, R.col = rowptr.data_ptr(), col.data_ptr()
                R.w_nbr = w_nbr.data_ptr()
                R.b_nbr = _lib.dptr(b)
                R.w_root = _lib.dptr(w_root)
            self._keep.append(rels)
            L = layers[li]
            L.n_rel, L.rel_host, L.h_in, L.h_out, L.act = len(conv.relations), rels, h_in, conv.out_dim, _ACT[conv.act]
            self.edges_per_layer.append(n_edges)
            h_in = conv.out_dim
        head = (_lib.Dense * max(len(model.head), 1))()
        for i, (w, b, act) in enumerate(model.head):
            w = w.to(dev, torch.float32).contiguous()
            b = None if b is None else b.to(dev, torch.float32).contiguous()
            self._keep += [w, b]
            head[i].in_, head[i].out, head[i].act = int(w.shape[1]), int(w.shape[0]), _ACT[act]
            head[i].w, head[i].b = w.data_ptr(), _lib.dptr(b)
        q = torch.as_tensor(queries, dtype=torch.int32).reshape(-1).to(dev).contiguous()
        hop_t = None if hop is None else hop.to(dev, torch.int8).contiguous()
        if prune and hop_t is None:
            raise ValueError("prune=True needs the BFS levels from the k-hop kernel")
        self._keep += [layers, head, q, hop_t]
        p = _lib.Plan()
        p.n_nodes, p.f_in, p.x = n, f_in, x.data_ptr()
        p.n_layers, p.layers_host = len(model.convs), layers
        p.n_head, p.head_host = len(model.head), head
        p.n_query, p.query, p.out_col = int(q.numel()), q.data_ptr(), int(out_col)
        p.prune, p.hop = int(bool(prune)), _lib.dptr(hop_t)
        p.zero_edge_rule = int(bool(zero_edge_rule))
        # "fp32": fp32 storage, dense transforms as 3xTF32 tcgen05 products (~1e-6 relative); "fp32_exact": the same plan with
        # exact fp32 FMA transforms (the dense_simt engine option, set around every call of this engine)
        p.precision = {"fp32": 0, "fp32_exact": 0, "bf16": 1, "bf16_act": 2}[precision]
        self.exact_transforms = precision == "fp32_exact"
        self.plan, self.n_query, self.device = p, int(q.numel()), dev
        self.stats = torch.zeros(4, dtype=torch.int64, device=dev)
        # workspace: the largest coalition tile that fits
        free = torch.cuda.mem_get_info(dev)[0] + torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
        budget = int(free * workspace_fraction)
        tile = 32 if tile_coalitions is None else int(tile_coalitions)
        while True:
            need = int(self.lib.xpgnn_forward_workspace_bytes(C.byref(p), tile))
            if need < 0:
                raise _lib.XpgnnError("invalid plan")
            if need <= budget or tile == 1:
                break
            tile //= 2
        if need > budget:
            raise _lib.XpgnnError("graph does not fit: %d bytes of workspace needed, %d available" % (need, budget))
        self.tile_coalitions = tile
        self.workspace = torch.empty(need, dtype=torch.uint8, device=dev)

    @staticmethod
    def _kept_edges(rowptr):
        return int(rowptr[-1].item())

    @staticmethod
    def _weight(w, k, dev):
        """[out, in] fp32 on device, zero-padded to ``k`` input columns (padded hetero features)."""
        w = w.to(dev, torch.float32)
        if w.shape[1] < k:
            w = torch.nn.functional.pad(w, (0, k - w.shape[1]))
        elif w.shape[1] > k:
            raise ValueError("weight expects %d input features, graph provides %d" % (w.shape[1], k))
        return w.contiguous()

    @property
    def edge_visits_per_coalition(self):
        """Edges of the computational graph examined per coalition (sum over conv layers)."""
        return sum(self.edges_per_layer)

    def __call__(self, act, n_coalitions, s0=0, n_s=None):
        """act: packed (N, W) int32/uint32 coalition bits.  Returns y (n_s, n_query) float32."""
        n_s = n_coalitions - s0 if n_s is None else n_s
        assert act.is_cuda and act.dim() == 2 and act.shape[0] == self.graph.n_nodes and act.is_contiguous()
        w = int(act.shape[1])
        y = torch.empty((n_s, self.n_query), dtype=torch.float32, device=self.device)
        old = _lib.set_options(dense_simt=1) if self.exact_transforms else None
        try:
            _lib.check(self.lib.xpgnn_forward(C.byref(self.plan), act.data_ptr(), w, int(s0), int(n_s), y.data_ptr(),
                                              self.workspace.data_ptr(), int(self.workspace.numel()),
                                              self.stats.data_ptr(), _lib.stream_ptr()))
        finally:
            if old is not None:
                _lib.set_options(**old)
        return y
