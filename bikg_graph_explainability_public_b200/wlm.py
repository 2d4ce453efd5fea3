"""Weighted linear surrogate behind the reference's ``wlm`` module (``wlm.py:17-520``).

``LinearRegression`` keeps the reference's module (its kaiming-uniform init consumes the global
CPU generator right after the masks -- that draw order is part of parity).  ``train_model`` keeps
the reference's signature; internally it evaluates every coalition with ``engine.MaskedForward``,
computes the SHAP kernel weights on the device and runs the fused fit kernels
(``xpgnn_wlm_fit``): one Adam(weight_decay=1e-2) step per coalition batch, in batch order.
"""
import numpy as np
import torch
from torch import nn

from . import _lib
from .kernels import shap_weights


class LinearRegression(nn.Module):
    """wlm.py:17-61."""

    def __init__(self, num_elements):
        assert isinstance(num_elements, int)
        super().__init__()
        self.layer = nn.Linear(int(num_elements), 1, bias=False)

    def forward(self, X):
        return self.layer(X)


def optimizer_assertions(params):
    """The argument checks of ``optimizer_scheduler`` (wlm.py:465-480)."""
    opt, lr, patience = params["optimizer"], params["lr"], params["lr_patience"]
    assert isinstance(opt, str), "Optimizer is not string"
    assert isinstance(lr, float) or isinstance(lr, int), "Learning rate given is not numeric"
    assert isinstance(patience, float) or isinstance(patience, int), "Patience for scheduler is not string"
    if opt.strip().lower() != "adam":
        raise NotImplementedError("Optimizer choice not available. Please choose between 'adam'")
    return abs(lr)


def fit_surrogate(coalitions, y, kern, w0, params, broadcast_y=True, want_losses=True):
    """Device fit.  y: (S,) float32 query predictions, kern: (S,) float64, w0: (N,) float32.
    Returns (weights (N,) float32 device tensor, losses list)."""
    lib = _lib.load()
    lr = optimizer_assertions(params)
    dev = coalitions.act.device
    s, n, b = coalitions.n_coalitions, coalitions.n_elements, coalitions.batch_size
    w = w0.detach().to(dev, torch.float32).clone().contiguous()
    y = y.to(dev, torch.float32).contiguous()
    kern = kern.to(dev, torch.float64).contiguous()
    steps = -(-s // b)
    losses = torch.zeros(steps, dtype=torch.float64, device=dev) if want_losses else None
    _lib.check(lib.xpgnn_wlm_fit(coalitions.act.data_ptr(), coalitions.words, n, s, b, y.data_ptr(), kern.data_ptr(),
                                 w.data_ptr(), float(lr), float(params["l1_lambda"]), 1e-2, int(bool(broadcast_y)),
                                 _lib.dptr(losses), _lib.stream_ptr()))
    return w, ([] if losses is None else losses.cpu().tolist())


def train_model(mask_loader, params, feat, edge_index, linear_model, arch, problem, element_index=None,
                node_type=None, edge_type=None, node_type_names=None, edge_type_names=None, padded_dims=None,
                engine=None):
    """Same positional contract as the reference (wlm.py:132-146).  ``mask_loader`` is the
    ``CoalitionSet`` returned by ``Mask.mask_generator``.  Returns (weights, losses, best_epoch)."""
    from .explainer import build_engine

    if engine is None:
        engine = build_engine(feat, edge_index, arch, element_index, node_type, edge_type, node_type_names,
                              edge_type_names, padded_dims)
    arch.eval()
    torch.empty((), dtype=torch.int64).random_()  # iter(DataLoader) base seed draw (wlm.py:210)
    y = engine(mask_loader.act, mask_loader.n_coalitions)[:, 0]
    kern = shap_weights(mask_loader.popcount, mask_loader.n_elements, mask_loader.batch_size)
    w0 = linear_model.layer.weight.detach().reshape(-1)
    w, losses = fit_surrogate(mask_loader, y, kern, w0, params, broadcast_y=engine.broadcast_y)
    with torch.no_grad():
        linear_model.layer.weight.copy_(w.reshape(1, -1).to(linear_model.layer.weight.device))
    best_epoch = int(np.argmin(losses)) if losses else 0
    return [w], losses, best_epoch  # "best" parameters are the final ones (wlm.py:94,258)


def regularizer(net, factor):
    """wlm.py:101-129 (host helper kept for API parity)."""
    a = torch.abs(torch.cat([p.view(-1) for p in net.parameters()]))
    return factor * (a.sum() / a.shape[0])


def weighted_mse_loss(input, target, weight):
    """wlm.py:491-520 (host helper kept for API parity)."""
    diff = (input.flatten() - target) ** 2
    return torch.mean(weight * diff) / weight.sum()
