"""SHAP kernel weights behind the reference's ``Kernel`` interface (``kernels.py:6-174``).

The binomial tables come from ``scipy.special.binom`` on the host (the very function the reference
calls, ``kernels.py:3,64,109``); the per-coalition weights, the per-batch shrinking-reference loop
of ``kernels.py:152-162`` and ``nan_to_num`` run on the device in float64.
"""
import numpy as np
import torch

from . import _lib
from .engine import require_cuda

_TABLES = {}


def _exact_table(m, dev):
    key = ("exact", m, str(dev))
    if key not in _TABLES:
        from scipy.special import binom

        _TABLES[key] = torch.from_numpy(binom(m, np.arange(m + 1)).astype(np.float64)).to(dev)
    return _TABLES[key]


def _approx_tables(dev):
    key = ("approx", str(dev))
    if key not in _TABLES:
        from scipy.special import binom

        tabs, ptr, ref = [], [0], 1000
        while ref > 0:  # kernels.py:152-162: ref = int(0.9 * ref) until it reaches 0
            tabs.append(binom(ref, np.arange(ref)).astype(np.float64))
            ptr.append(ptr[-1] + ref)
            ref = int(0.9 * ref)
        _TABLES[key] = (torch.from_numpy(np.concatenate(tabs)).to(dev),
                        torch.tensor(ptr, dtype=torch.int32, device=dev), len(tabs))
    return _TABLES[key]


def shap_weights(popcount, n_elements, batch_size):
    """popcount: int32 device tensor (active elements per coalition).  Returns float64 weights; the
    reference evaluates ``Kernel(mask).compute()`` once per batch (``wlm.py:421-422``), which matters
    for the approximate branch, hence ``batch_size``."""
    lib = _lib.load()
    dev = popcount.device
    s = int(popcount.numel())
    out = torch.empty(max(s, 1), dtype=torch.float64, device=dev)
    if n_elements - 1 <= 1000:
        tab = _exact_table(n_elements, dev)
        _lib.check(lib.xpgnn_shap_weights(popcount.data_ptr(), s, n_elements, batch_size, tab.data_ptr(), None, 1,
                                          out.data_ptr(), _lib.stream_ptr()))
    else:
        tab, ptr, nt = _approx_tables(dev)
        _lib.check(lib.xpgnn_shap_weights(popcount.data_ptr(), s, n_elements, batch_size, tab.data_ptr(),
                                          ptr.data_ptr(), nt, out.data_ptr(), _lib.stream_ptr()))
    return out[:s]


class Kernel:
    def __init__(self, mask):
        self.mask = mask

    def compute(self):
        dev = require_cuda()
        m = self.mask.to(dev)
        pop = m.sum(dim=1).to(torch.int32).contiguous()
        return shap_weights(pop, int(m.shape[1]), max(int(m.shape[0]), 1))
