"""tcgen05 dense transform (TF32x3 and BF16 modes) vs the exact fp32 SIMT path, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dense(lib, a, w, b, act, precision, accumulate=None):
    from bikg_graph_explainability_public_b200 import _lib

    m, k = a.shape
    n = w.shape[0]
    out = torch.zeros((m, n), device="cuda") if accumulate is None else accumulate.clone()
    _lib.check(lib.xpgnn_dense_rows(a.data_ptr(), m, k, k, w.data_ptr(), _lib.dptr(b), n, act, out.data_ptr(), n,
                                    int(accumulate is not None), precision, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("m,k,n", [(1000, 128, 128), (4096, 64, 16), (300, 32, 40), (129, 256, 64), (70000, 128, 128)])
def test_tensor_core_dense_matches_fp32(lib, m, k, n):
    g = torch.Generator(device="cuda").manual_seed(m + k + n)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g)
    ref = (a.double() @ w.double().T + b.double())
    simt = _dense(lib, a, w, b, 0, 0)
    assert torch.allclose(simt.double(), ref, rtol=1e-5, atol=1e-5)
    scale = float(ref.abs().max())
    x3 = _dense(lib, a, w, b, 0, 2)
    err3 = float((x3.double() - ref).abs().max()) / scale
    assert err3 < 2e-6, "TF32x3 error %g" % err3
    bf = _dense(lib, a, w, b, 0, 1)
    errb = float((bf.double() - ref).abs().max()) / scale
    assert errb < 2e-2, "bf16 error %g" % errb
    # fused epilogue: accumulate + ReLU, no bias
    prev = torch.randn(m, n, device="cuda", generator=g)
    ref2 = torch.relu(prev.double() + a.double() @ w.double().T)
    y = _dense(lib, a, w, None, 1, 2, accumulate=prev)
    assert float((y.double() - ref2).abs().max()) / scale < 2e-6
