"""Micro-benchmark of xpgnn_dense_rows (SIMT / BF16 tcgen05 / TF32x3 tcgen05) on an engine-sized problem."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bikg_graph_explainability_public_b200 import _lib

lib = _lib.load()
m, k, n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 128, 128
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
a = torch.randn(m, k, device="cuda")
w = torch.randn(n, k, device="cuda") / k ** 0.5
b = torch.randn(n, device="cuda")
out = torch.empty(m, n, device="cuda")
for prec in modes:
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.xpgnn_dense_rows(a.data_ptr(), m, k, k, w.data_ptr(), b.data_ptr(), n, 1, out.data_ptr(), n, 0, prec,
                                        _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # accuracy on the first 4096 rows against fp64 (act = ReLU in this call)
    ref = torch.relu(a[:4096].double() @ w.double().t() + b.double())
    err = ((out[:4096].double() - ref).abs().max() / ref.abs().max()).item()
    print("precision %d: %.3f ms  %.1f TFLOP/s (x%d MMAs)  %.0f GB/s  max err / max |ref| %.2e" % (
        prec, ms, 2.0 * m * k * n / ms / 1e9, 3 if prec == 2 else 1, 2.0 * m * k * 4 / ms / 1e6, err))
