#!/usr/bin/env python
"""Kernel-variant sweep on one workload of bench.py: the graph, model and coalition bits are built once, every variant
(a set of engine options, xpgnn_set_option) runs warm-up + timed
steps with the per-category CUDA-event profile on.  Prints one JSON line per variant.

  python tools/variants.py --workload c3 --coalitions 128 --variants "seg=0;seg=4;seg=8,seg_occ=6"
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--coalitions", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--variants", default="")
    ap.add_argument("--check", action="store_true", help="compare every variant's predictions with the first one's")
    args = ap.parse_args()
    from bikg_graph_explainability_public_b200 import _lib

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    wl = bench.Workload(args.workload)
    s = args.coalitions
    w = -(-s // 32)
    mask = bench.make_masks(s, wl.n, wl.c, wl.com_of, 1000, wl.com_of2).to(dev)
    act = torch.zeros((wl.n, w), dtype=torch.int32, device=dev)
    _lib.check(lib.xpgnn_pack_mask(mask.data_ptr(), s, wl.n, act.data_ptr(), w, None, _lib.stream_ptr()))
    cats = ["masked_degree", "spmm_invariant_l0", "spmm_tile_l1", "dense", "head", "compaction"]
    y0 = None
    for var in [v for v in args.variants.split(";")] or [""]:
        kv = {k: int(v) for k, v in (x.split("=") for x in var.split(",") if x)}
        old = _lib.set_options(**kv)
        try:
            arch, eng = wl.engine(dev, args.precision)
            for _ in range(args.warmup):
                y = eng(act, s)
            torch.cuda.synchronize()
            lib.xpgnn_profile(1)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(args.steps):
                y = eng(act, s)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / args.steps
            prof = (np.zeros(6), np.zeros(6, dtype=np.int64))
            lib.xpgnn_profile_read(prof[0].ctypes.data, prof[1].ctypes.data)
            lib.xpgnn_profile(0)
            out = {"variant": var, "ms_per_step": ms, "evals_per_s": s * len(wl.queries) / (ms / 1e3), "tile": eng.tile_coalitions,
                   "ms_per_launch": {c: (float(prof[0][i]) / max(int(prof[1][i]), 1)) for i, c in enumerate(cats)},
                   "launches": {c: int(prof[1][i]) for i, c in enumerate(cats)}}
            if args.check:
                yc = y.float().cpu().numpy()
                if y0 is None:
                    y0 = yc
                out["max_rel_diff_vs_first"] = float(np.max(np.abs(yc - y0) / np.maximum(np.abs(y0), 1e-6)))
            del eng
        except Exception as ex:  # keep sweeping
            out = {"variant": var, "error": repr(ex)}
        _lib.set_options(**old)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
