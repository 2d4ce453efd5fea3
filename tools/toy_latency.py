#!/usr/bin/env python
"""Latency of ``Explainer.run`` on the reference's own toy cases (BASELINE.json configs[0], [1]) -- launch-latency territory."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io as gio  # noqa: E402
from bikg_graph_explainability_public_b200 import Explainer  # noqa: E402

for name in ["c1_homo_gcn", "c1_homo_gcn_times3", "c2_hetero_gcn", "c4_hetero_sage"]:
    case = gio.load_case(name)
    meta = case["meta"]
    ts = []
    for rep in range(4):
        names, pathways, pnames = gio.fresh_inputs(case)
        arch = gio.build_arch(case)
        feat = {k: v.clone() for k, v in case["feat"].items()} if isinstance(case["feat"], dict) else case["feat"].clone()
        ei = {k: v.clone() for k, v in case["edge_index"].items()} if isinstance(case["edge_index"], dict) else case["edge_index"].clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ex = Explainer(feat, ei, arch, dict(meta["params"]), names, pathways, pnames, meta["element_type"], meta["problem"])
        ex.run(meta["element"], meta.get("times", 1))
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print("%-22s Explainer.run on B200: first %.3f s, then %s" % (name, ts[0], ["%.3f" % t for t in ts[1:]]), ex.last_stats)
