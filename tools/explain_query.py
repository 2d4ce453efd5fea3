#!/usr/bin/env python
"""Seconds per explained query through the drop-in API (BASELINE.json metric):
``Explainer(feat, edge_index, arch, params, names, pathways, pathway_names).run(query, 1)`` end to end
(host prep, k-hop cut, masks, perturbed forward, SHAP weights, surrogate fit, DataFrames) on a synthetic graph.

  python tools/explain_query.py --nodes 1000000 --edges 20000000 --graph rmat --communities 500
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rmat_edges(n, e, g, a=0.57, b=0.19, c=0.19):
    """R-MAT(0.57, 0.19, 0.19, 0.05) (hub rows give large receptive fields)."""
    bits = int(np.ceil(np.log2(n)))
    src = torch.zeros(e, dtype=torch.int64)
    dst = torch.zeros(e, dtype=torch.int64)
    for _ in range(bits):
        r = torch.rand(e, generator=g)
        sb = (r >= a + b).to(torch.int64)                      # quadrants c, d -> source bit 1
        db = ((r >= a) & (r < a + b) | (r >= a + b + c)).to(torch.int64)  # quadrants b, d -> target bit 1
        src = src * 2 + sb
        dst = dst * 2 + db
    return torch.stack([src % n, dst % n])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=100_000)
    ap.add_argument("--edges", type=int, default=2_000_000)
    ap.add_argument("--features", type=int, default=128)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--communities", type=int, default=50)
    ap.add_argument("--graph", default="uniform", choices=["uniform", "rmat"])
    ap.add_argument("--interpret-samples", type=int, default=64)
    ap.add_argument("--epochs", type=int, default=64)
    ap.add_argument("--names", default="str", choices=["str", "int"], help="community members as names or indices")
    ap.add_argument("--queries", type=int, default=2)
    ap.add_argument("--prune", type=int, default=1)
    ap.add_argument("--profile", action="store_true", help="per-category kernel ms of every query (CUDA events, xpgnn_profile)")
    ap.add_argument("--hostprof", action="store_true", help="cProfile of the last query (run with CUDA_LAUNCH_BLOCKING=1 so that "
                    "device time is charged to the launching call)")
    ap.add_argument("--device-inputs", action="store_true", help="feat / edge_index already on the GPU (device convention, SURVEY.md 8b)")
    args = ap.parse_args()

    from torch import nn

    from bikg_graph_explainability_public_b200 import nn as xnn
    from pathway_explanations.explainer import Explainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:  # one process per GPU: the coalitions of every query are sharded, one NCCL all-gather of the predictions
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))

    g = torch.Generator().manual_seed(1234)
    n, e = args.nodes, args.edges
    ei = torch.randint(0, n, (2, e), generator=g) if args.graph == "uniform" else rmat_edges(n, e, g)
    x = torch.randn(n, args.features, generator=g)
    com_of = torch.randperm(n, generator=g) % args.communities
    names = [str(i) for i in range(n)]
    order = torch.argsort(com_of, stable=True)
    bounds = torch.searchsorted(com_of[order], torch.arange(args.communities + 1))
    members = [order[bounds[c]:bounds[c + 1]].tolist() for c in range(args.communities)]
    pathways = members if args.names == "int" else [[names[i] for i in m] for m in members]
    pathway_names = ["community_%d" % c for c in range(args.communities)]
    torch.manual_seed(7)

    class GCN2(nn.Module):
        def __init__(self, f, h):
            super().__init__()
            self.conv = nn.ModuleList([xnn.GCNConv(f, h), nn.ReLU(), xnn.GCNConv(h, h), nn.ReLU()])
            self.fc = nn.ModuleList([xnn.Linear(h, 1)])

    arch = GCN2(args.features, args.hidden).eval().cuda()
    params = {"interpret_samples": args.interpret_samples, "epochs": args.epochs, "optimizer": "adam", "lr": 0.01,
              "l1_lambda": 1e-4, "lr_patience": 10, "seed": 1}  # the reference's config/configs.json
    indeg = torch.bincount(ei[1], minlength=n)
    queries = [int(torch.argmax(indeg))] + torch.randint(0, n, (args.queries,), generator=g).tolist()
    Explainer.engine_options = dict(Explainer.engine_options, prune=bool(args.prune), precision="fp32")
    if args.device_inputs:
        x, ei = x.cuda(), ei.cuda()
    out = []
    for q in queries[:args.queries]:
        pw = [list(p) for p in pathways]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.profile:
            from bikg_graph_explainability_public_b200 import _lib
            _lib.load().xpgnn_profile(1)
        ex = Explainer(x, ei, arch, dict(params), list(names), pw, list(pathway_names))
        cfg, pdf = ex.run(names[q], 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if args.profile:
            ms, cnt = np.zeros(6), np.zeros(6, dtype=np.int64)
            _lib.load().xpgnn_profile_read(ms.ctypes.data, cnt.ctypes.data)
            _lib.load().xpgnn_profile(0)
            cats = ["masked_degree", "spmm_l0", "spmm_l1", "dense", "head", "compaction"]
            print(json.dumps({"query": q, "kernel_ms": {c: [round(float(m), 2), int(k)] for c, m, k in zip(cats, ms, cnt)}}), flush=True)
        out.append({"query": q, "in_degree": int(indeg[q]), "seconds": dt, **ex.last_stats,
                    "top_community": None if pdf is None or len(pdf) == 0 else str(pdf.index[0]),
                    "score_checksum": None if pdf is None else float(pdf["score"].abs().sum()), "rank": rank, "world": world})
        print(json.dumps(out[-1]), flush=True)
    if args.hostprof and rank == 0:
        import cProfile
        import pstats

        pw = [list(p) for p in pathways]
        pr = cProfile.Profile()
        pr.enable()
        ex = Explainer(x, ei, arch, dict(params), list(names), pw, list(pathway_names))
        ex.run(names[queries[0]], 1)
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps({"graph": args.graph, "nodes": n, "edges": e, "communities": args.communities, "names": args.names,
                          "world": world, "s_per_explained_query_median": float(np.median([o["seconds"] for o in out]))}))


if __name__ == "__main__":
    main()
