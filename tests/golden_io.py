"""Loads the fixtures written by ``oracle/make_golden.py`` (reference outputs) -- test infrastructure."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["c1_homo_gcn", "c1_homo_gcn_times3", "c2_hetero_gcn", "gcn2_random", "sage2_shapley", "c4_hetero_sage",
         "gcn2_5arg", "c2_dict_out", "gcn2_graph"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    case = dict(meta=meta, z=z)
    if meta["hetero"]:
        case["feat"] = {k: torch.from_numpy(z["feat::" + k]) for k in meta["node_types"]}
        case["edge_index"] = {tuple(r): torch.from_numpy(z["edge_index::" + "|".join(r)]) for r in meta["relations"]}
    else:
        case["feat"] = torch.from_numpy(z["feat"])
        case["edge_index"] = torch.from_numpy(z["edge_index"])
        if "node_types" in z.files:  # 5-argument protocol: caller-provided type vectors (explainer.py:83-84)
            case["node_types"] = torch.from_numpy(z["node_types"])
            case["edge_types"] = torch.from_numpy(z["edge_types"])
    case["state"] = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    return case


def fresh_inputs(case):
    """Deep copies of the mutable python inputs (the reference and the port mutate them)."""
    m = case["meta"]
    return (json.loads(json.dumps(m["names"])), json.loads(json.dumps(m["pathways"])),
            json.loads(json.dumps(m["pathway_names"])))


def build_arch(case):
    """The oracle's CPU stand-in model carrying the golden weights (lowered by duck typing)."""
    from oracle.fixture_models import build_model

    return build_model(case["meta"]["model"], case["state"])


def golden_mask(case, r=0):
    z = case["z"]
    rows, n = (int(v) for v in z["mask_shape_%d" % r])
    return np.unpackbits(z["mask_bits_%d" % r], axis=1)[:, :n].astype(bool)
