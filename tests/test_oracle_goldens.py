"""The oracle port vs the fixtures produced by the UNMODIFIED reference (``oracle/make_golden.py``).

Runs anywhere (no reference tree, no GPU): this is what pins the oracle on the GPU box."""
import copy
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import golden_io as gio
from oracle import xpgnn_oracle as orc
from oracle.mt19937 import MT19937


@pytest.mark.parametrize("name", gio.CASES)
def test_oracle_explain_matches_reference_golden(name):
    case = gio.load_case(name)
    z, meta = case["z"], case["meta"]
    names, pathways, pnames = gio.fresh_inputs(case)
    arch = gio.build_arch(case)
    mt = MT19937.from_torch_state(z["rng_state"]) if "rng_state" in z.files else None
    o = orc.explain(copy.deepcopy(case["feat"]), copy.deepcopy(case["edge_index"]), arch, dict(meta["params"]), names,
                    pathways, pnames, meta["element_type"], meta["problem"], element=meta["element"],
                    times=meta["times"], mt=mt, node_types=copy.deepcopy(case.get("node_types")),
                    edge_types=copy.deepcopy(case.get("edge_types")))
    assert np.array_equal(o["subset"], z["subset"]) and np.array_equal(o["sub_edge_index"], z["sub_edge_index"])
    assert o["sub_ind"] == int(z["sub_ind"])
    for r, run in enumerate(o["runs"]):
        assert np.array_equal(run["mask"], gio.golden_mask(case, r))
        if run["pathway_rows"] is not None:
            assert np.array_equal(run["pathway_rows"], z["pathway_rows_%d" % r])
        assert run["batch_size"] == int(z["batch_size_%d" % r])
        assert np.array_equal(run["w0"], z["w0_%d" % r])
        assert np.array_equal(np.concatenate([b[1] for b in run["batches"]]), z["kernel_%d" % r])
        y = torch.cat([b[2] for b in run["batches"]]).numpy()
        np.testing.assert_allclose(y, z["y_%d" % r], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(run["weights"], z["weights_%d" % r], rtol=1e-5, atol=1e-7)
        closed = orc.train_wlm_closed_form(run["batches"], run["w0"], meta["params"])
        np.testing.assert_allclose(closed, z["weights_%d" % r], rtol=1e-4, atol=2e-6)
    assert [str(x) for x in o["config_val_df"].index] == [str(x) for x in z["cfg_names"]]
    if "pw_names" in z.files:
        assert [str(x) for x in o["pathway_df"].index] == [str(x) for x in z["pw_names"]]
        np.testing.assert_allclose(o["pathway_df"]["score"].values, z["pw_score"], rtol=1e-5, atol=1e-7)


def test_oracle_mask_stream_goldens():
    with open(os.path.join(gio.GOLDEN, "mask_stream.json")) as f:
        specs = json.load(f)
    for sp in specs:
        mt = MT19937(sp["seed"] + 2)
        params = dict(interpret_samples=sp["interpret_samples"], epochs=sp["epochs"])
        m, prow, bsz = orc.mask_generator(sp["n"], copy.deepcopy(sp["communities"]), params, mt)
        assert hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest() == sp["sha256_mask"], sp["name"]
        assert hashlib.sha256(prow.astype(np.int32).tobytes()).hexdigest() == sp["sha256_rows"], sp["name"]
        assert mt.consumed == sp["consumed"] and bsz == sp["batch_size"]


def test_oracle_shap_kernel_goldens():
    z = np.load(os.path.join(gio.GOLDEN, "shap_kernel.npz"))
    for key in [k[len("kernel_"):] for k in z.files if k.startswith("kernel_")]:
        rows, n = (int(v) for v in z["shape_" + key])
        m = np.unpackbits(z["mask_" + key], axis=1)[:, :n].astype(bool)
        assert np.array_equal(orc.shap_kernel(m), z["kernel_" + key]), key


def test_reference_known_answers():
    """Known-answer vectors held by the reference's own unit tests."""
    # tests/test_kernels.py:40-95: exact kernel of a 9-column mask; row with k active of M=9: (M-1)/(C(M,k) k (M-k))
    from math import comb

    m = np.zeros((9, 9), bool)
    for i in range(9):
        m[i, : i + 1] = True
    k = orc.shap_kernel(m)
    for i in range(8):
        kk = i + 1
        assert abs(k[i] - 8 / (comb(9, kk) * kk * (9 - kk))) < 1e-12
    assert k[8] == 0  # all active -> inf -> 0 (kernels.py:172)
    # tests/test_wlm.py:280-291 pins 1/315 as ~1/305 within 1e-3: k=4 of M=9 -> 8 / (126*4*5) = 1/315
    assert abs(k[3] - 1 / 315) < 1e-12
    # tests/test_pathways.py:452-494 (aggregate = mean of member importances)
    cv = torch.tensor([0.1, 0.4, 0.2, 0.3])
    assert np.allclose(orc.aggregate(cv, [[0, 1], [2], [1, 2, 3]]), [0.25, 0.2, 0.3])
    # tests/test_data.py:1761-1845 (build_edge_mask): block-diagonal offsets + both-endpoints-active rule
    ei = np.array([[0, 1, 2], [1, 2, 0]])
    keep, edges = orc.build_edge_mask(ei, np.array([[1, 1, 0], [1, 1, 1]], bool))
    assert edges.tolist() == [[0, 1, 2, 3, 4, 5], [1, 2, 0, 4, 5, 3]]
    assert keep.tolist() == [True, False, False, True, True, True]
    # tests/test_wlm.py:378-404 (weighted_mse_loss on 1-D inputs) and :300-337 (regularizer)
    p, t, w = torch.tensor([1.0, 2.0]), torch.tensor([0.0, 4.0]), torch.tensor([1.0, 3.0], dtype=torch.float64)
    assert abs(float(orc.weighted_mse_loss(p, t, w)) - ((1 * 1 + 3 * 4) / 2) / 4) < 1e-12
    assert abs(float(orc.regularizer(torch.tensor([[1.0, -3.0]]), 0.5)) - 1.0) < 1e-7


def test_oracle_khop_matches_reference_test_counts():
    """tests/test_explainer.py:647: the 2-hop computational graph around node "10" has 15 nodes."""
    case = gio.load_case("c1_homo_gcn")
    subset, sub_ei, sub_ind, _ = orc.k_hop_subgraph(case["edge_index"].numpy(), 10, 2)
    assert len(subset) == 15 and sub_ei.shape[1] == 34 and subset[sub_ind] == 10
    subset1, _, _, _ = orc.k_hop_subgraph(case["edge_index"].numpy(), 10, 1)
    assert subset1.tolist() == [9, 10, 11, 19, 28]  # tests/test_data.py:700-709
