"""CPU-side tests: the C-ABI library loads and exports what include/xpgnn_b200.h declares, host logic
(row plan, lowering, hetero flattening) and the loud failure without a GPU.  No compute calls."""
import math
import os
import random
import re

import numpy as np
import pytest
import torch

import golden_io as gio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    from bikg_graph_explainability_public_b200 import _lib

    header = open(os.path.join(ROOT, "include", "xpgnn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(xpgnn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.xpgnn_abi_version() == 1


def test_struct_layouts_match_header():
    """ctypes mirrors must have the C layout (64-bit pointers, int32 fields, natural alignment)."""
    import ctypes as C

    from bikg_graph_explainability_public_b200 import _lib

    assert C.sizeof(_lib.MaskPlan) == 4 * 4 + 9 * 8
    assert C.sizeof(_lib.Relation) == 6 * 4 + 5 * 8
    assert C.sizeof(_lib.Layer) == 4 + 4 + 8 + 3 * 4 + 4
    assert C.sizeof(_lib.Dense) == 3 * 4 + 4 + 2 * 8
    assert _lib.Plan.x.offset == 8 and _lib.Plan.layers_host.offset == 24 and _lib.Plan.query.offset == 56


def test_row_plan_follows_torch_float32_arithmetic():
    """masks.py:116-125: int / int64-tensor is reciprocal * int in float32."""
    from bikg_graph_explainability_public_b200.masks import row_plan
    from oracle.xpgnn_oracle import row_plan as oracle_plan

    rnd = random.Random(0)
    for _ in range(300):
        c = rnd.randint(1, 30)
        lens = [rnd.randint(1, rnd.choice([5, 50, 500, 5000])) for _ in range(c)]
        total = rnd.choice([1000, 4096, 64, 91, 16384])
        plan = row_plan(lens, total)
        assert plan == oracle_plan(lens, total)
        lp = torch.tensor(lens)
        for i, ln in enumerate(lens):
            frac = ln / torch.sum(lp)
            size = math.ceil(frac * total)
            si = math.ceil(frac * size)
            if si < 3:
                si, size = 1, 2
            assert plan[i] == (size, si)
    assert row_plan([1, 1, 10], 1000)[2] == (834, 696)  # fl32(len/sum) would give 695


@pytest.mark.parametrize("name", ["c1_homo_gcn", "c2_hetero_gcn", "gcn2_random", "sage2_shapley", "c4_hetero_sage"])
def test_lowering_of_fixture_models(name):
    from bikg_graph_explainability_public_b200.lowering import lower
    from bikg_graph_explainability_public_b200.model import Model

    case = gio.load_case(name)
    arch = gio.build_arch(case)
    m = lower(arch)
    spec = case["meta"]["model"]
    assert len(m.convs) == len(spec["conv_dims"])
    assert [c.out_dim for c in m.convs] == list(spec["conv_dims"])
    assert len(m.head) == len(spec["head_dims"]) - 1 and m.out_dim == 1
    n_rel = len(spec.get("relations", [None]))
    assert Model(arch).get_hops(n_rel if "relations" in spec else 0) == len(spec["conv_dims"])
    for c in m.convs:
        assert len(c.relations) == n_rel and c.act == "relu"


def test_lowering_rejects_unknown_modules():
    from torch import nn

    from bikg_graph_explainability_public_b200 import nn as xnn
    from bikg_graph_explainability_public_b200.lowering import lower

    class Weird(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = xnn.GCNConv(4, 4)
            self.norm = nn.BatchNorm1d(4)

    with pytest.raises(NotImplementedError):
        lower(Weird())
    with pytest.raises(NotImplementedError):
        lower(nn.Sequential(nn.Linear(3, 3)))


def test_product_layers_load_reference_checkpoint_keys():
    """State-dict keys of the product's layer containers equal PyG's (test_data/*.pth.tar layout)."""
    from torch import nn

    from bikg_graph_explainability_public_b200 import nn as xnn

    case = gio.load_case("c2_hetero_gcn")
    rels = [tuple(r) for r in case["meta"]["model"]["relations"]]

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.ModuleList([xnn.HeteroConv({r: xnn.GCNConv(84, 16) for r in rels}), nn.ReLU()])
            self.fc = nn.ModuleList([xnn.Linear(16, 16), nn.ReLU(), xnn.Linear(16, 32), nn.ReLU(), xnn.Linear(32, 1),
                                     nn.Sigmoid()])

    net = Net()
    net.load_state_dict(case["state"])  # strict: every key must match
    homo = gio.load_case("c1_homo_gcn")

    class Homo(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.ModuleList([xnn.GCNConv(84, 16), nn.ReLU()])
            self.fc = nn.ModuleList([xnn.Linear(16, 16), nn.ReLU(), xnn.Linear(16, 32), nn.ReLU(), xnn.Linear(32, 1),
                                     nn.Sigmoid()])

    Homo().load_state_dict(homo["state"])


def test_hetero_flatten_matches_oracle():
    from bikg_graph_explainability_public_b200.data import Data
    from oracle.xpgnn_oracle import flatten_hetero

    case = gio.load_case("c4_hetero_sage")
    out = Data(case["feat"], case["edge_index"]).preprocess_hetero_graph()
    ref = flatten_hetero(case["feat"], case["edge_index"])
    assert out[0] == ref[0] and out[1] == ref[1]
    for a, b in zip(out[2:6], ref[2:6]):
        assert torch.equal(a, b)
    assert out[6] == ref[6] and out[7] == ref[7] and out[8] == ref[8]


def test_explainer_assertions_and_loud_failure_without_gpu():
    from bikg_graph_explainability_public_b200 import Explainer, _lib

    case = gio.load_case("c1_homo_gcn")
    names, pathways, pnames = gio.fresh_inputs(case)
    arch = gio.build_arch(case)
    p = dict(case["meta"]["params"])
    with pytest.raises(AssertionError, match="Feature matrix is not torch tensor or dict"):
        Explainer([1, 2], case["edge_index"], arch, p, names, pathways, pnames)
    with pytest.raises(AssertionError, match="Length of list with pathway names"):
        Explainer(case["feat"], case["edge_index"], arch, p, names, pathways, pnames[:-1])
    with pytest.raises(AssertionError, match="Feature given is not a dict of node types"):
        Explainer(case["feat"], case["edge_index"], arch, p, names, pathways, pnames, element_type="gene")
    ex = Explainer(case["feat"], case["edge_index"], arch, p, names, pathways, pnames, problem="node")
    if not torch.cuda.is_available():
        with pytest.raises(_lib.XpgnnError, match="no CPU fallback"):
            ex.run("10", 1)


def test_word_range_partitions_tiles():
    from bikg_graph_explainability_public_b200.shard import word_range

    for n_words in (0, 1, 5, 16, 129):
        for world in (1, 2, 3, 8):
            spans = [word_range(n_words, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_words
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _pathways_numpy_formulation(communities, names):
    """The reference's algorithm (pathways.py:84-96, 131-134): intersect1d against the whole name array per community."""
    names_array = np.array(names, dtype=str)
    sub, inds = [], []
    for community in communities:
        common = np.intersect1d(np.array(community, dtype=str), names_array)
        if len(common) > 0:
            sub.append(common.tolist())
    for community in sub:
        _, ind, _ = np.intersect1d(names_array, np.array(community, dtype=str), return_indices=True)
        inds.append(ind.tolist())
    return sub, inds


@pytest.mark.parametrize("case", ["str", "int_members", "duplicates", "unicode"])
def test_name_resolution_fast_path_equals_intersect1d(case):
    """the hash-map name resolution returns exactly what the reference's per-community
    ``np.intersect1d`` returns (members as sorted unique names, indices of first occurrences in name order)."""
    from bikg_graph_explainability_public_b200.pathways import Pathways

    rng = np.random.default_rng(3)
    n = 500
    if case == "unicode":
        names = ["gène_%d" % i if i % 3 else "Ωmega%d" % i for i in range(n)]
    else:
        names = [str(i) for i in range(n)]
    sub_names = [names[i] for i in rng.permutation(n)[:300]]            # the computational graph keeps 300 nodes
    if case == "duplicates":
        sub_names = sub_names + sub_names[:25]                             # repeated names: first occurrence wins
    communities = []
    for c in range(12):
        members = rng.integers(0, n, size=int(rng.integers(1, 60))).tolist()  # with repeats, partly outside the subgraph
        communities.append(members if case == "int_members" else [names[i] for i in members])
    communities.append([names[i] for i in range(n) if names[i] not in set(sub_names)][:5])  # dropped entirely
    cnames = ["c%d" % i for i in range(len(communities))]
    ref_sub, ref_inds = _pathways_numpy_formulation(communities, sub_names)
    sub, sub_cnames, _ = Pathways(communities, cnames).comp_graph(sub_names)
    assert sub == ref_sub
    assert len(sub_cnames) == len(sub) == 12
    assert Pathways(sub, sub_cnames).names2inds(sub_names) == ref_inds
    # what Explainer.run does: reuse the index of comp_graph, skip the (identity) intersection of the filtered communities
    pw = Pathways(communities, cnames)
    sub2, sub_cnames2, _ = pw.comp_graph(sub_names)
    assert Pathways(sub2, sub_cnames2).names2inds(sub_names, index=pw.last_index, filtered=True) == ref_inds


def test_resolve_indices_equals_comp_graph_then_names2inds():
    """``Pathways.resolve_indices`` (the vectorised pass ``Explainer.run`` uses) returns, per community, the same index
    set as ``comp_graph`` followed by ``names2inds`` (the reference's two intersect1d passes, pathways.py:84-96,131-134),
    in ascending order -- the order ``Mask.mask_generator`` sorts them into anyway (masks.py:323) -- incl. duplicate
    names (first occurrence wins), duplicate members, communities that miss the subgraph and integer members."""
    import random

    from bikg_graph_explainability_public_b200.pathways import Pathways

    rnd = random.Random(3)
    names = ["n%d" % rnd.randrange(300) for _ in range(400)]
    coms = [["n%d" % rnd.randrange(500) for _ in range(rnd.randrange(1, 40))] for _ in range(30)] + [["zz1", "zz2"]]
    cn = ["c%d" % i for i in range(len(coms))]
    sub, sn, _ = Pathways([list(c) for c in coms], cn).comp_graph(names)
    want = [sorted(x) for x in Pathways(sub, sn).names2inds(names)]
    got, got_names = Pathways([list(c) for c in coms], cn).resolve_indices(names)
    assert got_names == sn and got == want and "c30" not in got_names
    coms_i = [[rnd.randrange(500) for _ in range(20)] for _ in range(10)]
    names_i = [str(i) for i in range(250)]
    sub, sn, _ = Pathways([list(c) for c in coms_i], None).comp_graph(names_i)
    want = [sorted(x) for x in Pathways(sub, sn).names2inds(names_i)]
    got, got_names = Pathways([list(c) for c in coms_i], None).resolve_indices(names_i)
    assert got_names == sn and got == want
