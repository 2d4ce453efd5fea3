"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *unmodified* reference package from ``/root/reference/src`` in this
container, or from its staged copy ``oracle/_ref/src`` on the GPU box (``oracle/stage_ref.py``: byte-for-byte, digests
in ``oracle/_ref/MANIFEST.json``), behind four shims:

1. ``torch_geometric`` -> ``oracle/pyg_standin`` (PyG 2.0.4 is not installed);
2. ``inspect.getargspec`` (removed in Python 3.11; used at ``model.py:104``);
3. ``ReduceLROnPlateau(verbose=...)`` (rejected by torch 2.11; ``wlm.py:484-486``;
   the scheduler is never stepped, so results are unaffected);
4. CUDA hidden, so the numpy ``in1d`` branch of ``data.py:431-448`` is taken.

Used by ``oracle/make_golden.py`` to generate the fixtures under ``tests/golden``
and by ``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is
absent) to pin the oracle port against the reference itself.
"""
import inspect
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree of the build container, else the byte-for-byte staged copy (oracle/stage_ref.py) that travels to the GPU box
REFERENCE_SRC = "/root/reference/src"
if not os.path.isdir(os.path.join(REFERENCE_SRC, "pathway_explanations")):
    REFERENCE_SRC = os.path.join(_HERE, "_ref", "src")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "pathway_explanations"))


def install_pyg_standin():
    p = os.path.join(_HERE, "pyg_standin")
    if p not in sys.path:
        sys.path.insert(0, p)


def import_reference():
    """Return the reference's ``pathway_explanations`` package (CPU only)."""
    assert reference_available(), "reference tree not present (GPU box?)"
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    import torch

    assert not torch.cuda.is_available(), (
        "run the reference harness with CUDA_VISIBLE_DEVICES='' (cupy is absent; data.py:11-13)"
    )
    install_pyg_standin()
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec  # model.py:104 only reads .args
    sched = torch.optim.lr_scheduler
    if not getattr(sched.ReduceLROnPlateau, "_xpgnn_shim", False):
        _orig = sched.ReduceLROnPlateau

        class ReduceLROnPlateau(_orig):
            _xpgnn_shim = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        sched.ReduceLROnPlateau = ReduceLROnPlateau
    # make sure *our* drop-in package of the same name is not picked up
    for name in [m for m in sys.modules if m == "pathway_explanations" or m.startswith("pathway_explanations.")]:
        mod = sys.modules[name]
        if REFERENCE_SRC not in (getattr(mod, "__file__", "") or ""):
            del sys.modules[name]
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import pathway_explanations  # noqa: E402

    assert REFERENCE_SRC in pathway_explanations.__file__
    return pathway_explanations


class Recorder:
    """Monkeypatch hooks that record the intermediates of ``Explainer.run``."""

    def __init__(self, ref):
        import pathway_explanations.data as rdata
        import pathway_explanations.explainer as rexp
        import pathway_explanations.masks as rmasks
        import pathway_explanations.wlm as rwlm

        self.comp_graph = []
        self.masks = []
        self.batches = []
        self.weights = []
        self.init_weights = []
        rec = self

        orig_cg = rdata.Data.comp_graph

        def comp_graph(self_, *a, **k):
            out = orig_cg(self_, *a, **k)
            rec.comp_graph.append(out)
            return out

        rdata.Data.comp_graph = comp_graph

        orig_mg = rmasks.Mask.mask_generator

        def mask_generator(self_):
            loader, rows = orig_mg(self_)
            rec.masks.append((loader.dataset.clone(), None if rows is None else rows.clone(),
                              loader.batch_size))
            return loader, rows

        rmasks.Mask.mask_generator = mask_generator

        orig_ko = rwlm.kernel_output

        def kernel_output(mask, *a, **k):
            kern, out = orig_ko(mask, *a, **k)
            rec.batches.append((mask.clone(), kern.clone(), out.clone()))
            return kern, out

        rwlm.kernel_output = kernel_output

        orig_tm = rwlm.train_model

        def train_model(loader, params, feat, ei, lin, *a, **k):
            rec.init_weights.append(lin.layer.weight.detach().clone().flatten())
            w, losses, best = orig_tm(loader, params, feat, ei, lin, *a, **k)
            rec.weights.append((w[0].detach().clone(), list(losses), best))
            return w, losses, best

        rwlm.train_model = train_model
        rexp.train_model = train_model
        self._undo = [
            (rdata.Data, "comp_graph", orig_cg),
            (rmasks.Mask, "mask_generator", orig_mg),
            (rwlm, "kernel_output", orig_ko),
            (rwlm, "train_model", orig_tm),
            (rexp, "train_model", orig_tm),
        ]

    def close(self):
        for obj, name, orig in self._undo:
            setattr(obj, name, orig)
