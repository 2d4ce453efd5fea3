"""Import alias so that reference user code (``from pathway_explanations.explainer import Explainer,
set_seed``; reference ``README.md:81``) runs unchanged on the B200-native engine."""
import sys as _sys

import bikg_graph_explainability_public_b200 as _impl
from bikg_graph_explainability_public_b200 import *  # noqa: F401,F403
from bikg_graph_explainability_public_b200 import (data, explainer, kernels, masks, model,  # noqa: F401
                                                  pathways, wlm)

__all__ = _impl.__all__
for _m in ("data", "explainer", "kernels", "masks", "model", "pathways", "wlm"):
    _sys.modules[__name__ + "." + _m] = getattr(_impl, _m)
