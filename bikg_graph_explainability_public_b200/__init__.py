"""B200-native perturbation engine for community-aware GNN explanations.

Drop-in for the perturbation hot path of ``pathway_explanations`` (XP-GNN): the public names of
the reference package (``__init__.py:1-9``) are re-exported, ``Explainer(...).run(query, repeats)``
keeps its contract, and the work runs in hand-written sm_100a CUDA kernels
(``csrc/`` -> ``libxpgnn_b200.so``, C ABI in ``include/xpgnn_b200.h``).
"""
from .data import Data
from .explainer import Explainer, set_seed
from .kernels import Kernel
from .masks import Mask
from .model import Model
from .pathways import Pathways
from .wlm import LinearRegression

__all__ = ["Data", "Explainer", "Kernel", "Mask", "Model", "Pathways", "LinearRegression", "set_seed"]
__version__ = "0.1.0"
