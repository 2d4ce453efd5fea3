"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the reference's perturbation hot path plus the harness that runs the
unmodified reference in the build container.  Importing this package makes the name
``torch_geometric`` resolve to the CPU stand-in under ``oracle/pyg_standin`` unless a real
PyG is installed (it is not in this image).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

if _ilu.find_spec("torch_geometric") is None:
    _sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pyg_standin"))
