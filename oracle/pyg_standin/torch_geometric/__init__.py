"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Minimal CPU stand-in for the third-party dependency ``torch_geometric==2.0.4``
(pinned by the reference in ``Dockerfile:12-16`` / ``dev_setup.sh:4-9``), which
is not installed in this image and is not part of ``/root/reference``.

Only the call sites of the reference's hot path are restated, from PyG 2.0.4's
published algorithms:

* ``torch_geometric.utils.subgraph.k_hop_subgraph``  (reference: ``data.py:7,331``)
* ``torch_geometric.utils.subgraph.get_num_hops``    (reference: ``model.py:4,52``)
* ``torch_geometric.nn.{MessagePassing,GCNConv,SAGEConv,HeteroConv,Linear}``
  (reference fixtures: ``tests/test_utils.py:7,46,62,133-139``)

Parity status: the conv arithmetic is *unpinned* by the reference's own tests
(they only assert 0<=y<=1, ``tests/test_model.py:78-79``); the k-hop node sets
are pinned by ``tests/test_data.py:700-1168`` and ``tests/test_pathways.py:43-141``.
"""
from . import nn, utils  # noqa: F401

__version__ = "2.0.4-standin"
