"""ORACLE / TEST INFRASTRUCTURE ONLY."""
from .subgraph import get_num_hops, k_hop_subgraph  # noqa: F401
