"""ORACLE / TEST INFRASTRUCTURE ONLY.

CPU restatement of ``torch_geometric.utils.subgraph`` (PyG 2.0.4) for the two
functions the reference calls (``data.py:331-333``, ``model.py:52``).
"""
import torch


def get_num_hops(model):
    """Number of ``MessagePassing`` sub-modules (PyG 2.0.4 semantics)."""
    from ..nn import MessagePassing

    return sum(1 for m in model.modules() if isinstance(m, MessagePassing))


def k_hop_subgraph(node_idx, num_hops, edge_index, relabel_nodes=False, num_nodes=None,
                   flow="source_to_target"):
    """k-hop in-neighbourhood + induced subgraph.

    flow = source_to_target: repeatedly add the *sources* of edges whose
    *target* lies in the previous frontier; then keep every edge with both ends
    in the node set, in original order.  ``subset`` is sorted ascending
    (``torch.unique``), ``inv`` is the rank of the seed(s) in it.
    """
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    assert flow in ("source_to_target", "target_to_source")
    if flow == "target_to_source":
        row, col = edge_index
    else:
        col, row = edge_index

    node_mask = row.new_empty(num_nodes, dtype=torch.bool)
    if isinstance(node_idx, (int, list, tuple)):
        node_idx = torch.tensor([node_idx], device=row.device, dtype=torch.long).flatten()
    else:
        node_idx = node_idx.to(row.device).flatten()

    subsets = [node_idx]
    for _ in range(num_hops):
        node_mask.fill_(False)
        node_mask[subsets[-1]] = True
        edge_mask = node_mask[row]
        subsets.append(col[edge_mask])

    subset, inv = torch.cat(subsets).unique(return_inverse=True)
    inv = inv[: node_idx.numel()]

    node_mask.fill_(False)
    node_mask[subset] = True
    edge_mask = node_mask[row] & node_mask[col]
    edge_index = edge_index[:, edge_mask]

    if relabel_nodes:
        relabel = row.new_full((num_nodes,), -1)
        relabel[subset] = torch.arange(subset.size(0), device=row.device)
        edge_index = relabel[edge_index]

    return subset, edge_index, inv, edge_mask
