"""Helpers the reference imports from torch_geometric.utils (never on the timed path)."""
import torch


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros(n, dtype=dtype or torch.float32)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype))


def to_networkx(data, node_attrs=None, edge_attrs=None, to_undirected=False):
    import networkx as nx
    g = nx.Graph() if to_undirected else nx.DiGraph()
    g.add_nodes_from(range(int(data.num_nodes)))
    ei = data.edge_index.tolist()
    vals = {k: getattr(data, k).tolist() for k in (edge_attrs or [])}
    for e, (u, v) in enumerate(zip(ei[0], ei[1])):
        g.add_edge(u, v, **{k: vals[k][e] for k in vals})
    return g


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    import numpy as np
    import scipy.sparse as sp
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    w = np.ones(edge_index.size(1)) if edge_attr is None else edge_attr.view(-1).numpy()
    return sp.coo_matrix((w, (edge_index[0].numpy(), edge_index[1].numpy())), shape=(n, n))
