"""Ensembles of independent rollouts (BASELINE config 5): a block-diagonal batch per GPU, no communication.

`collate` concatenates graphs the way the reference's loader does for a batch (data_loader.py:113-162 / PyG `Batch`: node
features stacked per type, edge indices offset by the node counts of the preceding graphs, edge attributes stacked); since
no edge joins two graphs, one step of the batch is exactly one step of every member.  Members may differ in size, in
topology and in `span` (generate mode picks the span by the nearest (G, R) grid point, graph_trajectory.py:1314-1316):
`EnsembleEngine.step(spans)` advances every member by its own span (z += span_i / 121, test.py:401-407 per graph).
"""
import torch

from .engine import DEFAULT_EDGE_TYPES, RolloutEngine


def collate(graphs, edge_types=DEFAULT_EDGE_TYPES):
    """graphs: list of (x_dict, edge_index_dict, edge_attr_dict or None).  Returns (x, ei, ea or None, ptr) with
    ptr[t] = [0, n_0, n_0 + n_1, ...] node offsets and ptr[e] edge offsets (keys: node types and edge types)."""
    node_types = list(graphs[0][0])
    ptr = {t: [0] for t in node_types}
    for x, _, _ in graphs:
        for t in node_types:
            ptr[t].append(ptr[t][-1] + x[t].shape[0])
    x = {t: torch.cat([g[0][t] for g in graphs], 0) for t in node_types}
    ei, ea = {}, {}
    with_ea = all(g[2] is not None for g in graphs)
    for e in edge_types:
        parts, ptr[e] = [], [0]
        for i, (_, eidx, _) in enumerate(graphs):
            off = torch.tensor([[ptr[e[0]][i]], [ptr[e[2]][i]]], dtype=eidx[e].dtype, device=eidx[e].device)
            parts.append(eidx[e] + off)
            ptr[e].append(ptr[e][-1] + eidx[e].shape[1])
        ei[e] = torch.cat(parts, 1)
        if with_ea:
            ea[e] = torch.cat([g[2][e].reshape(-1, 1) for g in graphs], 0)
    return x, ei, (ea if with_ea else None), ptr


class EnsembleEngine(RolloutEngine):
    """RolloutEngine over a block-diagonal batch; `step(spans)` takes one span per member (or a scalar for all)."""

    def set_graphs(self, graphs):
        x, ei, ea, ptr = collate(graphs, self.edge_types)
        self.graph_ptr = ptr
        self.__dict__.pop('_dz_cache', None)
        self.set_graph(x, ei, ea)
        return ptr

    def split(self, pred):
        """Per-member views of a prediction dict ('joint', 'grain', 'grain_area' by node, 'edge_event' / 'edge' by jj edge)."""
        keys = {'joint': 'joint', 'grain': 'grain', 'grain_area': 'grain', 'edge_event': self.edge_types[2], 'edge': self.edge_types[2]}
        n = len(self.graph_ptr['grain']) - 1
        out = []
        for i in range(n):
            out.append({k: v[self.graph_ptr[keys[k]][i]:self.graph_ptr[keys[k]][i + 1]] for k, v in pred.items()
                        if k in keys and v is not None})
        return out

    def member_features(self, i):
        return {t: self.x[t][self.graph_ptr[t][i]:self.graph_ptr[t][i + 1]] for t in self.node_types}
