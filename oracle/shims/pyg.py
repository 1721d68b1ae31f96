"""Minimal torch_geometric stand-in (only what infgen/modules/{layers,agent_decoder}.py touch).

TEST INFRASTRUCTURE ONLY - see oracle/shims/__init__.py for the semantics this fixes.
"""
import copy
import inspect
import torch
import torch.nn as nn


class MessagePassing(nn.Module):
    """aggr='add', node_dim=0, flow='source_to_target' (edge_index[0]=source j, edge_index[1]=target i)."""

    def __init__(self, aggr='add', node_dim=0, **kwargs):
        super().__init__()
        assert aggr == 'add' and node_dim == 0
        self._msg_params = None

    def propagate(self, edge_index, size=None, **kwargs):
        if self._msg_params is None:
            self._msg_params = list(inspect.signature(self.message).parameters.keys())
            self._upd_params = list(inspect.signature(self.update).parameters.keys())[1:]
        src, dst = edge_index[0], edge_index[1]
        n_dst = kwargs['q'].size(0) if 'q' in kwargs else int(dst.max()) + 1
        msg_kwargs = {}
        for name in self._msg_params:
            if name == 'index':
                msg_kwargs[name] = dst
            elif name == 'ptr':
                msg_kwargs[name] = None
            elif name == 'size_i':
                msg_kwargs[name] = n_dst
            elif name.endswith('_i'):
                msg_kwargs[name] = kwargs[name[:-2]][dst]
            elif name.endswith('_j'):
                msg_kwargs[name] = kwargs[name[:-2]][src]
            else:
                msg_kwargs[name] = kwargs.get(name)
        self._num_dst = n_dst
        msg = self.message(**msg_kwargs)
        out = msg.new_zeros((n_dst,) + tuple(msg.shape[1:]))
        out.index_add_(0, dst, msg)
        upd_kwargs = {name: kwargs.get(name) for name in self._upd_params}
        return self.update(out, **upd_kwargs)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """torch_geometric.utils.softmax: grouped by `index`, max-subtracted, denominator + 1e-16."""
    assert dim == 0
    n = int(index.max()) + 1 if num_nodes is None and index.numel() > 0 else (num_nodes or 0)
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    src_max = src.new_full(shape, float('-inf')).scatter_reduce(0, idx, src.detach(), 'amax', include_self=True)
    out = (src - src_max.gather(0, idx)).exp()
    out_sum = src.new_zeros(shape).scatter_add_(0, idx, out) + 1e-16
    return out / out_sum.gather(0, idx)


def dense_to_sparse(adj, mask=None):
    assert mask is None
    if adj.dim() == 2:
        idx = adj.nonzero().t()
        return idx, adj[idx[0], idx[1]]
    assert adj.dim() == 3
    b, r, c = adj.shape
    nz = adj.nonzero()
    row = nz[:, 0] * r + nz[:, 1]
    col = nz[:, 0] * c + nz[:, 2]
    return torch.stack([row, col], dim=0), adj[nz[:, 0], nz[:, 1], nz[:, 2]]


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None, return_edge_mask=False):
    assert subset.dtype == torch.bool and not relabel_nodes
    edge_mask = subset[edge_index[0]] & subset[edge_index[1]]
    edge_index = edge_index[:, edge_mask]
    if edge_attr is not None:
        edge_attr = edge_attr[edge_mask]
    if return_edge_mask:
        return edge_index, edge_attr, edge_mask
    return edge_index, edge_attr


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros(n, dtype=dtype or torch.float, device=index.device)
    return out.scatter_add_(0, index, out.new_ones(index.size(0)))


def coalesce(edge_index, *a, **k):
    return torch.unique(edge_index, dim=1)


class _Store(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class HeteroData(dict):
    """Nested dict with attribute access on node stores, `.num_graphs == 1`, deep `.clone()`."""
    num_graphs = 1

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _Store):
            v = _Store(v)
        super().__setitem__(k, v)

    def __missing__(self, k):
        s = _Store()
        super().__setitem__(k, s)
        return s

    def clone(self):
        def _c(v):
            if isinstance(v, torch.Tensor):
                return v.clone()
            if isinstance(v, dict):
                return type(v)({kk: _c(vv) for kk, vv in v.items()})
            return copy.deepcopy(v)
        out = HeteroData()
        for k, v in self.items():
            dict.__setitem__(out, k, _c(v))
        return out


class Batch(HeteroData):
    pass
