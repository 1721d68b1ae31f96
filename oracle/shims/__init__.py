"""Import stand-ins that let the UNMODIFIED reference (`/root/reference/infgen/modules/*`) run on CPU
in the build container, where torch_geometric / torch_cluster / lightning / easydict are not installed.

TEST INFRASTRUCTURE ONLY.  Nothing under `infgen_b200/` imports this package; it exists so that
`tests/golden/make_golden.py` can execute the real reference and write golden vectors, and so that
`oracle/agent_decoder_oracle.py` (the travelling CPU restatement) can be pinned against it.

The stand-ins *define* the third-party semantics parity is judged against:

* ``torch_cluster.radius``: strict ``dist^2 < r^2``, same batch id, the first ``max_num_neighbors``
  x-points by ascending index per y-point (torch_cluster 1.6.3 CUDA kernel behaviour),
  output rows ``(y_idx, x_idx)`` ordered by y then x.
* ``torch_cluster.radius_graph``: ``radius(x, x, max_num_neighbors + 1)`` with self loops dropped and
  rows flipped to ``(source=x_idx, target=y_idx)`` (flow='source_to_target').
* ``torch_geometric.utils.softmax``: per-target max-subtracted exp / (sum + 1e-16).
* ``torch_geometric.utils.dense_to_sparse`` on a [B,R,C] mask: row-major ``nonzero`` order, node offset b*R.
* ``MessagePassing.propagate(aggr='add', node_dim=0)``: ``_j`` args gathered by edge_index[0], ``_i`` by
  edge_index[1], messages summed with ``index_add_`` in edge order.
"""
import sys
import types
import importlib.machinery
from unittest import mock

import os

# the reference tree: where it lies in the build container, else the copy staged by oracle/make_ref.py (GPU box)
REFERENCE_ROOT = '/root/reference'
if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'infgen', 'modules')):
    REFERENCE_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), '_ref')


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []  # behave as a package so sub-imports resolve through sys.modules
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-ins and put the reference tree on sys.path. Idempotent."""
    if getattr(install, '_done', False):
        return
    import torch
    from . import pyg, cluster

    # --- torch_geometric -------------------------------------------------------------------------------
    tg = _module('torch_geometric')
    _module('torch_geometric.nn')
    _module('torch_geometric.nn.conv', MessagePassing=pyg.MessagePassing)
    _module('torch_geometric.utils', softmax=pyg.softmax, dense_to_sparse=pyg.dense_to_sparse,
            subgraph=pyg.subgraph, degree=pyg.degree, coalesce=pyg.coalesce)
    _module('torch_geometric.data', HeteroData=pyg.HeteroData, Batch=pyg.Batch, Dataset=object)
    tg.__version__ = '2.5.3-shim'

    # --- torch_cluster ---------------------------------------------------------------------------------
    _module('torch_cluster', radius=cluster.radius, radius_graph=cluster.radius_graph)

    # --- easydict / lightning_utilities (imported by infgen/utils/func.py) --------------------------------
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in {**(d or {}), **kw}.items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        __setattr__ = __setitem__

    _module('easydict', EasyDict=EasyDict)
    _module('lightning_utilities')
    _module('lightning_utilities.core')
    _module('lightning_utilities.core.rank_zero',
            rank_prefixed_message=lambda msg, rank=None: msg,
            rank_zero_only=lambda f: f)

    # --- plotting: agent_decoder.py star-imports infgen.utils.visualization, which needs TF/matplotlib/waymo.
    # None of it is on the decode path (only reachable through PLOT_EDGE* env switches).
    vis = _module('infgen.utils.visualization')
    vis.__all__ = []
    vis.plot_interact_edge = mock.MagicMock()

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install._done = True
