"""Operator-level entry points mirroring the reference's `infgen/modules/layers.py` modules.

Each function runs the same sm_100a kernels the decode loop uses, through the C ABI, on host tensors; they exist so
that parity tests can read like tests of the reference modules (`AttentionLayer(x, r, edge_index)` etc.).
"""
import ctypes as C
from typing import Optional, Tuple, Union
import numpy as np
import torch

from . import _capi


def _f32(t: torch.Tensor) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))


def attention_layer(dec, layer: str, x: Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]],
                    r: Optional[torch.Tensor], edge_index: torch.Tensor) -> torch.Tensor:
    """`AttentionLayer.forward(x, r, edge_index)` (layers.py:61-76); `layer` is the state_dict prefix.
    edge_index is PyG style: row 0 = source j, row 1 = target i."""
    if isinstance(x, torch.Tensor):
        x_src, x_dst = None, x
    else:
        x_src, x_dst = x
    n_dst = x_dst.shape[0]
    src, dst = edge_index[0].long(), edge_index[1].long()
    order = torch.argsort(dst, stable=True)
    counts = torch.bincount(dst, minlength=n_dst)
    ptr = np.zeros(n_dst + 1, dtype=np.int32)
    ptr[1:] = np.cumsum(counts.numpy())
    src_sorted = np.ascontiguousarray(src[order].numpy().astype(np.int32))
    r_sorted = _f32(r[order]) if r is not None else None
    xs = _f32(x_src) if x_src is not None else None
    xd = _f32(x_dst)
    out = np.zeros((n_dst, 128), dtype=np.float32)
    _capi.check(dec.lib.infgen_op_attention_layer(
        dec._h, layer.encode(), _capi.f32p(xs), 0 if xs is None else xs.shape[0], _capi.f32p(xd), n_dst,
        _capi.f32p(r_sorted), _capi.i32p(ptr), _capi.i32p(src_sorted), _capi.f32p(out)))
    return torch.from_numpy(out)


def fourier_embedding(dec, name: str, x: torch.Tensor, cat: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`FourierEmbedding.forward(continuous_inputs, categorical_embs)` (layers.py:142-160); `cat` = summed embs."""
    xn = _f32(x)
    cn = _f32(cat) if cat is not None else None
    out = np.zeros((xn.shape[0], 128), dtype=np.float32)
    _capi.check(dec.lib.infgen_op_fourier_embedding(dec._h, name.encode(), _capi.f32p(xn), xn.shape[0], xn.shape[1],
                                                    _capi.f32p(cn), _capi.f32p(out)))
    return torch.from_numpy(out)


def mlp_embedding(dec, name: str, x: torch.Tensor) -> torch.Tensor:
    """`MLPEmbedding.forward` (layers.py:189)."""
    xn = _f32(x)
    out = np.zeros((xn.shape[0], 128), dtype=np.float32)
    _capi.check(dec.lib.infgen_op_mlp_embedding(dec._h, name.encode(), _capi.f32p(xn), xn.shape[0], xn.shape[1],
                                                _capi.f32p(out)))
    return torch.from_numpy(out)


def mlp_layer(dec, name: str, x: torch.Tensor, n_out: int) -> torch.Tensor:
    """`MLPLayer.forward` (layers.py:213-215) for the heads that take a 128-d input."""
    xn = _f32(x)
    out = np.zeros((xn.shape[0], n_out), dtype=np.float32)
    _capi.check(dec.lib.infgen_op_mlp_layer(dec._h, name.encode(), _capi.f32p(xn), xn.shape[0], _capi.f32p(out)))
    return torch.from_numpy(out)
