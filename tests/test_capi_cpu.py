"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/infgen_b200.h declares,
the packed weight layout is consistent with the packer, and the engine refuses to run without a GPU."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import torch

from infgen_b200 import _capi
from infgen_b200.weights import make_state_dict, pack_state_dict, gemm_pack, agent_decoder_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from infgen_b200.build import build
    build()
    return _capi.load()


def test_header_symbols_are_exported_and_bound(lib):
    header = open(os.path.join(ROOT, 'include', 'infgen_b200.h')).read()
    declared = set(re.findall(r'\b(infgen_[a-z_0-9]+)\s*\(', header))
    assert declared, 'no declarations found'
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    assert lib.infgen_abi_version() == _capi.ABI_VERSION
    assert f'#define INFGEN_ABI_VERSION {_capi.ABI_VERSION}' in header


def test_weight_layout_and_packing(lib):
    sd = make_state_dict(0)
    blob = pack_state_dict(sd, lib)
    assert blob.dtype == np.float32 and blob.size == lib.infgen_weight_blob_floats()
    # offsets are 128-byte aligned, non-overlapping and in order
    prev_end = 0
    for i in range(lib.infgen_weight_count()):
        name = lib.infgen_weight_name(i)
        off, n = lib.infgen_weight_offset(name), lib.infgen_weight_numel(name)
        assert off % 32 == 0 and off >= prev_end
        prev_end = off + n
    assert lib.infgen_weight_offset(b'no_such_tensor') == -1
    # a packed Linear can be read back: W[k][n] lives at [k//4][n][k%4]
    w = sd['t_attn_layers.2.to_out.weight'].numpy()
    off = lib.infgen_weight_offset(b't_attn_layers.2.w_out')
    packed = blob[off:off + 128 * 128].reshape(32, 128, 4)
    assert np.array_equal(packed.transpose(0, 2, 1).reshape(128, 128), w.T)
    # missing tensors are reported, not silently zero
    bad = dict(sd)
    del bad['a2a_attn_layers.0.to_q.weight']
    with pytest.raises(KeyError):
        pack_state_dict(bad, lib)


def test_gemm_pack_padding():
    w = np.arange(3 * 5, dtype=np.float32).reshape(3, 5)        # N=3, K=5 -> K4=2, N_pad=4
    p = gemm_pack(w, 4).reshape(2, 4, 4)
    for n in range(3):
        for k in range(5):
            assert p[k // 4, n, k % 4] == w[n, k]
    assert p[1, :, 1:].sum() == 0 and p[:, 3].sum() == 0


def test_state_dict_spec_covers_reference_names():
    spec = agent_decoder_spec()
    for key in ('t_attn_layers.5.to_k_r.weight', 'pt2a_attn_layers.0.attn_prenorm_x_dst.bias',
                'token_predict_head.mlp.3.weight', 'r_t_emb.mlps.3.0.weight', 'fusion_emb.mlp.6.bias'):
        assert key in spec
    assert 't_attn_layers.0.to_k.bias' not in spec                # layers.py:33 bias=False


@pytest.mark.skipif(torch.cuda.is_available(), reason='only meaningful without a GPU')
def test_engine_fails_loudly_without_gpu(lib):
    from infgen_b200.agent_decoder import B200AgentDecoder
    with pytest.raises(RuntimeError, match='no CUDA device|CUDA'):
        B200AgentDecoder(make_state_dict(0))
