"""Weight naming, synthetic initialisation and packing for the decode path.

The source of truth for weights is the reference module's `state_dict()`; names and shapes below are those of
`InfGenAgentDecoder` (reference `infgen/modules/agent_decoder.py:187-290`, `infgen/modules/layers.py:32-58,
126-139, 170-177, 206-211`), verified against a live instance under `oracle/shims` (tests/golden/make_golden.py
loads a dict produced here with `load_state_dict(strict=True)`).
"""
from collections import OrderedDict
from typing import Dict, Tuple
import numpy as np
import torch

from .config import HIDDEN, NUM_HEADS, HEAD_DIM, NUM_FREQ, FOURIER_IN, NUM_LAYERS, SEED_LAYERS, TOKEN_SIZE

GRID_SIZE = 1961     # Attr_Tokenizer(grid_range=150, grid_interval=3, radius=75).grid_size (attr_tokenizer.py:24-43)
ANGLE_SIZE = 120     # 360 / angle_interval (attr_tokenizer.py:20)

NON_BIPARTITE_STACKS = ('t_attn_layers', 'a2a_attn_layers', 'a2sa_attn_layers')   # agent_decoder.py:222-243

Spec = "OrderedDict[str, Tuple[Tuple[int, ...], str]]"   # name -> (shape, kind); kind in {linear_w, bias, ln_w, ln_b, emb}


def _linear(spec, prefix, n_in, n_out, bias=True):
    spec[f'{prefix}.weight'] = ((n_out, n_in), 'linear_w')
    if bias:
        spec[f'{prefix}.bias'] = ((n_out,), 'bias')


def _ln(spec, prefix, n=HIDDEN):
    spec[f'{prefix}.weight'] = ((n,), 'ln_w')
    spec[f'{prefix}.bias'] = ((n,), 'ln_b')


def _attention_layer(spec, prefix, has_pos_emb=True):
    """layers.py:32-58. `attn_prenorm_x_dst` aliases `_x_src` when not bipartite but is still a state_dict key."""
    d = NUM_HEADS * HEAD_DIM
    _linear(spec, f'{prefix}.to_q', HIDDEN, d)
    _linear(spec, f'{prefix}.to_k', HIDDEN, d, bias=False)
    _linear(spec, f'{prefix}.to_v', HIDDEN, d)
    if has_pos_emb:
        _linear(spec, f'{prefix}.to_k_r', HIDDEN, d, bias=False)
        _linear(spec, f'{prefix}.to_v_r', HIDDEN, d)
    _linear(spec, f'{prefix}.to_s', HIDDEN, d)
    _linear(spec, f'{prefix}.to_g', d + HIDDEN, d)
    _linear(spec, f'{prefix}.to_out', d, HIDDEN)
    _linear(spec, f'{prefix}.ff_mlp.0', HIDDEN, 4 * HIDDEN)
    _linear(spec, f'{prefix}.ff_mlp.3', 4 * HIDDEN, HIDDEN)
    _ln(spec, f'{prefix}.attn_prenorm_x_src')
    _ln(spec, f'{prefix}.attn_prenorm_x_dst')
    if has_pos_emb:
        _ln(spec, f'{prefix}.attn_prenorm_r')
    _ln(spec, f'{prefix}.attn_postnorm')
    _ln(spec, f'{prefix}.ff_prenorm')
    _ln(spec, f'{prefix}.ff_postnorm')


def _fourier(spec, prefix, input_dim):
    """layers.py:126-139."""
    spec[f'{prefix}.freqs.weight'] = ((input_dim, NUM_FREQ), 'emb')
    for d in range(input_dim):
        _linear(spec, f'{prefix}.mlps.{d}.0', FOURIER_IN, HIDDEN)
        _ln(spec, f'{prefix}.mlps.{d}.1')
        _linear(spec, f'{prefix}.mlps.{d}.3', HIDDEN, HIDDEN)
    _ln(spec, f'{prefix}.to_out.0')
    _linear(spec, f'{prefix}.to_out.2', HIDDEN, HIDDEN)


def _mlp_embedding(spec, prefix, input_dim):
    """layers.py:170-177."""
    _linear(spec, f'{prefix}.mlp.0', input_dim, 128)
    _ln(spec, f'{prefix}.mlp.1')
    _linear(spec, f'{prefix}.mlp.3', 128, HIDDEN)
    _ln(spec, f'{prefix}.mlp.4')
    _linear(spec, f'{prefix}.mlp.6', HIDDEN, HIDDEN)


def _mlp_layer(spec, prefix, input_dim, output_dim, hidden_dim=HIDDEN):
    """layers.py:206-211."""
    _linear(spec, f'{prefix}.mlp.0', input_dim, hidden_dim)
    _ln(spec, f'{prefix}.mlp.1', hidden_dim)
    _linear(spec, f'{prefix}.mlp.3', hidden_dim, output_dim)


def agent_decoder_spec() -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """Every key of `InfGenAgentDecoder.state_dict()` for ours_standard / ours_long_term, in module order."""
    s = OrderedDict()
    s['type_a_emb.weight'] = ((4, HIDDEN), 'emb')
    _mlp_embedding(s, 'shape_emb', 3)
    s['state_a_emb.weight'] = ((4, HIDDEN), 'emb')
    _fourier(s, 'x_a_emb', 2)
    _fourier(s, 'r_t_emb', 4)
    _fourier(s, 'r_pt2a_emb', 3)
    _fourier(s, 'r_a2a_emb', 3)
    _fourier(s, 'r_pt2sa_emb', 3)
    _fourier(s, 'r_a2sa_emb', 3)
    for name in ('veh', 'ped', 'cyc'):
        _mlp_embedding(s, f'token_emb_{name}', 8)
    _mlp_embedding(s, 'token_emb_grid', 2)
    s['no_token_emb.weight'] = ((1, HIDDEN), 'emb')
    s['bos_token_emb.weight'] = ((1, HIDDEN), 'emb')
    s['invalid_offset_token_emb.weight'] = ((1, HIDDEN), 'emb')
    _mlp_embedding(s, 'fusion_emb', 4 * HIDDEN)
    for stack, n, pos in (('t_attn_layers', NUM_LAYERS, True), ('pt2a_attn_layers', NUM_LAYERS, True),
                          ('a2a_attn_layers', NUM_LAYERS, True), ('pt2sa_attn_layers', SEED_LAYERS, True),
                          ('a2sa_attn_layers', SEED_LAYERS, True), ('occ2sa_attn_layers', SEED_LAYERS, False)):
        for i in range(n):
            _attention_layer(s, f'{stack}.{i}', has_pos_emb=pos)
    _mlp_layer(s, 'token_predict_head', HIDDEN, TOKEN_SIZE)
    _mlp_layer(s, 'state_predict_head', HIDDEN, 3)
    _mlp_layer(s, 'seed_state_predict_head', HIDDEN, 2)
    _mlp_layer(s, 'seed_type_predict_head', HIDDEN, 3)
    _mlp_layer(s, 'seed_shape_predict_head', HIDDEN, 3)
    _mlp_layer(s, 'seed_pos_rel_token_predict_head', HIDDEN, GRID_SIZE)
    _mlp_layer(s, 'seed_offset_xy_predict_head', HIDDEN, 2)
    _mlp_layer(s, 'seed_agent_occ_embed', GRID_SIZE, HIDDEN)
    _mlp_layer(s, 'seed_heading_rel_token_predict_head', HIDDEN, ANGLE_SIZE)
    _mlp_layer(s, 'grid_agent_occ_head', HIDDEN, GRID_SIZE)
    _mlp_layer(s, 'grid_pt_occ_head', HIDDEN, GRID_SIZE)
    _mlp_layer(s, 'grid_index_head', HIDDEN, GRID_SIZE)
    return s


def map_decoder_spec() -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """Every key of `InfGenMapDecoder.state_dict()` (map_decoder.py:46-64; ours_standard.yaml: input_dim 2, 3 layers), in
    module order.  SURVEY.md section 8f row f1: the next row after the decode path."""
    s = OrderedDict()
    s['type_pt_emb.weight'] = ((17, HIDDEN), 'emb')
    s['side_pt_emb.weight'] = ((4, HIDDEN), 'emb')
    s['polygon_type_emb.weight'] = ((4, HIDDEN), 'emb')
    s['light_pl_emb.weight'] = ((4, HIDDEN), 'emb')
    _fourier(s, 'r_pt2pt_emb', 3)
    for i in range(3):
        _attention_layer(s, f'pt2pt_layers.{i}', has_pos_emb=True)
    _mlp_layer(s, 'token_predict_head', HIDDEN, 1024)
    _mlp_embedding(s, 'token_emb', 22)
    return s


def _fill(spec, seed: int, perturb: bool) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for name, (shape, kind) in spec.items():
        if kind == 'linear_w':
            fan_out, fan_in = shape
            bound = float(np.sqrt(6.0 / (fan_in + fan_out)))
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == 'emb':
            a = rng.normal(0.0, 0.02, size=shape)
        elif kind == 'bias':
            a = rng.normal(0.0, 0.05, size=shape) if perturb else np.zeros(shape)
        elif kind == 'ln_w':
            a = 1.0 + rng.normal(0.0, 0.1, size=shape) if perturb else np.ones(shape)
        elif kind == 'ln_b':
            a = rng.normal(0.0, 0.05, size=shape) if perturb else np.zeros(shape)
        else:
            raise ValueError(kind)
        sd[name] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return sd


def make_map_state_dict(seed: int = 0, perturb: bool = True) -> Dict[str, torch.Tensor]:
    """Seed-fixed synthetic `InfGenMapDecoder` weights (same distributions as `make_state_dict`)."""
    sd = _fill(map_decoder_spec(), seed, perturb)
    for i in range(3):                     # pt2pt layers are not bipartite: the dst norm aliases the src norm
        for leaf in ('weight', 'bias'):
            sd[f'pt2pt_layers.{i}.attn_prenorm_x_dst.{leaf}'] = sd[f'pt2pt_layers.{i}.attn_prenorm_x_src.{leaf}'].clone()
    return sd


def make_state_dict(seed: int = 0, perturb: bool = True) -> Dict[str, torch.Tensor]:
    """Seed-fixed synthetic weights with the reference's names/shapes (there are no checkpoints offline).

    The reference initialiser (`weight_init`, infgen/utils/func.py:177-194) is xavier-uniform Linear weights, zero
    biases, N(0, 0.02) embeddings, identity LayerNorm.  With `perturb=True` biases and LayerNorm affines are
    additionally randomised so parity tests exercise every parameter (a zero bias hides an omitted bias add).
    numpy's PCG64 stream is used so the values are bit-identical in every container.
    """
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for name, (shape, kind) in agent_decoder_spec().items():
        if kind == 'linear_w':
            fan_out, fan_in = shape
            bound = float(np.sqrt(6.0 / (fan_in + fan_out)))
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == 'emb':
            a = rng.normal(0.0, 0.02, size=shape)
        elif kind == 'bias':
            a = rng.normal(0.0, 0.05, size=shape) if perturb else np.zeros(shape)
        elif kind == 'ln_w':
            a = 1.0 + rng.normal(0.0, 0.1, size=shape) if perturb else np.ones(shape)
        elif kind == 'ln_b':
            a = rng.normal(0.0, 0.05, size=shape) if perturb else np.zeros(shape)
        else:
            raise ValueError(kind)
        sd[name] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    # non-bipartite layers alias attn_prenorm_x_dst to attn_prenorm_x_src (layers.py:52-53): one parameter saved
    # under two keys, so a real checkpoint always holds identical values for the pair
    for stack in NON_BIPARTITE_STACKS:
        n = NUM_LAYERS if stack in ('t_attn_layers', 'a2a_attn_layers') else SEED_LAYERS
        for i in range(n):
            for leaf in ('weight', 'bias'):
                sd[f'{stack}.{i}.attn_prenorm_x_dst.{leaf}'] = sd[f'{stack}.{i}.attn_prenorm_x_src.{leaf}'].clone()
    return sd


# ----------------------------------------------------------------------------------------------------------------
# packing `state_dict()` into the library's blob (layout owned by csrc/engine.cu: build_layout)
# ----------------------------------------------------------------------------------------------------------------
def _np(t) -> np.ndarray:
    return t.detach().cpu().numpy().astype(np.float32) if hasattr(t, 'detach') else np.asarray(t, dtype=np.float32)


def gemm_pack(weight: np.ndarray, n_pad: int = None) -> np.ndarray:
    """nn.Linear weight [N, K] -> K-major blocks [ceil(K/4)][N_pad][4] (zero padded), see csrc/common.cuh block_gemm."""
    n, k = weight.shape
    n_pad = n_pad or n
    k4 = (k + 3) // 4
    wt = np.zeros((k4 * 4, n_pad), dtype=np.float32)
    wt[:k, :n] = weight.T
    return np.ascontiguousarray(wt.reshape(k4, 4, n_pad).transpose(0, 2, 1)).reshape(-1)


def _pad_vec(v: np.ndarray, n_pad: int) -> np.ndarray:
    out = np.zeros(n_pad, dtype=np.float32)
    out[:v.shape[0]] = v
    return out


def _pack_attention(out, sd, p, has_pos):
    g = lambda k: _np(sd[f'{p}.{k}'])
    for dst, src in (('ln_src', 'attn_prenorm_x_src'), ('ln_dst', 'attn_prenorm_x_dst'), ('ln_post', 'attn_postnorm'),
                     ('ln_ffpre', 'ff_prenorm'), ('ln_ffpost', 'ff_postnorm')):
        out[f'{p}.{dst}.g'], out[f'{p}.{dst}.b'] = g(f'{src}.weight'), g(f'{src}.bias')
    out[f'{p}.w_qs'] = gemm_pack(np.concatenate([g('to_q.weight'), g('to_s.weight')]))
    out[f'{p}.b_qs'] = np.concatenate([g('to_q.bias'), g('to_s.bias')])
    out[f'{p}.w_kv'] = gemm_pack(np.concatenate([g('to_k.weight'), g('to_v.weight')]))
    out[f'{p}.b_kv'] = np.concatenate([np.zeros(HIDDEN, np.float32), g('to_v.bias')])      # to_k has no bias
    if has_pos:
        out[f'{p}.w_kr'] = g('to_k_r.weight').reshape(-1)                                  # [out][in] row-major
        out[f'{p}.ln_r.g'], out[f'{p}.ln_r.b'] = g('attn_prenorm_r.weight'), g('attn_prenorm_r.bias')
        out[f'{p}.w_vr'] = gemm_pack(g('to_v_r.weight'))
        out[f'{p}.b_vr'] = g('to_v_r.bias')
    out[f'{p}.w_g'], out[f'{p}.b_g'] = gemm_pack(g('to_g.weight')), g('to_g.bias')
    out[f'{p}.w_out'], out[f'{p}.b_out'] = gemm_pack(g('to_out.weight')), g('to_out.bias')
    out[f'{p}.w_ff1'], out[f'{p}.b_ff1'] = gemm_pack(g('ff_mlp.0.weight')), g('ff_mlp.0.bias')
    out[f'{p}.w_ff2'], out[f'{p}.b_ff2'] = gemm_pack(g('ff_mlp.3.weight')), g('ff_mlp.3.bias')


def _pack_fourier(out, sd, p, dim):
    g = lambda k: _np(sd[f'{p}.{k}'])
    out[f'{p}.freqs'] = g('freqs.weight').reshape(-1)
    for d in range(dim):
        q = f'{p}.mlps.{d}'
        out[f'{q}.w0'], out[f'{q}.b0'] = gemm_pack(g(f'mlps.{d}.0.weight')), g(f'mlps.{d}.0.bias')
        out[f'{q}.ln.g'], out[f'{q}.ln.b'] = g(f'mlps.{d}.1.weight'), g(f'mlps.{d}.1.bias')
        out[f'{q}.w3'], out[f'{q}.b3'] = gemm_pack(g(f'mlps.{d}.3.weight')), g(f'mlps.{d}.3.bias')
    out[f'{p}.out_ln.g'], out[f'{p}.out_ln.b'] = g('to_out.0.weight'), g('to_out.0.bias')
    out[f'{p}.w_out'], out[f'{p}.b_out'] = gemm_pack(g('to_out.2.weight')), g('to_out.2.bias')


def _pack_mlp_embedding(out, sd, p):
    g = lambda k: _np(sd[f'{p}.mlp.{k}'])
    out[f'{p}.w0'], out[f'{p}.b0'] = gemm_pack(g('0.weight')), g('0.bias')
    out[f'{p}.ln1.g'], out[f'{p}.ln1.b'] = g('1.weight'), g('1.bias')
    out[f'{p}.w3'], out[f'{p}.b3'] = gemm_pack(g('3.weight')), g('3.bias')
    out[f'{p}.ln4.g'], out[f'{p}.ln4.b'] = g('4.weight'), g('4.bias')
    out[f'{p}.w6'], out[f'{p}.b6'] = gemm_pack(g('6.weight')), g('6.bias')


def _pack_mlp_layer(out, sd, p):
    g = lambda k: _np(sd[f'{p}.mlp.{k}'])
    n_out = g('3.weight').shape[0]
    n_pad = (n_out + 127) // 128 * 128
    out[f'{p}.w0'], out[f'{p}.b0'] = gemm_pack(g('0.weight')), g('0.bias')
    out[f'{p}.ln.g'], out[f'{p}.ln.b'] = g('1.weight'), g('1.bias')
    out[f'{p}.w3'], out[f'{p}.b3'] = gemm_pack(g('3.weight'), n_pad), _pad_vec(g('3.bias'), n_pad)


def packed_tensors(sd: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
    """Every packed tensor of the library layout, keyed by the library's names."""
    out: Dict[str, np.ndarray] = {}
    for k in ('type_a_emb', 'state_a_emb', 'no_token_emb', 'bos_token_emb', 'invalid_offset_token_emb'):
        out[k] = _np(sd[f'{k}.weight']).reshape(-1)
    _pack_mlp_embedding(out, sd, 'shape_emb')
    for name, dim in (('x_a_emb', 2), ('r_t_emb', 4), ('r_pt2a_emb', 3), ('r_a2a_emb', 3), ('r_pt2sa_emb', 3),
                      ('r_a2sa_emb', 3)):
        _pack_fourier(out, sd, name, dim)
    for name in ('token_emb_veh', 'token_emb_ped', 'token_emb_cyc', 'token_emb_grid', 'fusion_emb'):
        _pack_mlp_embedding(out, sd, name)
    for stack, n, pos in (('t_attn_layers', NUM_LAYERS, True), ('pt2a_attn_layers', NUM_LAYERS, True),
                          ('a2a_attn_layers', NUM_LAYERS, True), ('pt2sa_attn_layers', SEED_LAYERS, True),
                          ('a2sa_attn_layers', SEED_LAYERS, True), ('occ2sa_attn_layers', SEED_LAYERS, False)):
        for i in range(n):
            _pack_attention(out, sd, f'{stack}.{i}', pos)
    for name in ('token_predict_head', 'state_predict_head', 'seed_state_predict_head', 'seed_type_predict_head',
                 'seed_shape_predict_head', 'seed_pos_rel_token_predict_head', 'seed_offset_xy_predict_head',
                 'seed_agent_occ_embed', 'seed_heading_rel_token_predict_head', 'grid_agent_occ_head',
                 'grid_pt_occ_head'):
        _pack_mlp_layer(out, sd, name)
    return out


def packed_map_tensors(map_sd: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
    """Packed tensors of the map encoder (`InfGenMapDecoder.state_dict()` names, map_decoder.py:46-64); library names
    carry the prefix `map.`."""
    sd = {f'map.{k}': v for k, v in map_sd.items()}
    out: Dict[str, np.ndarray] = {}
    for k in ('type_pt_emb', 'polygon_type_emb', 'light_pl_emb'):
        out[f'map.{k}'] = _np(sd[f'map.{k}.weight']).reshape(-1)
    _pack_fourier(out, sd, 'map.r_pt2pt_emb', 3)
    for i in range(3):
        _pack_attention(out, sd, f'map.pt2pt_layers.{i}', True)
    _pack_mlp_layer(out, sd, 'map.token_predict_head')
    _pack_mlp_embedding(out, sd, 'map.token_emb')
    return out


def pack_state_dict(sd: Dict[str, torch.Tensor], lib, map_sd: Dict[str, torch.Tensor] = None) -> np.ndarray:
    """One float32 blob in the layout `lib` (libinfgen_b200.so) reports; raises if a tensor is missing or mis-sized.

    `sd` uses the reference's `InfGenAgentDecoder.state_dict()` names (a full-model checkpoint's
    `encoder.agent_encoder.` prefix is stripped by the caller), `map_sd` the names of `InfGenMapDecoder.state_dict()`
    (`encoder.map_encoder.`).  Either may be None: an engine that only serves the other module keeps zeros there.
    """
    tensors = packed_tensors(sd) if sd is not None else {}
    if map_sd is not None:
        tensors.update(packed_map_tensors(map_sd))
    blob = np.zeros(int(lib.infgen_weight_blob_floats()), dtype=np.float32)
    n = lib.infgen_weight_count()
    for i in range(n):
        name = lib.infgen_weight_name(i).decode()
        if (name.startswith('map.') and map_sd is None) or (not name.startswith('map.') and sd is None):
            continue                               # the module this engine does not serve
        if name not in tensors:
            raise KeyError(f'library expects packed tensor {name!r} that the packer did not produce')
        arr = np.ascontiguousarray(tensors[name], dtype=np.float32).reshape(-1)
        numel, off = int(lib.infgen_weight_numel(name.encode())), int(lib.infgen_weight_offset(name.encode()))
        if arr.size != numel:
            raise ValueError(f'{name}: packed {arr.size} floats, library expects {numel}')
        blob[off:off + numel] = arr
    extra = set(tensors) - {lib.infgen_weight_name(i).decode() for i in range(n)}
    if extra:
        raise KeyError(f'packer produced tensors the library does not know: {sorted(extra)[:5]}')
    return blob
