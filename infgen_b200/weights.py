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
