"""CPU restatement of the MOTION BRANCH of the reference's teacher-forced pass `InfGenAgentDecoder.forward`
(/root/reference/infgen/modules/agent_decoder.py:1104-1240).

TEST INFRASTRUCTURE ONLY (see oracle/agent_decoder_oracle.py).  Restated:
  * inputs / embedding of every (agent, column) from the ground-truth token stream   :1108-1139 (_agent_token_embedding
    :332-424, training branch)
  * masks                                                                                :1142-1161
  * temporal / agent<->agent / map->agent edges with EVERY column as destination         :1163-1183 (the builders of
    :540-758 in their training form: no inference mask)
  * 6 x {t_attn, pt2a_attn, a2a_attn}                                                     :1204-1216
  * token_predict_head / state_predict_head on every (agent, column)                     :1218-1230
The reference appends 10 seed rows per graph (_pad_feat :511-526).  In the motion branch they receive no edge and are
the source of none (temporal: `hist_mask[-10:] = False` :553-556, agent / map edges: filtered by `pad_mask` :635, :713),
so the agents' features do not depend on them and they are left out here; the golden vectors of the reference (which
carries them) pin that equivalence.
Pinned by tests/test_oracle_forward_vs_golden.py against tests/golden/case_fwd_*.npz (tests/golden/make_golden_forward.py).
"""
from typing import Dict

import torch

from .agent_decoder_oracle import (INVALID, ENTER, EXIT, SEED_TYPE, H, mlp_embedding, mlp_layer, attention_layer,
                                   fourier_embedding, wrap_angle, angle_between, build_vector_a, embed_columns, build_grid,
                                   interaction_edges, map_edges)


def forward_masks(state: torch.Tensor, valid: torch.Tensor):
    """agent_decoder.py:1142-1161 + the source mask of _build_temporal_edge (:545-571).  Returns (hist_mask, interact_mask)."""
    A, T = state.shape
    is_bos, is_eos = state == ENTER, state == EXIT
    bos = torch.where(is_bos.any(1), is_bos.long().argmax(1), torch.tensor(0))
    eos = torch.where(is_eos.any(1), is_eos.long().argmax(1), torch.tensor(T - 1))
    col = torch.arange(T)[None].expand(A, T)
    temporal = torch.ones_like(valid)
    motion = (col > bos[:, None]) & (col <= eos[:, None])
    temporal[motion] = valid[motion]
    interact = valid.clone()
    interact[is_bos] = True
    hist = temporal.clone()
    hist[col < bos[:, None]] = False                                   # temporal_attn_to_invalid = False (:545-550)
    return hist, interact


def forward_motion(scene: Dict, W: Dict[str, torch.Tensor], cfg) -> Dict[str, torch.Tensor]:
    ag = scene['agent']
    nh = cfg.num_historical_steps
    pos, head = ag['token_pos'].clone(), ag['token_heading'].clone()
    token, state = ag['token_idx'].clone(), ag['state_idx'].clone()
    grid_a = ag['grid_token_idx'].clone()
    type_a = ag['type'].long()
    valid = ag['raw_agent_valid_mask'].clone()
    shape_a = ag['shape'][:, nh - 1]
    vocab = torch.stack([ag['trajectory_token_veh'], ag['trajectory_token_ped'], ag['trajectory_token_cyc']])
    pt_pos, pt_ori = scene['pt_token']['position'], scene['pt_token']['orientation']
    x_pt = scene['map_enc']['x_pt']
    A, T = state.shape
    P = pt_pos.shape[0]
    # ---- embedding (:332-424) ------------------------------------------------------------------------------------------
    tok_tab = []
    for ti, nm in enumerate(('veh', 'ped', 'cyc')):
        e = mlp_embedding(W, f'token_emb_{nm}', vocab[ti][:, -1].flatten(1, 2))
        tok_tab.append(torch.cat([e, W['bos_token_emb.weight'], W['no_token_emb.weight']]))
    tok_tab = torch.stack(tok_tab)
    grid = build_grid(cfg.grid_range, cfg.grid_interval, cfg.pl2seed_radius)
    grid_tab = torch.cat([mlp_embedding(W, 'token_emb_grid', grid), W['invalid_offset_token_emb.weight']])
    tok_emb = torch.zeros(A, T, H)
    for ti in range(3):
        m = type_a == ti
        tok_emb[m] = tok_tab[ti][token[m]]
    inv = state == INVALID
    types_at = type_a[:, None].repeat(1, T)
    types_at[inv] = SEED_TYPE
    shapes_at = shape_a[:, None].repeat(1, T, 1)
    shapes_at[inv] = 0.1
    cat = W['type_a_emb.weight'][types_at] + mlp_embedding(W, 'shape_emb', shapes_at.reshape(-1, 3)).view(A, T, H)
    mv, hv = build_vector_a(pos, head, state)
    feat = embed_columns(W, tok_emb, mv, hv, cat, state, grid_tab[grid_a])
    # ---- masks and edges ------------------------------------------------------------------------------------------------
    hist, interact = forward_masks(state, valid)
    is_bos = state == ENTER
    bos = torch.where(is_bos.any(1), is_bos.long().argmax(1), torch.tensor(0))
    col = torch.arange(T)[None].expand(A, T)
    start = torch.clamp(bos - cfg.time_span / cfg.shift + 1, min=0)
    hist = hist.clone()
    hist[~(col >= start[:, None])] = False
    nz = (hist.unsqueeze(2) & hist.unsqueeze(1)).nonzero()            # [agent, src col, dst col] (:575-579)
    src, dst = nz[:, 0] * T + nz[:, 1], nz[:, 0] * T + nz[:, 2]
    keep = (dst - src > 0) & (dst - src <= cfg.time_span / cfg.shift)
    src, dst = src[keep], dst[keep]
    p, h, hvf, invf = pos.reshape(-1, 2), head.reshape(-1), hv.reshape(-1, 2), inv.reshape(-1)
    rp = p[src] - p[dst]
    rh = wrap_angle(h[src] - h[dst])
    rp[invf[src] & ~invf[dst]] = -1.0
    rp[~invf[src] & invf[dst]] = 1.0
    rh[invf[src] & ~invf[dst]] = -1.0
    rp[invf[src] & invf[dst]] = -2.0
    rh[invf[src] & invf[dst]] = -2.0
    raw_t = torch.stack([rp.norm(p=2, dim=-1), angle_between(hvf[dst], rp), rh, (src - dst).float()], dim=-1)
    e_t = (src, dst, raw_t, fourier_embedding(W, 'r_t_emb', raw_t))
    all_cols = torch.ones(A, T, dtype=torch.bool)
    e_a = interaction_edges(W, cfg, pos, head, state, hv, interact, all_cols)
    e_m = map_edges(W, cfg, pos, head, state, hv, interact, all_cols, pt_pos, pt_ori)
    # ---- 6 x {temporal, map, agent} (:1204-1216) ---------------------------------------------------------------------------
    x_pt_tiled = x_pt.repeat(T, 1)                                     # step-major copies (`repeat`, not repeat_interleave)
    for i in range(6):
        x = attention_layer(W, f't_attn_layers.{i}', feat.reshape(-1, H), feat.reshape(-1, H), e_t[3], e_t[0], e_t[1], False)
        x = x.view(A, T, H).transpose(0, 1).reshape(-1, H)
        x = attention_layer(W, f'pt2a_attn_layers.{i}', x_pt_tiled, x, e_m[3], e_m[0], e_m[1], True)
        x = attention_layer(W, f'a2a_attn_layers.{i}', x, x, e_a[3], e_a[0], e_a[1], False)
        feat = x.view(T, A, H).transpose(0, 1)
    return {'x_a': feat, 'next_token_prob': mlp_layer(W, 'token_predict_head', feat),
            'next_state_prob': mlp_layer(W, 'state_predict_head', feat), 'hist_mask': hist, 'interact_mask': interact,
            'n_edges': torch.tensor([e_t[0].numel(), e_m[0].numel(), e_a[0].numel()])}
