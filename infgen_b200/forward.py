"""Host mirror of the MOTION BRANCH of the reference's teacher-forced pass `InfGenAgentDecoder.forward(data, map_enc)`
(/root/reference/infgen/modules/agent_decoder.py:1104-1240), on the GPU through the C ABI
`infgen_load_scenes` + `infgen_forward` of an engine created with `teacher_forced = 1`.

The reference pushes all (agent, column) rows - plus 10 seed rows per graph - through the 18 AttentionLayers at once.  Here
the columns run in order through the closed-loop machinery (embedding of the column from the given token stream, edges
whose destination is that column, K/V of the earlier columns from the temporal ring): the features of a column depend
only on earlier columns, so the result is the same.  The seed rows receive no edge and are the source of none in this branch
(agent_decoder.py:553-556, :635, :713); they are not part of the row space, and `x_a` is returned for the agents only.

Returned keys (same dtypes / shapes as the reference's, :1497-1507): `x_a` [A,T,128], `ego_pos`, `next_token_prob`
[A,T,2048], `next_token_idx` [A,T,10], `next_token_idx_gt`, `next_token_eval_mask`, `next_state_prob` [A,T,3],
`next_state_idx` [A,T,1], `next_state_idx_gt`, `next_state_eval_mask`.  The seed / occupancy / refine branches
(:1232-1386) are not built.  No CPU fallback.
"""
import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _capi
from .config import STATE_TOKEN, TOKEN_SIZE

INVALID, VALID, ENTER, EXIT = (STATE_TOKEN[k] for k in ('invalid', 'valid', 'enter', 'exit'))


def forward_masks(state: torch.Tensor, valid: torch.Tensor, window: int):
    """History (temporal source AND destination) and interaction masks of the teacher-forced pass
    (agent_decoder.py:1142-1161, 545-571)."""
    A, T = state.shape
    is_bos, is_eos = state == ENTER, state == EXIT
    bos = torch.where(is_bos.any(1), is_bos.long().argmax(1), torch.tensor(0))
    eos = torch.where(is_eos.any(1), is_eos.long().argmax(1), torch.tensor(T - 1))
    col = torch.arange(T)[None].expand(A, T)
    hist = torch.ones_like(valid)
    motion = (col > bos[:, None]) & (col <= eos[:, None])
    hist[motion] = valid[motion]
    hist[col < bos[:, None]] = False
    hist[col < torch.clamp(bos - window + 1, min=0)[:, None]] = False
    interact = valid.clone()
    interact[is_bos] = True
    return hist, interact


def eval_masks(state: torch.Tensor, mask: torch.Tensor, av_index: int):
    """next_token_eval_mask / next_state_eval_mask (agent_decoder.py:1388-1417, 1443)."""
    m = mask.clone()
    bos_idx = torch.nonzero(state == ENTER)
    eos_idx = torch.nonzero(state == EXIT)
    T = state.shape[1]
    tok = (m * m.roll(shifts=-1, dims=1) * m.roll(shifts=1, dims=1)).clone()
    for a, c in bos_idx.tolist():
        tok[a, c:c + 1] = 1
        tok[a, c + 1:c + 2] = mask[a, c + 2:c + 3]
    tok[eos_idx[:, 0], eos_idx[:, 1]] = 0
    st = (m * m.roll(shifts=-1, dims=1) * m.roll(shifts=1, dims=1)).clone()
    for a, c in bos_idx.tolist():
        st[a, :c] = 0
        st[a, c:c + 1] = 1
        st[a, c + 1:c + 2] = mask[a, c + 2:c + 3]
    for a, c in eos_idx.tolist():
        st[a, c + 1:] = 1
        st[a, c:c + 1] = mask[a, c - 1:c]
    tok[:, 0] = mask[:, 0] * mask[:, 1]
    st[:, 0] = mask[:, 0] * mask[:, 1]
    tok[:, -1] = 0
    st[:, -1] = 0
    st[av_index] = 0
    return tok.bool(), st.bool()


def teacher_forced_forward(dec, data: Dict, map_enc: Dict) -> Dict:
    ag = data['agent']
    state = ag['state_idx'].long()
    A, T = state.shape
    eng = dec._fwd.get(T)
    if eng is None:
        eng = dec._fwd[T] = type(dec)(dec._state_dict_ref, dec.cfg, device=dec.device, use_cuda_graph=False,
                                      vocab=dec._vocab_ref, teacher_forced_cols=T)
    cfg = dec.cfg
    valid = ag['raw_agent_valid_mask'].bool()
    hist, interact = forward_masks(state, valid, cfg.window)
    av = int(torch.as_tensor(ag['av_index']).reshape(-1)[0])
    cap = (A + 3) // 4 * 4
    P = int(data['pt_token']['position'].shape[0])

    def rows(x, dtype):                                   # [A, ...] -> [cap, ...] contiguous numpy
        x = np.ascontiguousarray(torch.as_tensor(x).cpu().numpy().astype(dtype))
        out = np.zeros((cap,) + x.shape[1:], dtype=dtype)
        out[:A] = x
        return out
    pos_h, head_h = rows(ag['token_pos'].float(), np.float32), rows(ag['token_heading'].float(), np.float32)
    state_h, token_h = rows(state, np.int32), rows(ag['token_idx'], np.int32)
    grid_h = rows(ag['grid_token_idx'], np.int32)
    tsrc_h, int_h = rows(hist, np.uint8), rows(interact, np.uint8)
    type_r = rows(ag['type'], np.int32)
    shape_r = rows(ag['shape'][:, cfg.num_historical_steps - 1].float(), np.float32)
    n_rows, ego, sid = (np.array([v], dtype=np.int32) for v in (A, av, 0))
    pt_ptr = np.array([0, P], dtype=np.int32)
    pt_pos = np.ascontiguousarray(data['pt_token']['position'][:, :2].float().cpu().numpy())
    pt_ori = np.ascontiguousarray(data['pt_token']['orientation'].float().cpu().numpy())
    x_pt = np.ascontiguousarray(map_enc['x_pt'].float().cpu().numpy())
    sb = _capi.SceneBatch(
        n_scenes=1, row_capacity=cap, n_cols=T, n_iters=0, n_rows=_capi.i32p(n_rows), ego_row=_capi.i32p(ego),
        scene_id=_capi.i32p(sid), pos_hist=_capi.f32p(pos_h), head_hist=_capi.f32p(head_h), state_hist=_capi.i32p(state_h),
        token_hist=_capi.i32p(token_h), grid_hist=_capi.i32p(grid_h), tsrc_hist=_capi.u8p(tsrc_h),
        interact_hist=_capi.u8p(int_h), type=_capi.i32p(type_r), shape=_capi.f32p(shape_r), pt_ptr=_capi.i32p(pt_ptr),
        pt_pos=_capi.f32p(pt_pos), pt_ori=_capi.f32p(pt_ori), x_pt=_capi.f32p(x_pt))
    _capi.check(eng.lib.infgen_load_scenes(eng._h, C.byref(sb), _capi.HOST))
    x_a = np.empty((T, cap, 128), dtype=np.float32)
    logits = np.empty((T, cap, TOKEN_SIZE), dtype=np.float32)
    st_logits = np.empty((T, cap, 3), dtype=np.float32)
    _capi.check(eng.lib.infgen_forward(eng._h, _capi.f32p(x_a), _capi.f32p(logits), _capi.f32p(st_logits), _capi.HOST))
    next_token_prob = torch.from_numpy(logits[:, :A]).transpose(0, 1).contiguous()
    next_state_prob = torch.from_numpy(st_logits[:, :A]).transpose(0, 1).contiguous()
    token_idx = ag['token_idx'].long()
    tok_mask, st_mask = eval_masks(state, valid, av)
    state_gt = state.roll(shifts=-1, dims=1).clone()
    state_gt[state_gt == EXIT] = 2                        # valid_state_type.index('exit') (:1451)
    return {
        'x_a': torch.from_numpy(x_a[:, :A]).transpose(0, 1).contiguous(),
        'ego_pos': ag['token_pos'][[av]],
        'next_token_prob': next_token_prob,
        'next_token_idx': torch.topk(torch.softmax(next_token_prob, dim=-1), k=10, dim=-1)[1],
        'next_token_idx_gt': token_idx.roll(shifts=-1, dims=1),
        'next_token_eval_mask': tok_mask,
        'next_state_prob': next_state_prob,
        'next_state_idx': next_state_prob.softmax(dim=-1).argmax(dim=-1, keepdim=True),
        'next_state_idx_gt': state_gt,
        'next_state_eval_mask': st_mask,
    }
