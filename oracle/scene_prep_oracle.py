"""CPU restatement of the reference's per-scene preparation of the agent stream.

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() as the checker
of the CUDA path (infgen_b200/csrc/prep.cuh); nothing under infgen_b200/ imports it.

Restated, in plain torch (fp32, CPU), with the reference's operation order:

  tokenize_agent    `TokenProcessor._tokenize_agent`  /root/reference/infgen/datasets/preprocess.py:364-550
                    (clean_heading :315-322, _extrapolate_agent_to_prev_token_step :324-343, _get_agent_shape :345-353,
                    _match_agent_token :552-660, cal_polygon_contour :24-55)
  fetch_enterings   `InfGen._fetch_enterings`          /root/reference/infgen/model/infgen.py:1008-1090
                    (Attr_Tokenizer.encode_pos / encode_heading, infgen/modules/attr_tokenizer.py:77-89, 101-104)

Pinned: tests/test_oracle_prep_vs_golden.py compares both functions with golden vectors written by the UNMODIFIED
reference functions (tests/golden/make_golden_prep.py, build container).
"""
import math
from typing import Dict

import torch

INVALID, VALID, ENTER, EXIT = 0, 1, 2, 3          # configs/ours_standard.yaml:11-15
SHIFT, CURRENT_STEP = 5, 10                       # preprocess.py:277-279


def wrap_angle(a: torch.Tensor) -> torch.Tensor:  # infgen/utils/func.py:58-62
    return -math.pi + (a + math.pi) % (2 * math.pi)


def angle_between(ctr: torch.Tensor, nbr: torch.Tensor) -> torch.Tensor:   # infgen/utils/func.py:30-34
    return torch.atan2(ctr[..., 0] * nbr[..., 1] - ctr[..., 1] * nbr[..., 0], (ctr[..., :2] * nbr[..., :2]).sum(dim=-1))


def box_contour(pos, head, wl):                   # preprocess.py:24-55: corners lf, rf, rb, lb
    x, y = pos[..., 0], pos[..., 1]
    w, l = wl[..., 0], wl[..., 1]
    hc, hs = 0.5 * head.cos(), 0.5 * head.sin()
    lc, ls, wc, ws = l * hc, l * hs, w * hc, w * hs
    return torch.stack([torch.stack((x + lc - ws, y + ls + wc), -1), torch.stack((x + lc + ws, y + ls - wc), -1),
                        torch.stack((x - lc + ws, y - ls - wc), -1), torch.stack((x - lc - ws, y - ls + wc), -1)], dim=-2)


def _rot(theta):                                  # row-vector convention [[c, s], [-s, c]]
    c, s = theta.cos(), theta.sin()
    m = theta.new_zeros(theta.shape[0], 2, 2)
    m[:, 0, 0], m[:, 0, 1], m[:, 1, 0], m[:, 1, 1] = c, s, -s, c
    return m


def tokenize_agent(raw: Dict[str, torch.Tensor], vocab: Dict[str, torch.Tensor], predict_state: bool = True) -> Dict:
    """raw: valid_mask[A,N] bool, heading[A,N], position[A,N,>=2], velocity[A,N,2], type[A] (0 veh, 1 ped, 2 cyc),
    shape[A,N,3]; vocab: {'veh','ped','cyc'} -> [2048,6,4,2].  Returns the token-stream fields the reference writes into
    data['agent'] (:533-546).  Inputs are not modified."""
    valid = raw['valid_mask'].clone()
    heading = raw['heading'].clone().float()
    pos = raw['position'][..., :2].clone().float().contiguous()
    vel = raw['velocity'].clone().float()
    a_type = raw['type'].long()
    A, N = valid.shape
    # clean_heading (:315-322): a jump of more than 1.5 rad between valid neighbours repeats the previous heading
    pairs = valid[:, :-1] & valid[:, 1:]
    for i in range(N - 1):
        jump = torch.abs(wrap_angle(heading[:, i] - heading[:, i + 1])) > 1.5
        ch = jump & pairs[:, i]
        heading[:, i + 1][ch] = heading[:, i][ch]
    wl = torch.tensor([[2.0, 4.8], [1.0, 2.0], [1.0, 1.0]])[a_type]                     # _get_agent_shape (:345-353)
    tables = torch.stack([vocab['veh'], vocab['ped'], vocab['cyc']]).float()
    token_traj = tables[a_type][:, :, -1]                                                # [A,2048,4,2] last sub-step
    # extrapolate to the previous token step (:324-343)
    first = torch.max(valid, dim=1).indices
    for i, t in enumerate(first.tolist()):
        n = t % SHIFT
        if t == CURRENT_STEP and not bool(valid[i, CURRENT_STEP - SHIFT]):
            n = SHIFT
        if n > 0:
            vel[i, t - n:t] = vel[i, t]
            valid[i, t - n:t] = True
            heading[i, t - n:t] = heading[i, t]
            for j in range(n):
                pos[i, t - j - 1] = pos[i, t - j] - vel[i, t] * 0.1
    win = valid.unfold(1, SHIFT + 1, SHIFT)
    token_valid = win[:, :, 0] & win[:, :, -1]                                           # [A,T]
    # _match_agent_token (:552-660): closed-loop nearest-contour matching
    idx_l, contour_l = [], []
    prev_h, prev_p = heading[:, 0], pos[:, 0]
    ar = torch.arange(A)
    for i in range(SHIFT, N, SHIFT):
        ok = valid[:, i - SHIFT] & valid[:, i]
        world = torch.bmm(token_traj.flatten(1, 2), _rot(prev_h)).reshape(*token_traj.shape) + prev_p[:, None, None, :]
        cur = box_contour(pos[:, i], heading[:, i], wl)
        k = torch.argmin(torch.norm(world - cur[:, None], dim=-1).sum(-1), dim=-1)
        con = world[ar, k]
        prev_h = heading[:, i].clone()
        d = con[:, 0] - con[:, 3]
        prev_h[ok] = torch.arctan2(d[:, 1], d[:, 0])[ok]
        prev_p = pos[:, i].clone()
        prev_p[ok] = con.mean(dim=1)[ok]
        idx_l.append(k)
        contour_l.append(con)
    token_index = torch.stack(idx_l, 1)
    token_contour = torch.stack(contour_l, 1)
    token_pos = token_contour.mean(dim=2)
    d = token_contour[:, :, 0] - token_contour[:, :, 3]
    token_heading = torch.arctan2(d[..., 1], d[..., 0])
    T = token_index.shape[1]
    # states (:438-447)
    bos = torch.argmax(token_valid.long(), dim=1)
    eos = T - 1 - torch.argmax(torch.flip(token_valid.long(), dims=[1]), dim=1)
    state = torch.ones_like(token_index)
    step = torch.arange(T)[None].repeat(A, 1)
    state[step == bos[:, None]] = ENTER
    state[step == eos[:, None]] = EXIT
    state[(step < bos[:, None]) | (step > eos[:, None])] = INVALID
    state[state[:, -1] == EXIT, -1] = VALID
    token_valid = token_valid.clone()
    token_valid[state == ENTER] = False
    token_pos[state == INVALID] = 0.
    token_heading[state == INVALID] = 0.
    for i in range(SHIFT, N, SHIFT):
        is_bos = state[:, i // SHIFT - 1] == ENTER
        token_pos[is_bos, i // SHIFT - 1] = pos[is_bos, i].clone()
    token_index[state == INVALID] = -1
    token_index[state == ENTER] = -2
    raw_valid = token_valid.clone()
    if predict_state:
        token_valid = torch.ones_like(token_valid).bool()
    shape = raw['shape'].clone()
    for i in range(A):                                                                    # (:521-524)
        j = int(torch.nonzero(torch.all(shape[i] != 0., dim=-1))[0])
        shape[i, :] = shape[i, j]
    return {'token_idx': token_index, 'state_idx': state, 'token_contour': token_contour, 'token_pos': token_pos,
            'token_heading': token_heading, 'agent_valid_mask': token_valid, 'raw_agent_valid_mask': raw_valid,
            'shape': shape, 'valid_mask': valid, 'heading': heading, 'position_xy': pos}


def _apply_rot(x, theta):                         # attr_tokenizer.py:46-55
    return torch.bmm(x, _rot(theta))


def encode_pos(cells, x, y, theta_y):             # attr_tokenizer.py:77-89
    c = x - y
    c = _apply_rot(c[:, None], -(theta_y - math.pi / 2).expand(x.shape[0]))[:, 0]
    dist = ((c[:, None] - cells[None]) ** 2).sum(-1).sqrt()
    index = torch.argmin(dist, dim=-1)
    return index.long(), c - cells[index]


def encode_heading(h, angle_interval):            # attr_tokenizer.py:101-104
    h = (wrap_angle(h) + torch.pi) / (2 * torch.pi) * 360
    return (h // angle_interval).long()


def fetch_enterings(tok: Dict[str, torch.Tensor], pt_pos: torch.Tensor, av_index: int, cells: torch.Tensor,
                    radius: float = 75.0, angle_interval: float = 3.0) -> Dict[str, torch.Tensor]:
    """tok: token_pos[A,T,2], token_heading[A,T], state_idx[A,T]; pt_pos[P,>=2]; cells[G,2] = Attr_Tokenizer.grid.
    One scene (the reference loops over the graphs of a batch, infgen.py:1021)."""
    pos, head, state = tok['token_pos'], tok['token_heading'], tok['state_idx']
    A, T = state.shape
    ego_p, ego_h = pos[av_index], head[av_index]
    grid = torch.full((A, T), -1, dtype=torch.long)
    off = torch.zeros_like(pos)
    sort_indices = torch.zeros((A, T), dtype=torch.long)
    pt_grid = torch.full((T, pt_pos.shape[0]), -1, dtype=torch.long)
    pos_xy = torch.zeros((A, T, 2))
    bos_l, in_l = [], []
    for t in range(T):
        is_bos = state[:, t] == ENTER
        inv = state[:, t] == INVALID
        inr = ((pos[:, t] - ego_p[[t]]) ** 2).sum(-1).sqrt() <= radius
        m = ~inv & inr
        g, o = encode_pos(cells, pos[m, t], ego_p[[t]], ego_h[[t]])
        grid[m, t] = g
        off[m, t] = o
        pos_xy[m, t] = pos[m, t] - ego_p[[t]]
        hv = torch.stack([ego_h[[t]].cos(), ego_h[[t]].sin()], dim=-1)
        dist = angle_between(hv, pos[:, t] - ego_p[[t]])
        dist[~(is_bos & inr)] = torch.inf
        sd, si = dist.sort()
        si[torch.isinf(sd)] = av_index
        sort_indices[:, t] = si
        bos_l.append(is_bos)
        in_l.append(inr)
        inr_pt = ((pt_pos[:, :2] - ego_p[None, t]) ** 2).sum(-1).sqrt() <= radius
        g, _ = encode_pos(cells, pt_pos[inr_pt, :2], ego_p[[t]], ego_h[[t]])
        pt_grid[t, inr_pt] = g
    rel = head - ego_h[None]
    return {'grid_token_idx': grid, 'grid_offset_xy': off, 'heading_token_idx': encode_heading(rel, angle_interval),
            'pos_xy': pos_xy, 'heading_theta': wrap_angle(rel), 'sort_indices': sort_indices,
            'inrange_mask': torch.stack(in_l, 1), 'bos_mask': torch.stack(bos_l, 1), 'pt_grid_token_idx': pt_grid}


# ---- map side of the preparation ----------------------------------------------------------------------------------------
def match_token_map(traj_pos, traj_theta, pl_idx_list, side, sample_pt) -> Dict[str, torch.Tensor]:
    """`InfGen.match_token_map` /root/reference/infgen/model/infgen.py:918-984 (noise=False): every 5 m map polyline
    (three points, `_tokenize_map` preprocess.py:118-130) is moved into the frame of its first point (:927-935), matched to
    the vocabulary entry with the smallest summed squared distance over the three sample points (:936-937), and the
    [polygon, side, slot] mask of the scene is built from the token counts per polygon and side (:955-971).
    traj_pos [P,3,2] (any float dtype), traj_theta [P], pl_idx_list [P], side [P] uint8, sample_pt [V,3,2]."""
    traj_pos = torch.as_tensor(traj_pos).to(torch.float)
    traj_theta = torch.as_tensor(traj_theta).to(torch.float)
    pl_idx_list = torch.as_tensor(pl_idx_list)
    side = torch.as_tensor(side)
    sample_pt = torch.as_tensor(sample_pt).to(torch.float)
    P = traj_pos.shape[0]
    cos, sin = traj_theta.cos(), traj_theta.sin()
    rot = traj_theta.new_zeros(P, 2, 2)
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1] = cos, -sin, sin, cos
    local = torch.bmm(traj_pos - traj_pos[:, 0:1], rot)
    distance = torch.sum((sample_pt[None] - local.unsqueeze(1)) ** 2, dim=(-2, -1))
    token_idx = torch.argmin(distance, dim=1)
    polygons = pl_idx_list.unique()
    token2pl = torch.stack([torch.arange(P), pl_idx_list.long()])
    counts = torch.stack([torch.stack([((pl_idx_list == pl) & (side == k)).sum() for k in range(3)]) for pl in polygons]).float()
    longest = int(counts.max().item())
    traj_mask = torch.arange(longest)[None, None, :] < counts[:, :, None]
    position = torch.cat([traj_pos[:, 0, :], torch.zeros(P, 1)], dim=-1)
    return {'token_idx': token_idx, 'position': position, 'orientation': traj_theta.clone(), 'height': position[:, -1],
            'traj_mask': traj_mask, 'token2pl': token2pl, 'distance': distance}


def sample_pt_pred(traj_mask: torch.Tensor) -> Dict[str, torch.Tensor]:
    """`InfGen.sample_pt_pred` /root/reference/infgen/model/infgen.py:986-1006: a third of the slots 1.. of every
    (polygon, side) row is drawn without replacement from ONE global permutation (torch.randperm - the same torch RNG
    stream as the reference when seeded alike), masked out of the valid set, and the slot before each masked slot
    predicts it."""
    n_pl, n_side, L = traj_mask.shape
    raw = torch.arange(1, L).repeat(n_pl, n_side, 1)
    k = (L - 1) // 3
    masked = raw.view(-1)[torch.randperm(raw.numel())[:n_pl * n_side * k].reshape(n_pl, n_side, k)]
    masked = torch.sort(masked, -1)[0]
    valid = traj_mask.clone()
    valid.scatter_(2, masked, False)
    pred = traj_mask.clone()
    pred.scatter_(2, masked, False)
    keep = torch.ones_like(pred)
    keep.scatter_(2, masked - 1, False)
    pred.masked_fill_(keep, False)
    pred = pred * torch.roll(traj_mask, shifts=-1, dims=2)
    target = torch.roll(pred, shifts=1, dims=2)
    return {'pt_valid_mask': valid[traj_mask], 'pt_pred_mask': pred[traj_mask], 'pt_target_mask': target[traj_mask]}
