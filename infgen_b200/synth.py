"""Synthetic Waymo-shaped scenes in the token layout the decode path consumes.

There is no dataset offline, so scenes are generated: a lane-grid map cut into 5 m map tokens, agents sitting on
lanes around the ego with constant speed / yaw-rate tracks, tokenised to the reference layout
(`infgen/datasets/preprocess.py:364-550`): columns at 2 Hz (column c <-> raw step 5(c+1)), state in
{0 invalid, 1 valid, 2 enter, 3 exit}, motion token id in [0,2048) or -1 (invalid) / -2 (BOS at the enter column),
ego-centric grid token in [0,1961) or -1 (`infgen/model/infgen.py:1008-1075`).

The returned object is a plain nested dict of torch CPU tensors with exactly the keys
`InfGenAgentDecoder.inference` reads (`infgen/modules/agent_decoder.py:1609-1628, 1648-1650, 702-703, 1802`), so
the same scene feeds the reference (under oracle/shims), the CPU oracle and the CUDA path.
All randomness is numpy PCG64 so scenes are bit-identical in every container.
"""
import os
from typing import Dict, Optional
import numpy as np
import torch

from .config import DecoderConfig, SHIFT, TOKEN_SIZE, HIDDEN, STATE_TOKEN
from .grid import PositionGrid

_VOCAB_PATH = os.path.join(os.path.dirname(__file__), 'tokens', 'agent_vocab_555_s2.npz')
_VOCAB_CACHE: Optional[Dict[str, torch.Tensor]] = None
TYPE_NAMES = ('veh', 'ped', 'cyc')


def load_vocab() -> Dict[str, torch.Tensor]:
    """Motion-token vocabulary: {'veh','ped','cyc'} -> float32 [2048, 6, 4, 2] (sub-step, box corner, xy)."""
    global _VOCAB_CACHE
    if _VOCAB_CACHE is None:
        z = np.load(_VOCAB_PATH)
        _VOCAB_CACHE = {k: torch.from_numpy(np.ascontiguousarray(z[k])) for k in TYPE_NAMES}
    return _VOCAB_CACHE


def _make_map(rng: np.random.Generator, num_tokens: int, extent: float = 150.0):
    """Lane grid: straight lanes in x and y plus gentle sinusoidal ones, one token per 5 m. Returns pos[P,3], orient[P]."""
    per_lane = int(2 * extent / 5.0)
    n_lanes = int(np.ceil(num_tokens / per_lane))
    pos, ori = [], []
    offsets = np.linspace(-extent + 8.0, extent - 8.0, (n_lanes + 1) // 2)
    s = np.arange(per_lane) * 5.0 - extent + 2.5
    for i in range(n_lanes):
        off = offsets[i // 2]
        amp = 0.0 if i % 3 else rng.uniform(2.0, 6.0)
        wave = amp * np.sin(s / 40.0 + rng.uniform(0, 6.28))
        dwave = amp / 40.0 * np.cos(s / 40.0)
        if i % 2 == 0:      # runs along +x
            p = np.stack([s, off + wave], -1)
            o = np.arctan2(dwave, np.ones_like(s))
        else:               # runs along +y
            p = np.stack([off + wave, s], -1)
            o = np.arctan2(np.ones_like(s), dwave)
        pos.append(p)
        ori.append(o)
    pos = np.concatenate(pos)[:num_tokens]
    ori = np.concatenate(ori)[:num_tokens]
    pos3 = np.concatenate([pos, np.zeros((pos.shape[0], 1))], -1)
    return pos3.astype(np.float32), ori.astype(np.float32)


def _match_token(vocab_end: np.ndarray, box0: np.ndarray, dx: float, dy: float, dth: float) -> int:
    """Nearest vocabulary token for a local-frame displacement: compare the 4 box corners at the last sub-step."""
    c, s = np.cos(dth), np.sin(dth)
    corners = np.stack([box0[:, 0] * c - box0[:, 1] * s + dx, box0[:, 0] * s + box0[:, 1] * c + dy], -1)
    d = np.linalg.norm(vocab_end - corners[None], axis=-1).sum(-1)
    return int(np.argmin(d))


def make_scene(seed: int, num_agents: int = 64, num_map_tokens: int = 2048, num_steps: int = 91,
               ragged: float = 0.0, ego_index: int = 0, cfg: Optional[DecoderConfig] = None, box: float = 60.0) -> Dict:
    """One synthetic scene.

    ragged: fraction of non-ego agents that enter late / exit early (state tokens enter/exit/invalid in history
    and beyond); 0 gives every agent [enter, valid, valid, ...] which is what a fully observed track tokenises to.
    box: agents start on lanes within +-box metres of the ego (smaller = more crowded).
    """
    cfg = cfg or DecoderConfig()
    rng = np.random.default_rng(seed)
    vocab = load_vocab()
    grid = PositionGrid(cfg.grid_range, cfg.grid_interval, cfg.pl2seed_radius, cfg.angle_interval)
    A, T = num_agents, num_steps // SHIFT
    pt_pos, pt_ori = _make_map(rng, num_map_tokens)
    P = pt_pos.shape[0]

    # --- agents on lanes near the ego -------------------------------------------------------------------
    near = np.nonzero((np.abs(pt_pos[:, 0]) < box) & (np.abs(pt_pos[:, 1]) < box))[0]
    anchor = rng.choice(near, size=A, replace=len(near) < A)
    a_type = rng.choice(3, size=A, p=[0.7, 0.2, 0.1]).astype(np.uint8)
    a_type[ego_index] = 0
    base_shape = np.array([[4.6, 2.0, 1.6], [0.8, 0.8, 1.7], [1.8, 0.7, 1.7]], dtype=np.float32)
    shape = base_shape[a_type] * rng.uniform(0.9, 1.1, size=(A, 1)).astype(np.float32)
    speed = rng.uniform(0.0, 12.0, size=A) * np.where(a_type == 1, 0.15, 1.0)
    yaw_rate = rng.normal(0.0, 0.02, size=A)
    h0 = pt_ori[anchor] + rng.normal(0.0, 0.05, size=A)
    p0 = pt_pos[anchor, :2] + rng.normal(0.0, 0.6, size=(A, 2))
    tt = np.arange(num_steps) * 0.1
    heading = h0[:, None] + yaw_rate[:, None] * tt[None]
    vel = speed[:, None, None] * np.stack([np.cos(heading), np.sin(heading)], -1)
    position = p0[:, None] + np.cumsum(vel * 0.1, axis=1) - vel[:, :1] * 0.1
    # shift everything so the ego is at the origin at the current step (col 1 <-> raw step 10)
    position = position - position[ego_index, 10][None, None]
    pt_pos[:, :2] -= 0.0  # map stays put; only agents are re-centred (the ego need not sit on a lane)
    heading = (heading + np.pi) % (2 * np.pi) - np.pi

    # --- enter / exit columns -----------------------------------------------------------------------------
    enter_col = np.zeros(A, dtype=np.int64)
    exit_col = np.full(A, T, dtype=np.int64)          # T = never exits
    for a in range(A):
        if a == ego_index or rng.uniform() >= ragged:
            continue
        kind = rng.integers(0, 3)
        if kind == 1 and a > ego_index:
            # the reference only supports filtered-out rows *before* the ego (it subtracts just those from
            # batch_size_a, agent_decoder.py:1648-1649), so late-entering agents are placed before it
            kind = 0
        if kind == 0:                                 # enters at the current column
            enter_col[a] = 1
        elif kind == 1:                               # enters later (only visible to the insertion stage / GT)
            enter_col[a] = int(rng.integers(2, T - 2))
        else:                                         # present from the start, exits early
            exit_col[a] = int(rng.integers(1, T - 1))
    col = np.arange(T)[None]
    state = np.full((A, T), STATE_TOKEN['valid'], dtype=np.int64)
    state[col == enter_col[:, None]] = STATE_TOKEN['enter']
    state[col == exit_col[:, None]] = STATE_TOKEN['exit']
    state[(col < enter_col[:, None]) | (col > exit_col[:, None])] = STATE_TOKEN['invalid']

    # --- tokenise ---------------------------------------------------------------------------------------
    token_idx = np.full((A, T), -1, dtype=np.int64)
    token_pos = np.zeros((A, T, 2), dtype=np.float32)
    token_heading = np.zeros((A, T), dtype=np.float32)
    for a in range(A):
        v = vocab[TYPE_NAMES[a_type[a]]].numpy()
        vocab_end, box0 = v[:, -1], v[0, 0]
        for c in range(T):
            if state[a, c] == STATE_TOKEN['invalid']:
                continue
            s1, s0 = SHIFT * (c + 1), SHIFT * c
            token_pos[a, c] = position[a, s1]
            token_heading[a, c] = heading[a, s1]
            if state[a, c] == STATE_TOKEN['enter']:
                token_idx[a, c] = -2
                continue
            d = position[a, s1] - position[a, s0]
            ch, sh = np.cos(heading[a, s0]), np.sin(heading[a, s0])
            token_idx[a, c] = _match_token(vocab_end, box0, d[0] * ch + d[1] * sh, -d[0] * sh + d[1] * ch,
                                           heading[a, s1] - heading[a, s0])
    token_valid = (state == STATE_TOKEN['valid']) | (state == STATE_TOKEN['exit'])
    valid_raw = np.zeros((A, num_steps), dtype=bool)
    for a in range(A):
        lo = SHIFT * (enter_col[a] + 1) if enter_col[a] > 0 else 0
        hi = SHIFT * (exit_col[a] + 1) if exit_col[a] < T else num_steps - 1
        valid_raw[a, lo:hi + 1] = True

    # --- ego-centric grid tokens (history GT; the decode loop recomputes them for generated columns) ---------
    token_pos_t = torch.from_numpy(token_pos)
    token_heading_t = torch.from_numpy(token_heading)
    grid_idx = torch.full((A, T), -1, dtype=torch.long)
    for c in range(T):
        ok = torch.from_numpy(state[:, c] != STATE_TOKEN['invalid'])
        ego_p, ego_h = token_pos_t[[ego_index], c], token_heading_t[[ego_index], c]
        ok &= ((token_pos_t[:, c] - ego_p) ** 2).sum(-1).sqrt() <= cfg.pl2seed_radius
        if bool(ok.any()):
            grid_idx[ok, c] = grid.encode_pos(token_pos_t[ok, c], ego_p, ego_h)

    agent = {
        'num_nodes': A,
        'av_index': torch.tensor([ego_index], dtype=torch.long),
        'id': torch.arange(100, 100 + A, dtype=torch.long),
        'type': torch.from_numpy(a_type),
        'shape': torch.from_numpy(shape)[:, None, :].repeat(1, num_steps, 1).contiguous(),
        'position': torch.from_numpy(np.concatenate([position, np.zeros((A, num_steps, 1))], -1).astype(np.float32)),
        'heading': torch.from_numpy(heading.astype(np.float32)),
        'velocity': torch.from_numpy(vel.astype(np.float32)),
        'valid_mask': torch.from_numpy(valid_raw),
        'token_idx': torch.from_numpy(token_idx),
        'state_idx': torch.from_numpy(state),
        'token_pos': token_pos_t,
        'token_heading': token_heading_t,
        'raw_agent_valid_mask': torch.from_numpy(token_valid),
        'grid_token_idx': grid_idx,
        'trajectory_token_veh': vocab['veh'],
        'trajectory_token_ped': vocab['ped'],
        'trajectory_token_cyc': vocab['cyc'],
    }
    rng_x = np.random.default_rng(seed + 7919)
    scene = {
        'agent': agent,
        'pt_token': {'position': torch.from_numpy(pt_pos), 'orientation': torch.from_numpy(pt_ori), 'num_nodes': P},
        'batch_size_a': torch.tensor([A], dtype=torch.long),
        'scenario_id': [f'synth-{seed}'],
        'num_graphs': 1,
        # stand-in for InfGenMapDecoder's output (map encoder is SURVEY section 8 row f1, outside the loop)
        'map_enc': {'x_pt': torch.from_numpy(rng_x.normal(0.0, 1.0, size=(P, HIDDEN)).astype(np.float32))},
    }
    return scene


def expand_token_traj_all(scene: Dict) -> torch.Tensor:
    """`data['agent']['token_traj_all']` [A,2048,6,4,2] as the reference stores it (preprocess.py:356-362).

    Only the reference / oracle need the per-agent expansion (200 KB per agent); the CUDA path indexes the three
    per-type tables by `type` instead.
    """
    ag = scene['agent']
    tables = torch.stack([ag['trajectory_token_veh'], ag['trajectory_token_ped'], ag['trajectory_token_cyc']])
    return tables[ag['type'].long()].contiguous()


def make_map_tokens(scene, seed: int = 0, n_polygons: int = 64):
    """Synthetic categorical fields of the map tokens of `scene` (schema of `TokenProcessor._tokenize_map`,
    reference preprocess.py:693-761, consumed by `InfGenMapDecoder.forward`, map_decoder.py:70-90): token type (17
    classes), polygon type (4), vocabulary index (1024), the owning polygon and its light type (4), prediction masks.
    numpy PCG64, so the values are bit-identical in every container."""
    import numpy as np
    rng = np.random.default_rng(10_000 + seed)
    P = int(scene['pt_token']['position'].shape[0])
    poly = np.sort(rng.integers(0, n_polygons, size=P))
    pred = rng.random(P) < 0.3
    return {
        'position': scene['pt_token']['position'], 'orientation': scene['pt_token']['orientation'],
        'type': torch.from_numpy(rng.integers(0, 17, size=P).astype(np.uint8)),
        'pl_type': torch.from_numpy(rng.integers(0, 4, size=P).astype(np.uint8)),
        'token_idx': torch.from_numpy(rng.integers(0, 1024, size=P).astype(np.int64)),
        'polygon': torch.from_numpy(poly.astype(np.int64)),
        'polygon_light_type': torch.from_numpy(rng.integers(0, 4, size=n_polygons).astype(np.uint8)),
        'pt_pred_mask': torch.from_numpy(pred), 'pt_valid_mask': torch.ones(P, dtype=torch.bool),
        'pt_target_mask': torch.from_numpy(pred.copy()),
    }


def make_raw_map(seed: int, n_polygons: int = 24, extent: float = 120.0):
    """Synthetic RAW map of a scene, schema of the reference's pre-processed scenario files (`map_point` / `map_polygon` /
    point -> polygon edges): lanes as polygons whose points are sampled every ~1 m along straight,
    curved and kinked centre lines, some with a second point type (road edge) running beside the centre line.  Input of
    `TokenProcessor._tokenize_map` (reference preprocess.py:693-761), whose `map_save` fields `match_token_map` consumes.
    numpy PCG64: bit-identical in every container."""
    rng = np.random.default_rng(20_000 + seed)
    pos, ori, typ, pl = [], [], [], []
    pl_type = []
    for i in range(n_polygons):
        n = int(rng.integers(12, 90))
        s = np.arange(n) * rng.uniform(0.8, 1.3)
        x0, y0 = rng.uniform(-extent, extent, size=2)
        th0 = rng.uniform(-np.pi, np.pi)
        kind = i % 4
        if kind == 0:                                    # straight
            th = np.full(n, th0)
        elif kind == 1:                                  # arc
            th = th0 + s / rng.uniform(25.0, 80.0) * rng.choice([-1.0, 1.0])
        elif kind == 2:                                  # S curve
            th = th0 + 0.4 * np.sin(s / rng.uniform(8.0, 20.0))
        else:                                            # kink (a polyline split, preprocess.py:66-77)
            th = np.where(np.arange(n) < n // 2, th0, th0 + rng.uniform(0.5, 1.2))
        dx, dy = np.cos(th), np.sin(th)
        step = np.diff(s, prepend=0.0)
        px, py = x0 + np.cumsum(dx * step), y0 + np.cumsum(dy * step)
        groups = [(16, 0.0)] if i % 3 else [(16, 0.0), (12, 1.8)]
        for t, lateral in groups:
            pos.append(np.stack([px - lateral * dy, py + lateral * dx, np.zeros(n)], -1))
            ori.append(th)
            typ.append(np.full(n, t))
            pl.append(np.full(n, i))
        pl_type.append(int(rng.integers(0, 4)))
    pos = np.concatenate(pos).astype(np.float32)
    M = pos.shape[0]
    return {
        'map_point': {'num_nodes': M, 'position': torch.from_numpy(pos),
                      'orientation': torch.from_numpy(np.concatenate(ori).astype(np.float32)),
                      'type': torch.from_numpy(np.concatenate(typ).astype(np.uint8))},
        'map_polygon': {'num_nodes': n_polygons, 'type': torch.from_numpy(np.array(pl_type, dtype=np.uint8)),
                        'light_type': torch.full((n_polygons,), 3, dtype=torch.uint8)},
        ('map_point', 'to', 'map_polygon'): {
            'edge_index': torch.stack([torch.arange(M), torch.from_numpy(np.concatenate(pl).astype(np.int64))])},
    }
