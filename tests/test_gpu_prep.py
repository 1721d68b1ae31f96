"""Row f2 on the GPU: `infgen_prepare_scene` (k_tokenize_agents + k_fetch_enterings, through the C ABI and the host
mirror `B200ScenePrep`) against the golden vectors of the UNMODIFIED reference functions and against the oracle."""
import os
import numpy as np
import pytest
import torch

from tests.golden.cases import PREP_CASES, build_prep_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
INT_KEYS = ('token_idx', 'state_idx', 'agent_valid_mask', 'raw_agent_valid_mask', 'grid_token_idx', 'heading_token_idx',
            'sort_indices', 'inrange_mask', 'bos_mask', 'pt_grid_token_idx')
FLT_KEYS = ('token_contour', 'token_pos', 'token_heading', 'shape', 'grid_offset_xy', 'pos_xy', 'heading_theta')


@pytest.fixture(scope='module')
def prep():
    from infgen_b200.config import DecoderConfig
    from infgen_b200.weights import make_state_dict
    from infgen_b200.agent_decoder import B200AgentDecoder
    from infgen_b200.scene_prep import B200ScenePrep
    dec = B200AgentDecoder(make_state_dict(0), DecoderConfig(), device=0)
    yield B200ScenePrep(dec)
    dec.close()


def _cost_gap(raw, want, a, c, k_got, k_want):
    """Summed corner distance (preprocess.py:606-608) of two candidate tokens of agent a at token step c, evaluated in
    the frame the REFERENCE had at that step (its matched box of step c - 1): the margin by which the argmin was decided."""
    from oracle.scene_prep_oracle import tokenize_agent, box_contour
    from infgen_b200.synth import load_vocab
    import math
    tok = tokenize_agent(raw, load_vocab())                       # extrapolated / cleaned raw arrays
    pos, heading, valid = tok['position_xy'], tok['heading'], tok['valid_mask']
    ty = int(raw['type'][a])
    voc = load_vocab()[('veh', 'ped', 'cyc')[ty]][:, -1]           # [2048,4,2]
    wl = torch.tensor([[2.0, 4.8], [1.0, 2.0], [1.0, 1.0]])[ty]
    i = 5 * (c + 1)
    if c == 0:
        ph, pp = heading[a, 0], pos[a, 0]
    elif bool(valid[a, i - 10] & valid[a, i - 5]):
        con = torch.as_tensor(want['token_contour'][a, c - 1])
        d = con[0] - con[3]
        ph, pp = torch.atan2(d[1], d[0]), con.mean(0)
    else:
        ph, pp = heading[a, i - 5], pos[a, i - 5]
    cs, sn = math.cos(float(ph)), math.sin(float(ph))
    rot = torch.tensor([[cs, sn], [-sn, cs]])
    cur = box_contour(pos[a, i], heading[a, i], wl)
    cost = lambda k: float(torch.norm(voc[k] @ rot + pp - cur, dim=-1).sum())
    return abs(cost(int(k_got)) - cost(int(k_want)))


def _check_tokens(raw, got, want):
    """Token / state streams: bit-exact, except that the closed-loop match of an agent may leave the reference's track at
    a step whose argmin is decided by less than 1e-4 m (device cosf / sinf / atan2f differ from the host libm in the
    last ulp); from that step on the agent follows a different - equally valid - token sequence and is not compared."""
    g_tok, w_tok = np.asarray(got['token_idx']), np.asarray(want['token_idx'])
    diverged = {}
    for a, c in np.argwhere(g_tok != w_tok):
        if a in diverged:
            continue
        gap = _cost_gap(raw, want, int(a), int(c), g_tok[a, c], w_tok[a, c])
        assert gap < 1e-4, f'agent {a} step {c}: token {g_tok[a, c]} vs {w_tok[a, c]}, cost gap {gap}'
        diverged[int(a)] = int(c)
    assert len(diverged) <= max(1, g_tok.shape[0] // 16), diverged
    return diverged


def _mask_diverged(x, diverged, col_axis=1):
    x = np.array(x, copy=True)
    for a, c in diverged.items():
        x[a, c:] = 0
    return x


def run_gpu(prep, raw, pt_pos):
    data = {'agent': {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in raw.items()},
            'pt_token': {'position': pt_pos.clone()}}
    return prep.tokenize(data)['agent']


@pytest.mark.parametrize('name', list(PREP_CASES))
def test_prepare_scene_matches_reference_golden(prep, name):
    raw, pt_pos, cfg, spec = build_prep_case(name)
    got = run_gpu(prep, raw, pt_pos)
    gold = np.load(os.path.join(GOLD, f'case_prep_{name}.npz'))
    div = _check_tokens(raw, {k: got[k].numpy() for k in ('token_idx',)}, gold)
    for k in INT_KEYS:                                   # bit-exact: tokens, states, cells, heading bins, order, masks
        if k in ('pt_grid_token_idx', 'sort_indices'):
            if not div:
                assert np.array_equal(got[k].numpy(), gold[k]), k
            continue
        assert np.array_equal(_mask_diverged(got[k].numpy().astype(gold[k].dtype), div), _mask_diverged(gold[k], div)), k
    for k in FLT_KEYS:                                   # fp32 positions in metres / radians: 1e-4 rel + 1e-4 abs (north-star: 1e-3 rel)
        np.testing.assert_allclose(_mask_diverged(got[k].numpy(), div), _mask_diverged(gold[k], div), rtol=1e-4, atol=1e-4, err_msg=k)


@pytest.mark.parametrize('seed,agents,ragged', [(41, 5, 0.0), (42, 33, 0.7), (43, 128, 0.4)])
def test_prepare_scene_matches_oracle(prep, seed, agents, ragged):
    """Other sizes (one agent tile, ragged, 128 agents) against the oracle on the same seeded tracks; an argmin decided by
    less than 1e-4 m of summed corner distance may legitimately differ and is reported as such."""
    from oracle.scene_prep_oracle import tokenize_agent, fetch_enterings
    from infgen_b200.synth import load_vocab
    from infgen_b200.grid import PositionGrid
    from tests.golden import cases
    cases.PREP_CASES['_tmp'] = dict(scene_seed=seed, agents=agents, map_tokens=256, ragged=ragged, ego=min(2, agents - 1))
    try:
        raw, pt_pos, cfg, spec = build_prep_case('_tmp')
    finally:
        del cases.PREP_CASES['_tmp']
    got = run_gpu(prep, raw, pt_pos)
    tok = tokenize_agent(raw, load_vocab())
    grid = PositionGrid(cfg.grid_range, cfg.grid_interval, cfg.pl2seed_radius, cfg.angle_interval)
    ent = fetch_enterings(tok, pt_pos, spec['ego'], grid.cells, cfg.pl2seed_radius, cfg.angle_interval)
    want = {**tok, **ent}
    wn = {k: v.numpy() for k, v in want.items()}
    div = _check_tokens(raw, {'token_idx': got['token_idx'].numpy()}, wn)
    for k in INT_KEYS:
        if k in ('pt_grid_token_idx', 'sort_indices'):
            if not div:
                assert np.array_equal(got[k].numpy(), wn[k]), k
            continue
        assert np.array_equal(_mask_diverged(got[k].numpy().astype(wn[k].dtype), div), _mask_diverged(wn[k], div)), k
    for k in FLT_KEYS:
        np.testing.assert_allclose(_mask_diverged(got[k].numpy(), div), _mask_diverged(wn[k], div), rtol=1e-4, atol=1e-4, err_msg=k)


def test_prepare_scene_rejects_bad_arguments(prep):
    from infgen_b200 import _capi
    raw, pt_pos, cfg, spec = build_prep_case('a16')
    raw['av_idx'] = torch.tensor([99])
    with pytest.raises(_capi.InfgenError):
        run_gpu(prep, raw, pt_pos)


# ---- map side: infgen_match_map_tokens / B200ScenePrep.match_token_map ------------------------------------------------------
from tests.golden.cases import MAPMATCH_CASES                      # noqa: E402


def _map_case(name):
    gold = np.load(os.path.join(GOLD, f'case_mapmatch_{name}.npz'))
    data = {'map_save': {'traj_pos': torch.from_numpy(gold['in_traj_pos']), 'traj_theta': torch.from_numpy(gold['in_traj_theta']),
                         'pl_idx_list': torch.from_numpy(gold['in_pl_idx_list'])},
            'pt_token': {'side': torch.from_numpy(gold['in_side']), 'num_nodes': int(gold['in_traj_pos'].shape[0])}}
    return gold, data


@pytest.mark.parametrize('name', list(MAPMATCH_CASES))
def test_map_match_matches_reference_golden(prep, name):
    """Vocabulary match of every 5 m map polyline, [polygon, side, slot] mask, positions / orientations and the
    pt_token -> polygon edges against the unmodified reference method; the prediction masks under the reference's seed."""
    from oracle.scene_prep_oracle import match_token_map
    from tests.test_oracle_prep_vs_golden import _sample_pt
    gold, data = _map_case(name)
    data = prep.match_token_map(data, want_distance=True)
    pt = data['pt_token']
    got, want = pt['token_idx'].numpy(), gold['token_idx']
    # bit-exact, except where the argmin is decided by less than 1e-5 m^2 (device cosf / sinf differ from libm in the last ulp)
    bad = np.nonzero(got != want)[0]
    if len(bad):
        dist = match_token_map(gold['in_traj_pos'], gold['in_traj_theta'], gold['in_pl_idx_list'], gold['in_side'],
                               _sample_pt())['distance'].numpy()
        for t in bad:
            assert abs(dist[t, got[t]] - dist[t, want[t]]) < 1e-5, (t, got[t], want[t])
    assert len(bad) <= max(1, len(want) // 200)
    assert np.array_equal(pt['traj_mask'].numpy(), gold['traj_mask'])
    assert np.array_equal(data[('pt_token', 'to', 'map_polygon')]['edge_index'].numpy(), gold['token2pl'])
    for k in ('position', 'orientation', 'height'):
        assert np.array_equal(pt[k].numpy(), gold[k]), k
    assert pt['token_idx'].dtype == torch.long and pt['traj_mask'].dtype == torch.bool
    torch.manual_seed(MAPMATCH_CASES[name]['mask_seed'])
    data = prep.sample_pt_pred(data)
    for k in ('pt_valid_mask', 'pt_pred_mask', 'pt_target_mask'):
        assert np.array_equal(pt[k].numpy(), gold[k]), k


def test_map_match_full_size_properties(prep):
    """8,192 polylines against the 1,024-entry vocabulary: every vocabulary entry placed anywhere in the plane at any
    heading is matched to itself (distance ~ 0), and a scene's counts add up to its polylines."""
    from tests.test_oracle_prep_vs_golden import _sample_pt
    sp = _sample_pt()
    rng = np.random.default_rng(5)
    P = 8192
    ids = rng.integers(0, sp.shape[0], size=P)
    theta = rng.uniform(-np.pi, np.pi, size=P).astype(np.float32)
    origin = rng.uniform(-150.0, 150.0, size=(P, 1, 2)).astype(np.float32)
    c, s = np.cos(theta)[:, None], np.sin(theta)[:, None]
    loc = sp.numpy()[ids]                                          # [P,3,2] in the token frame; world = loc @ R(theta)^T + o
    world = np.stack([loc[..., 0] * c - loc[..., 1] * s, loc[..., 0] * s + loc[..., 1] * c], -1) + origin
    world = world - world[:, :1] + origin                          # first point at the origin of the frame
    pl = np.sort(rng.integers(0, 300, size=P))
    side = rng.integers(0, 3, size=P).astype(np.uint8)
    data = {'map_save': {'traj_pos': torch.from_numpy(world), 'traj_theta': torch.from_numpy(theta),
                         'pl_idx_list': torch.from_numpy(pl.astype(np.float32))},
            'pt_token': {'side': torch.from_numpy(side), 'num_nodes': P}}
    data = prep.match_token_map(data, want_distance=True)
    pt = data['pt_token']
    d_self = ((sp.numpy()[pt['token_idx'].numpy()] - sp.numpy()[ids]) ** 2).sum((-2, -1))
    assert float(pt['match_distance'].max()) < 1e-3            # found an entry as close as the planted one ...
    assert float(d_self.max()) < 1e-2                          # ... which is the planted one or a duplicate of it
    assert int(pt['traj_mask'].sum()) == P
