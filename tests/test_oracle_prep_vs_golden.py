"""Row f2 oracle pin: oracle/scene_prep_oracle.py against golden vectors written by the UNMODIFIED reference functions
(TokenProcessor._tokenize_agent, InfGen._fetch_enterings; tests/golden/make_golden_prep.py)."""
import os
import numpy as np
import pytest
import torch

from tests.golden.cases import PREP_CASES, build_prep_case
from infgen_b200.synth import load_vocab
from infgen_b200.grid import PositionGrid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
INT_KEYS = ('token_idx', 'state_idx', 'agent_valid_mask', 'raw_agent_valid_mask', 'grid_token_idx', 'heading_token_idx',
            'sort_indices', 'inrange_mask', 'bos_mask', 'pt_grid_token_idx')
FLT_KEYS = ('token_contour', 'token_pos', 'token_heading', 'shape', 'grid_offset_xy', 'pos_xy', 'heading_theta')


def run_oracle(name):
    from oracle.scene_prep_oracle import tokenize_agent, fetch_enterings
    raw, pt_pos, cfg, spec = build_prep_case(name)
    tok = tokenize_agent(raw, load_vocab())
    grid = PositionGrid(cfg.grid_range, cfg.grid_interval, cfg.pl2seed_radius, cfg.angle_interval)
    ent = fetch_enterings(tok, pt_pos, spec['ego'], grid.cells, cfg.pl2seed_radius, cfg.angle_interval)
    return {**tok, **ent}


@pytest.mark.parametrize('name', list(PREP_CASES))
def test_prep_oracle_matches_reference(name):
    gold = np.load(os.path.join(GOLD, f'case_prep_{name}.npz'))
    got = run_oracle(name)
    for k in INT_KEYS:                                   # bit-exact: indices, states, masks
        assert np.array_equal(got[k].numpy(), gold[k]), k
    for k in FLT_KEYS:                                   # same torch ops on the same CPU: 1e-6 abs covers BLAS differences
        np.testing.assert_allclose(got[k].numpy(), gold[k], rtol=0, atol=1e-5, err_msg=k)


# ---- map side: match_token_map / sample_pt_pred -----------------------------------------------------------------------------
from tests.golden.cases import MAPMATCH_CASES                      # noqa: E402


def _sample_pt():
    from infgen_b200.map_encoder import load_map_vocab
    traj_src = load_map_vocab()
    return traj_src[:, torch.linspace(0, traj_src.shape[1] - 1, steps=3).long()]       # init_map_token, infgen.py:203-211


@pytest.mark.parametrize('name', list(MAPMATCH_CASES))
def test_map_match_oracle_matches_reference(name):
    from oracle.scene_prep_oracle import match_token_map, sample_pt_pred
    gold = np.load(os.path.join(GOLD, f'case_mapmatch_{name}.npz'))
    got = match_token_map(gold['in_traj_pos'], gold['in_traj_theta'], gold['in_pl_idx_list'], gold['in_side'], _sample_pt())
    for k in ('token_idx', 'traj_mask', 'token2pl'):
        assert np.array_equal(got[k].numpy(), gold[k]), k
    for k in ('position', 'orientation', 'height'):
        assert np.array_equal(got[k].numpy(), gold[k]), k
    torch.manual_seed(MAPMATCH_CASES[name]['mask_seed'])
    masks = sample_pt_pred(got['traj_mask'])
    for k in ('pt_valid_mask', 'pt_pred_mask', 'pt_target_mask'):
        assert np.array_equal(masks[k].numpy(), gold[k]), k
    assert masks['pt_pred_mask'].sum() > 0 and (masks['pt_pred_mask'].sum() == masks['pt_target_mask'].sum())
