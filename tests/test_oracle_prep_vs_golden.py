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
