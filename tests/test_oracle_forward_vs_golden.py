"""Rows a15 / f4 oracle pin: oracle/forward_oracle.py (motion branch of the teacher-forced `forward`) against golden
vectors written by the UNMODIFIED reference method (tests/golden/make_golden_forward.py)."""
import os
import numpy as np
import pytest
import torch

from tests.golden.cases import FWD_CASES, build_fwd_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', list(FWD_CASES))
def test_forward_oracle_matches_reference(name):
    from oracle.forward_oracle import forward_motion
    scene, sd, cfg, spec = build_fwd_case(name)
    with torch.no_grad():
        got = forward_motion(scene, sd, cfg)
    gold = np.load(os.path.join(GOLD, f'case_fwd_{name}.npz'))
    np.testing.assert_allclose(got['x_a'].numpy(), gold['x_a'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got['next_state_prob'].numpy(), gold['next_state_prob'], rtol=1e-5, atol=1e-5)
    tv, ti = got['next_token_prob'].topk(8, dim=-1)
    assert np.array_equal(ti[..., 0].numpy(), gold['top8_index'][..., 0])          # greedy token of every (agent, column)
    np.testing.assert_allclose(tv.numpy(), gold['top8_logit'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got['next_token_prob'][:, 5].numpy(), gold['logit_col5'], rtol=1e-5, atol=1e-5)
