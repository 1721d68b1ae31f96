"""Pin the CPU oracle against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py).

Bar: fp32 outputs within 1e-4 abs / 1e-3 rel of the reference, greedy token and state indices exact.
(In the build container the oracle is bit-identical to the reference; tolerances leave room for a different BLAS.)
"""
import os
import numpy as np
import pytest
import torch

from tests.golden.cases import CASES, build_case
from oracle.agent_decoder_oracle import rollout

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _stack_ragged(xs):
    """Per-iteration tensors whose row count grows (insertion): pad with NaN to the final row count."""
    n = max(x.shape[0] for x in xs)
    return torch.stack([torch.cat([x, x.new_full((n - x.shape[0], *x.shape[1:]), float('nan'))]) for x in xs])


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference_golden(name):
    scene, sd, cfg, spec = build_case(name)
    z = np.load(os.path.join(GOLD, f'case_{name}.npz'))
    inserting = bool(spec.get('debug_force_enter'))
    if inserting:
        assert 'Number of total inserted agents' in str(z['log_message'])
        r = rollout(scene, sd, cfg, collect_trace=True, debug_force_enter=True)
        assert len(r['insert_log']) == z['pos_a'].shape[0] - int(z['n_rows'][0]) > 0
        assert [t['n_rows'] for t in r['trace']] == z['n_rows'].tolist()
    else:
        assert str(z['log_message']) == 'No agents inserted!'
        r = rollout(scene, sd, cfg, collect_trace=True, assume_no_insertion=True)
    out, tr = r['out'], r['trace']
    assert out['ego_index'] == int(z['ego_index'])
    for k in ('next_token_idx', 'next_state_idx', 'agent_id', 'pred_valid', 'valid_mask', 'pred_type'):
        assert np.array_equal(out[k].numpy(), z[k]), k
    for k in ('pos_a', 'head_a', 'pred_traj', 'pred_head', 'pred_state', 'pred_shape', 'eval_shape'):
        np.testing.assert_allclose(out[k].numpy(), z[k], rtol=1e-3, atol=1e-4, err_msg=k)
    head_in = _stack_ragged([t['head_in'] for t in tr]).numpy()
    np.testing.assert_allclose(head_in, z['head_in'], rtol=1e-3, atol=1e-4)
    logits = _stack_ragged([t['token_logits'] for t in tr])
    top_v, top_i = logits.nan_to_num(-1e30).topk(8, dim=-1)
    np.testing.assert_allclose(top_v.numpy(), z['top8_logit'], rtol=1e-3, atol=1e-4)
    assert np.array_equal(top_i[..., 0].numpy(), z['top8_index'][..., 0])
    if 'token_logits' in z:
        np.testing.assert_allclose(logits.numpy(), z['token_logits'], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(_stack_ragged([t['state_logits'] for t in tr]).numpy(), z['state_logits'],
                               rtol=1e-3, atol=1e-4)
    if inserting:
        for k in ('next_state_prob_seed', 'next_pos_rel_prob_seed', 'grid_agent_occ_seed', 'grid_pt_occ_seed',
                  'grid_agent_occ_gt_seed'):
            np.testing.assert_allclose(out[k].numpy(), z[k], rtol=1e-3, atol=1e-4, err_msg=k)
