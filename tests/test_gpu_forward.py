"""Rows a15 / f4 on the GPU: the motion branch of the teacher-forced `InfGenAgentDecoder.forward` through the C ABI
(`infgen_forward`) and the host mirror `B200AgentDecoder.forward`, against golden vectors of the UNMODIFIED reference method
(tests/golden/make_golden_forward.py) and against the oracle."""
import os
import numpy as np
import pytest
import torch

from tests.golden.cases import FWD_CASES, build_fwd_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RTOL, ATOL = 1e-3, 2e-4          # north-star: 1e-3 rel fp32 (+ an absolute floor for values near zero)


def run_gpu(scene, sd, cfg, **kw):
    from infgen_b200.agent_decoder import B200AgentDecoder
    dec = B200AgentDecoder(sd, cfg, device=0, **kw)
    try:
        return dec.forward(scene, scene['map_enc'])
    finally:
        dec.close()


@pytest.mark.parametrize('name', list(FWD_CASES))
def test_forward_matches_reference_golden(name):
    scene, sd, cfg, spec = build_fwd_case(name)
    got = run_gpu(scene, sd, cfg)
    gold = np.load(os.path.join(GOLD, f'case_fwd_{name}.npz'))
    np.testing.assert_allclose(got['x_a'].numpy(), gold['x_a'], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got['next_state_prob'].numpy(), gold['next_state_prob'], rtol=RTOL, atol=ATOL)
    tv, ti = got['next_token_prob'].topk(8, dim=-1)
    np.testing.assert_allclose(tv.numpy(), gold['top8_logit'], rtol=RTOL, atol=ATOL)
    # greedy token of every (agent, column): exact, unless the reference's own top-2 margin is below the tolerance
    margin = gold['top8_logit'][..., 0] - gold['top8_logit'][..., 1]
    differ = ti[..., 0].numpy() != gold['top8_index'][..., 0]
    assert not np.any(differ & (margin > 2 * ATOL)), int(differ.sum())
    np.testing.assert_allclose(got['next_token_prob'][:, 5].numpy(), gold['logit_col5'], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got['next_token_prob'].sum(-1).numpy(), gold['logit_sum'], rtol=RTOL, atol=5e-3)
    for k in ('next_token_idx_gt', 'next_token_eval_mask', 'next_state_idx_gt', 'next_state_eval_mask'):
        assert np.array_equal(got[k].numpy(), gold[k]), k
    assert np.array_equal(got['next_state_idx'].numpy(), gold['next_state_idx'])
    assert got['next_token_idx'].shape == gold['next_token_idx'].shape and got['next_token_idx'].dtype == torch.long
    assert np.array_equal(got['next_token_idx'][..., 0].numpy()[~differ], gold['next_token_idx'][..., 0][~differ])


@pytest.mark.parametrize('agents,ragged,path', [(7, 0.5, None), (40, 0.4, 'rows'), (130, 0.3, None)])
def test_forward_matches_oracle(monkeypatch, agents, ragged, path):
    """Other row counts (partial tile, the row-tile kernels k_attn + k_node_tc forced, more rows than one wave of
    clusters) against the oracle on the same seeded scene."""
    from oracle.forward_oracle import forward_motion
    from infgen_b200.config import DecoderConfig
    from infgen_b200.weights import make_state_dict
    from infgen_b200.synth import make_scene
    if path:
        monkeypatch.setenv('INFGEN_LAYER_PATH', path)
    cfg = DecoderConfig(motion_beam_size=1, insert_beam_size=1)
    sd = make_state_dict(3)
    scene = make_scene(50 + agents, num_agents=agents, num_map_tokens=640, num_steps=91, ragged=ragged, ego_index=min(4, agents - 1), cfg=cfg)
    got = run_gpu(scene, sd, cfg)
    with torch.no_grad():
        want = forward_motion(scene, sd, cfg)
    for k in ('x_a', 'next_token_prob', 'next_state_prob'):
        torch.testing.assert_close(got[k], want[k], rtol=RTOL, atol=ATOL, msg=lambda m, k=k: f'{k}: {m}')


def test_forward_needs_teacher_forced_engine():
    """infgen_forward on a decode engine is a state error, not a wrong answer."""
    import ctypes as C
    from infgen_b200 import _capi
    from infgen_b200.agent_decoder import B200AgentDecoder
    scene, sd, cfg, spec = build_fwd_case('a24')
    dec = B200AgentDecoder(sd, cfg, device=0)
    try:
        dec.inference(scene, scene['map_enc'])
        rc = dec.lib.infgen_forward(dec._h, None, None, None, _capi.HOST)
        assert rc != 0 and b'teacher_forced' in dec.lib.infgen_last_error()
    finally:
        dec.close()
