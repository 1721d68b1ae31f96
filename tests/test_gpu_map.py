"""Map encoder (`InfGenMapDecoder.forward`, map_decoder.py:70-130) on the GPU, through the C ABI
(`infgen_map_setup` / `infgen_map_encode`), against
  (a) the golden vectors written by the UNMODIFIED reference (tests/golden/make_golden_map.py): a sparse map and a dense one
      whose graph is decided by the max_num_neighbors truncation, and
  (b) the CPU oracle on a batch of fresh maps of different sizes, and
  (c) end to end: map encoder + agent decode on ONE engine with x_pt kept in HBM == agent decode fed the oracle's x_pt.
Bar: x_pt / logits within 1e-3 rel (+1e-4 abs), the pt2pt edge list and the arg-max tokens exact."""
import os
import numpy as np
import pytest
import torch

from tests.golden.make_golden_map import build_map_case, MAP_CASES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
RTOL, ATOL = 1e-3, 1e-4


def _close(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    err = np.abs(a - b)
    bad = err > ATOL + RTOL * np.abs(b)
    assert not bad.any(), f'{what}: {bad.sum()}/{bad.size} out of tolerance, max abs err {err.max():.3e}'


def _data(pt):
    d = {'pt_token': {k: pt[k] for k in ('position', 'orientation', 'type', 'pl_type', 'token_idx', 'pt_pred_mask',
                                         'pt_valid_mask', 'pt_target_mask')}}
    d['pt_token']['light_type'] = pt['polygon_light_type'][pt['polygon']]      # map_decoder.py:85-86
    return d


@pytest.mark.parametrize('name', list(MAP_CASES))
def test_map_encoder_matches_reference_golden(name):
    from infgen_b200.map_encoder import B200MapEncoder
    pt, sd, traj = build_map_case(name)
    z = np.load(os.path.join(GOLD, f'case_{name}.npz'))
    enc = B200MapEncoder.from_state_dict(sd, traj)
    out = enc.forward(_data(pt))
    # the pt2pt graph the kernel built: same edges in the same order (target-major, ascending source)
    P = pt['position'].shape[0]
    cnt = enc._dec.debug_read('map_cnt', (P,), np.int32)
    src = enc._dec.debug_read('map_src', (P, 101), np.int32)
    enc.close()
    want_cnt = np.bincount(z['edge_dst'], minlength=P)
    assert np.array_equal(cnt, want_cnt)
    got_src = np.concatenate([src[p, :cnt[p]] for p in range(P)])
    assert np.array_equal(got_src, z['edge_src'])
    _close(out['x_pt'].numpy(), z['x_pt'], f'{name} x_pt')
    _close(out['map_next_token_prob'].numpy(), z['map_next_token_prob'], f'{name} logits')
    assert np.array_equal(out['map_next_token_idx'].numpy()[:, 0], z['map_next_token_idx'][:, 0])
    assert out['map_next_token_idx'].shape == z['map_next_token_idx'].shape


def test_map_encoder_batch_matches_oracle():
    """Three maps of different sizes (one of full BASELINE size, 2048 tokens) in one call == the oracle per map; tokens of
    different scenes never connect."""
    from oracle.map_decoder_oracle import map_encode
    from infgen_b200.config import DecoderConfig
    from infgen_b200.synth import make_scene, make_map_tokens
    from infgen_b200.weights import make_map_state_dict
    from infgen_b200.map_encoder import B200MapEncoder, load_map_vocab, map_token_fields, encode_on
    cfg = DecoderConfig()
    sd = make_map_state_dict(11)
    traj = load_map_vocab()
    pts = []
    for i, n in enumerate((300, 2048, 640)):
        scene = make_scene(70 + i, num_agents=4, num_map_tokens=n, num_steps=91, ragged=0.0, ego_index=0, cfg=cfg)
        pts.append(make_map_tokens(scene, 20 + i))
    enc = B200MapEncoder.from_state_dict(sd, traj)
    xs, _, ptr = encode_on(enc._dec._h, enc._dec.lib, [map_token_fields(_data(p)) for p in pts], 10.0, True, False)
    enc.close()
    for i, p in enumerate(pts):
        q = dict(p)
        q['light_type'] = p['polygon_light_type'][p['polygon']]
        with torch.no_grad():
            want = map_encode(sd, q, traj)
        _close(xs[ptr[i]:ptr[i + 1]], want['x_pt'].numpy(), f'map {i} x_pt')


def test_map_encoder_feeds_agent_decoder_in_hbm():
    """`InfGenDecoder.inference` (infgen_decoder.py:123-130): map encoder, then the agent decode on its x_pt.  One engine
    holding both weight sets, x_pt never leaving HBM, against the agent decode fed the ORACLE's x_pt."""
    from oracle.map_decoder_oracle import map_encode
    from infgen_b200.config import DecoderConfig
    from infgen_b200.synth import make_scene, make_map_tokens
    from infgen_b200.weights import make_map_state_dict, make_state_dict
    from infgen_b200.agent_decoder import B200AgentDecoder
    from infgen_b200.map_encoder import load_map_vocab
    cfg = DecoderConfig(motion_beam_size=1, disable_insertion=True)
    sd, msd, traj = make_state_dict(4), make_map_state_dict(12), load_map_vocab()
    scene = make_scene(81, num_agents=16, num_map_tokens=512, num_steps=91, ragged=0.2, ego_index=1, cfg=cfg)
    pt = make_map_tokens(scene, 31)
    q = dict(pt)
    q['light_type'] = pt['polygon_light_type'][pt['polygon']]
    with torch.no_grad():
        x_ref = map_encode(msd, q, traj)['x_pt']
    data = dict(scene)
    data['pt_token'] = dict(scene['pt_token'])
    data['pt_token'].update({k: pt[k] for k in ('type', 'pl_type', 'token_idx', 'pt_pred_mask', 'pt_valid_mask', 'pt_target_mask')})
    data['pt_token']['light_type'] = q['light_type']
    dec = B200AgentDecoder(sd, cfg, map_state_dict=msd, map_traj_src=traj)
    fused = dec.inference(data, None)                         # map encoder + decode, x_pt stays on the device
    split = dec.inference(data, {'x_pt': x_ref})              # decode fed the oracle's map encoding
    dec.close()
    assert torch.equal(fused['next_token_idx'], split['next_token_idx'])
    _close(fused['pred_traj'].numpy(), split['pred_traj'].numpy(), 'pred_traj')
