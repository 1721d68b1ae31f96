"""Host-side logic without a GPU: scene preparation (agent_decoder.py:1609-1657, 1695-1719) and the output dict
(agent_decoder.py:2303-2389).  The CPU oracle rolls a small scene out; its final state is written into the staging buffers
the device would fill, and `assemble_outputs` must reproduce the oracle's output dict key by key (SURVEY.md section 8b:
keys, dtypes, shapes) - which also checks every field `prepare_scene` derives from the inputs."""
import numpy as np
import torch

from infgen_b200.config import DecoderConfig
from infgen_b200.host import prepare_scene, HostBatch, assemble_outputs
from infgen_b200.synth import make_scene
from infgen_b200.weights import make_state_dict


def _rollout(ragged):
    from oracle.agent_decoder_oracle import rollout
    cfg = DecoderConfig(motion_beam_size=1, disable_insertion=True)
    sd = make_state_dict(2)
    scene = make_scene(31, num_agents=9, num_map_tokens=256, num_steps=91, ragged=ragged, ego_index=3, cfg=cfg)
    want = rollout(scene, sd, cfg)['out']
    return cfg, scene, want


def _check(cfg, scene, want):
    sh = prepare_scene(scene, scene['map_enc'], cfg)
    hb = HostBatch([sh], cfg, [0], pin=False)
    n, nh = sh.n_rows, cfg.num_historical_steps
    assert n == want['pos_a'].shape[0]
    # what the device writes (engine.cu:infgen_read)
    hb.out_pos[:n] = want['pos_a']
    hb.out_head[:n] = want['head_a']
    hb.out_pred_traj[:n] = want['pred_traj'][:, nh:]
    hb.out_pred_head[:n] = want['pred_head'][:, nh:]
    hb.out_pred_state[:n] = want['pred_state'][:, nh:]
    hb.out_hist_traj[:n] = want['pred_traj'][:, 1:nh]
    hb.out_hist_head[:n] = want['pred_head'][:, 1:nh]
    ncol = want['next_token_idx'].shape[1]
    hb.out_next_token[:n, :ncol] = want['next_token_idx'].to(torch.int32)
    hb.out_next_state[:n, :ncol] = want['next_state_idx'].to(torch.int32)
    hb.out_n_rows[0] = n
    got = assemble_outputs(hb, [sh], cfg)[0]
    tensor_keys = [k for k, v in want.items() if isinstance(v, torch.Tensor)]
    assert set(tensor_keys) <= set(got), sorted(set(tensor_keys) - set(got))
    # the oracle leaves out the two non-tensor log fields of the reference dict (agent_decoder.py:2387-2388)
    assert set(got) - {'agent_labels', 'log_message'} == set(want), sorted(set(got) ^ set(want))
    for k in tensor_keys:
        g, w = got[k], want[k]
        assert tuple(g.shape) == tuple(w.shape), (k, g.shape, w.shape)
        assert g.dtype == w.dtype, (k, g.dtype, w.dtype)
        if g.dtype.is_floating_point:
            np.testing.assert_allclose(g.numpy(), w.numpy(), rtol=0, atol=0, err_msg=k)
        else:
            assert torch.equal(g, w), k
    assert got['ego_index'] == want['ego_index']
    assert isinstance(got['agent_labels'], list) and isinstance(got['log_message'], str)


def test_output_dict_matches_oracle_all_rows_kept():
    _check(*_rollout(0.0))


def test_output_dict_matches_oracle_ragged_scene():
    """Late-entering / early-exiting tracks: rows invalid at the current step are filtered (agent_decoder.py:1609-1628),
    the ego index shifts, the history masks differ from all-true."""
    _check(*_rollout(0.5))
