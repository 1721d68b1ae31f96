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


def test_batch_assembly_equals_per_scene_assembly_with_insertion_records():
    """The batch-wide output assembly (tables over the whole row space, one index_put_ per dense insertion tensor, recycled
    storage) returns exactly what the per-scene assembly returns - for two consecutive calls that reuse the pooled
    storage with different records (the first call's records must be gone from the second call's tensors)."""
    from infgen_b200.host import DenseBatchPool, _assemble_outputs_per_scene
    cfg = DecoderConfig(motion_beam_size=1, insert_beam_size=3)
    datas = [make_scene(40 + i, num_agents=6 + i, num_map_tokens=256, num_steps=91, ragged=0.3 * (i % 2), ego_index=i % 3, cfg=cfg)
             for i in range(4)]
    shs = [prepare_scene(d, d['map_enc'], cfg) for d in datas]
    hb = HostBatch(shs, cfg, list(range(4)), pin=False)
    rng = np.random.default_rng(3)
    for name in ('out_pos', 'out_head', 'out_pred_traj', 'out_pred_head', 'out_hist_traj', 'out_hist_head', 'out_rec_pos_prob',
                 'out_rec_agent_occ', 'out_rec_pt_occ', 'out_rec_occ_gt', 'out_rec_state_prob', 'out_pred_shape'):
        t = getattr(hb, name)
        t.copy_(torch.from_numpy(rng.standard_normal(tuple(t.shape)).astype(np.float32)))
    hb.out_pred_state.copy_(torch.from_numpy(rng.integers(0, 4, size=tuple(hb.out_pred_state.shape)).astype(np.float32)))
    hb.out_next_token.copy_(torch.from_numpy(rng.integers(-2, 2048, size=tuple(hb.out_next_token.shape)).astype(np.int32)))
    hb.out_next_state.copy_(torch.from_numpy(rng.integers(0, 4, size=tuple(hb.out_next_state.shape)).astype(np.int32)))
    hb.out_pred_type.copy_(torch.from_numpy(rng.integers(0, 3, size=tuple(hb.out_pred_type.shape)).astype(np.int32)))
    pool = DenseBatchPool(depth=1)                       # depth 1: every call recycles the storage of the call before
    for call in range(3):
        for b, sh in enumerate(shs):
            k = int(rng.integers(0, 12)) if not (call == 1 and b == 2) else 0      # one scene without insertions
            hb.out_n_rows[b] = sh.n_rows + k
            pairs = rng.permutation(10 * (hb.S - 1))[:k]
            meta = np.stack([pairs // 10 + 1, pairs % 10 + 1], -1).astype(np.int32)           # [iteration, slot]
            hb.out_rec_meta[b * hb.cap + sh.n_rows: b * hb.cap + sh.n_rows + k] = torch.from_numpy(meta)
        pool.next_generation()
        got = assemble_outputs(hb, shs, cfg, pool)
        want = _assemble_outputs_per_scene(hb, shs, cfg, None)
        for b, (x, y) in enumerate(zip(got, want)):
            assert set(x) == set(y)
            for k_, v in y.items():
                if isinstance(v, torch.Tensor):
                    assert x[k_].dtype == v.dtype and tuple(x[k_].shape) == tuple(v.shape), (call, b, k_)
                    assert torch.equal(x[k_], v), (call, b, k_)
                else:
                    assert x[k_] == v, (call, b, k_)


def test_engine_group_split_policy():
    """`B200AgentDecoder._split`: contiguous, balanced groups that cover every scene once; one engine for small batches, for
    motion-only engines and for traced runs; never more than `max_engines` groups."""
    from infgen_b200.agent_decoder import B200AgentDecoder
    d = B200AgentDecoder.__new__(B200AgentDecoder)
    d.scenes_per_engine, d.max_engines, d.trace = 2, 4, False
    d.cfg = DecoderConfig(disable_insertion=False)
    assert d._split(1) is None and d._split(3) is None
    assert d._split(4) == [[0, 1], [2, 3]]
    assert [len(g) for g in d._split(8)] == [2, 2, 2, 2]
    assert [len(g) for g in d._split(19)] == [4, 5, 5, 5]
    for n in range(4, 70):
        g = d._split(n)
        assert 2 <= len(g) <= 4 and sum(g, []) == list(range(n)) and max(map(len, g)) - min(map(len, g)) <= 1
    d.scenes_per_engine = 7
    assert d._split(13) is None and [len(g) for g in d._split(19)] == [6, 6, 7]
    d.scenes_per_engine = 0
    assert d._split(64) is None
    d.scenes_per_engine, d.trace = 2, True
    assert d._split(64) is None
    d.trace, d.cfg = False, DecoderConfig(disable_insertion=True)
    assert d._split(64) is None
