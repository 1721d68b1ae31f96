"""Golden vectors of the motion branch of the teacher-forced pass, written by the UNMODIFIED reference
`InfGenAgentDecoder.forward` (/root/reference/infgen/modules/agent_decoder.py:1104-1240) on CPU (build container only).

    python tests/golden/make_golden_forward.py         # writes tests/golden/case_fwd_*.npz

The reference method is called as is and runs to its end (seed and refine branches included); the keys of the motion
branch are taken from its return value (`x_a`, `next_token_prob`, `next_token_idx`, `next_token_idx_gt`,
`next_token_eval_mask`, `next_state_prob`, `next_state_idx`, `next_state_idx_gt`, `next_state_eval_mask`).  The extra
fields `forward` reads at its top (grid offsets, heading tokens, entering order; `InfGen._fetch_enterings`) come from
oracle/scene_prep_oracle.py, which is pinned against the reference's own `_fetch_enterings`."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')
from tests.golden.cases import FWD_CASES, build_fwd_case                      # noqa: E402
from oracle.ref_runner import build_reference_decoder, to_hetero               # noqa: E402
from oracle.scene_prep_oracle import fetch_enterings                           # noqa: E402
from infgen_b200.grid import PositionGrid                                      # noqa: E402


MOTION_KEYS = ('next_token_idx', 'next_token_idx_gt', 'next_token_eval_mask', 'next_state_prob', 'next_state_idx',
               'next_state_idx_gt', 'next_state_eval_mask')


def reference_forward(scene, sd, cfg):
    dec = build_reference_decoder(sd, cfg)
    data = to_hetero(scene)
    ag = data['agent']
    grid = PositionGrid(cfg.grid_range, cfg.grid_interval, cfg.pl2seed_radius, cfg.angle_interval)
    ent = fetch_enterings({k: ag[k] for k in ('token_pos', 'token_heading', 'state_idx')}, data['pt_token']['position'],
                          int(ag['av_index'][0]), grid.cells, cfg.pl2seed_radius, cfg.angle_interval)
    assert torch.equal(ent['grid_token_idx'], ag['grid_token_idx'])
    for k in ('grid_offset_xy', 'heading_token_idx', 'pos_xy', 'heading_theta', 'sort_indices', 'pt_grid_token_idx'):
        ag[k] = ent[k]
    A, P = ag['token_idx'].shape[0], data['pt_token']['position'].shape[0]
    av = int(ag['av_index'][0])
    ag['batch'], ag['ptr'] = torch.zeros(A, dtype=torch.long), torch.tensor([0, A])
    data['pt_token']['batch'], data['pt_token']['ptr'] = torch.zeros(P, dtype=torch.long), torch.tensor([0, P])
    data['ego_pos'], data['ego_heading'] = ag['token_pos'][[av]], ag['token_heading'][[av]]
    data.num_graphs = 1
    torch.manual_seed(0)                                          # (the seed branch draws random evaluation masks)
    with torch.no_grad():
        out = dec.forward(data, {'x_pt': scene['map_enc']['x_pt'].clone()})
    res = {k: out[k] for k in MOTION_KEYS}
    res['x_a'] = out['x_a'][:A]                                   # rows [A, A + 10) are the seed rows (_pad_feat)
    res['next_token_prob'] = out['next_token_prob']
    return res


def main():
    for name in (sys.argv[1:] or list(FWD_CASES)):
        scene, sd, cfg, spec = build_fwd_case(name)
        out = reference_forward(scene, sd, cfg)
        top_v, top_i = out['next_token_prob'].topk(8, dim=-1)
        save = {'x_a': out['x_a'].numpy(), 'next_state_prob': out['next_state_prob'].numpy(), 'top8_logit': top_v.numpy(),
                'top8_index': top_i.numpy(), 'logit_sum': out['next_token_prob'].sum(-1).numpy(),
                'logit_col5': out['next_token_prob'][:, 5].numpy()}
        save.update({k: out[k].numpy() for k in MOTION_KEYS if k != 'next_state_prob'})
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f'case_fwd_{name}.npz')
        np.savez_compressed(path, **save)
        print(f'{name}: x_a {tuple(out["x_a"].shape)}, {os.path.getsize(path) / 1e6:.2f} MB')


if __name__ == '__main__':
    main()
