"""Golden vectors of the reference MAP ENCODER (`InfGenMapDecoder.forward`, map_decoder.py:70-130), written by running the
UNMODIFIED reference on CPU through oracle/shims (build container only):

    python tests/golden/make_golden_map.py        # writes tests/golden/case_map_*.npz

Inputs are regenerated from seeds (`infgen_b200.synth.make_scene` / `make_map_tokens`, `weights.make_map_state_dict`), so only
outputs are stored: x_pt [P,128], the token-head logits of the predicted tokens and the pt2pt edge list.
"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')

from infgen_b200.config import DecoderConfig                             # noqa: E402
from infgen_b200.synth import make_scene, make_map_tokens               # noqa: E402
from infgen_b200.weights import make_map_state_dict                     # noqa: E402

MAP_CASES = {
    # sparse lanes: ~3 neighbours per token inside the 10 m radius
    'map_p384': {'scene_seed': 41, 'num_map_tokens': 384, 'weights_seed': 7, 'tokens_seed': 3, 'shrink': 1.0},
    # the same tokens pulled towards the origin (positions x 0.08): most tokens have more than 100 candidates, so the
    # max_num_neighbors truncation and its self-loop quirk decide the graph
    'map_dense_p384': {'scene_seed': 41, 'num_map_tokens': 384, 'weights_seed': 8, 'tokens_seed': 4, 'shrink': 0.08},
}


def build_map_case(name='map_p384'):
    c = MAP_CASES[name]
    cfg = DecoderConfig()
    scene = make_scene(c['scene_seed'], num_agents=8, num_map_tokens=c['num_map_tokens'], num_steps=91,
                       ragged=0.0, ego_index=0, cfg=cfg)
    pt = make_map_tokens(scene, c['tokens_seed'])
    pt['position'] = (pt['position'] * c['shrink']).contiguous()
    sd = make_map_state_dict(c['weights_seed'])
    traj = torch.from_numpy(np.load(os.path.join(ROOT, 'infgen_b200', 'tokens', 'map_traj_token5.npz'))['traj_src'])
    return pt, sd, traj


@torch.no_grad()
def run_reference_map(pt, sd, traj):
    from oracle import shims                       # the reference is only needed (and only present) in the build container
    shims.install()
    from torch_geometric.data import HeteroData
    from infgen.modules.map_decoder import InfGenMapDecoder
    enc = InfGenMapDecoder(dataset='waymo', input_dim=2, hidden_dim=128, num_historical_steps=11, pl2pl_radius=10,
                           num_freq_bands=64, num_layers=3, num_heads=8, head_dim=16, dropout=0.1,
                           map_token={'traj_src': traj})
    enc.load_state_dict(sd, strict=True)
    enc.eval()
    captured = {}
    orig = enc.pt2pt_layers[0].forward

    def spy(x, r, edge_index):
        captured['edge_index'] = edge_index.clone()
        return orig(x, r, edge_index)
    enc.pt2pt_layers[0].forward = spy
    d = HeteroData()
    d['pt_token'] = {k: pt[k] for k in ('position', 'orientation', 'type', 'pl_type', 'token_idx', 'pt_pred_mask',
                                        'pt_valid_mask', 'pt_target_mask')}
    P = pt['position'].shape[0]
    d[('pt_token', 'to', 'map_polygon')] = {'edge_index': torch.stack([torch.arange(P), pt['polygon']])}
    d['map_polygon'] = {'light_type': pt['polygon_light_type']}
    out = enc(d)
    return out, captured['edge_index']


def main():
    torch.manual_seed(0)
    for name in (sys.argv[1:] or list(MAP_CASES)):
        pt, sd, traj = build_map_case(name)
        out, ei = run_reference_map(pt, sd, traj)
        path = os.path.join(os.path.dirname(__file__), f'case_{name}.npz')
        np.savez_compressed(path, x_pt=out['x_pt'].numpy(), map_next_token_prob=out['map_next_token_prob'].numpy(),
                            map_next_token_idx=out['map_next_token_idx'].numpy(), edge_src=ei[0].numpy(),
                            edge_dst=ei[1].numpy())
        print('wrote', path, {k: tuple(v.shape) for k, v in out.items() if isinstance(v, torch.Tensor)}, 'edges',
              ei.shape[1])


if __name__ == '__main__':
    main()
