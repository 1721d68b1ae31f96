"""Golden vectors of the map side of the per-scene preparation, written by the UNMODIFIED
reference functions on CPU (build container only):

    TokenProcessor._tokenize_map   /root/reference/infgen/datasets/preprocess.py:693-761  (produces the INPUTS, stored too:
                                   the polyline resampling is host-side scipy code that is not part of the row)
    InfGen.match_token_map         /root/reference/infgen/model/infgen.py:918-984
    InfGen.sample_pt_pred          /root/reference/infgen/model/infgen.py:986-1006 (under torch.manual_seed(spec seed))

The two InfGen methods are compiled on their own from the file (ast, unmodified - the module itself needs Lightning / TF)
and called with a stand-in `self` carrying the reference's `map_token` dict (`init_map_token`, :202-211) and noise=False.

    python tests/golden/make_golden_mapmatch.py       # writes tests/golden/case_mapmatch_*.npz"""
import ast
import os
import sys
import types
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden.cases import MAPMATCH_CASES                      # noqa: E402
from infgen_b200.synth import make_raw_map                         # noqa: E402
from infgen_b200.map_encoder import load_map_vocab                 # noqa: E402
from oracle import shims                                           # noqa: E402


def reference_methods(*names):
    src = open(os.path.join(shims.REFERENCE_ROOT, 'infgen', 'model', 'infgen.py')).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'InfGen')
    out = []
    for name in names:
        fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)
        ns = {'torch': torch, 'np': np}
        exec(compile(ast.Module(body=[fn], type_ignores=[]), f'infgen/model/infgen.py:{name}', 'exec'), ns)
        out.append(ns[name])
    return out


def main():
    shims.install()
    from infgen.datasets.preprocess import TokenProcessor
    match, sample = reference_methods('match_token_map', 'sample_pt_pred')
    traj_src = load_map_vocab().numpy()
    idx = torch.linspace(0, traj_src.shape[1] - 1, steps=3).long()          # init_map_token, infgen.py:203-211
    map_token = {'traj_src': torch.from_numpy(traj_src).float(), 'sample_pt': torch.from_numpy(traj_src[:, idx]).float()}
    stub = types.SimpleNamespace(map_token=map_token, noise=False)
    for name in (sys.argv[1:] or list(MAPMATCH_CASES)):
        spec = MAPMATCH_CASES[name]
        data = TokenProcessor._tokenize_map(make_raw_map(spec['seed'], spec['polygons']))
        inputs = {'traj_pos': data['map_save']['traj_pos'].numpy().copy(), 'traj_theta': data['map_save']['traj_theta'].numpy().copy(),
                  'pl_idx_list': data['map_save']['pl_idx_list'].numpy().copy(), 'side': data['pt_token']['side'].numpy().copy()}
        data = match(stub, data)
        torch.manual_seed(spec['mask_seed'])
        data = sample(stub, data)
        pt = data['pt_token']
        save = {f'in_{k}': v for k, v in inputs.items()}
        save.update({k: pt[k].numpy() for k in ('traj_mask', 'position', 'orientation', 'height', 'token_idx', 'pt_valid_mask',
                                                'pt_pred_mask', 'pt_target_mask')})
        save['token2pl'] = data[('pt_token', 'to', 'map_polygon')]['edge_index'].numpy()
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f'case_mapmatch_{name}.npz')
        np.savez_compressed(path, **save)
        print(f"{name}: P={pt['num_nodes']} polygons={pt['traj_mask'].shape[0]} longest={pt['traj_mask'].shape[2]} "
              f"distinct tokens {len(np.unique(save['token_idx']))}; traj_pos {inputs['traj_pos'].dtype} "
              f"{inputs['traj_pos'].shape}; {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == '__main__':
    main()
