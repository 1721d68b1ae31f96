"""Golden vectors of the per-scene preparation, written by the UNMODIFIED reference
functions on CPU (build container only):

    TokenProcessor._tokenize_agent   /root/reference/infgen/datasets/preprocess.py:364-550 (imported through oracle/shims)
    InfGen._fetch_enterings          /root/reference/infgen/model/infgen.py:1008-1090 - the module itself needs
                                     Lightning / TF, so the function's source is compiled on its own from the file
                                     (ast, unmodified) and called with a stand-in `self` that carries the reference
                                     Attr_Tokenizer and the state constants

    python tests/golden/make_golden_prep.py         # writes tests/golden/case_prep_*.npz

Inputs are regenerated from seeds (infgen_b200.synth.make_scene raw tracks), only the outputs are stored."""
import ast
import os
import sys
import types
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden.cases import PREP_CASES, build_prep_case          # noqa: E402
from oracle import shims                                           # noqa: E402


def reference_fetch_enterings():
    shims.install()
    from torch_geometric.data import HeteroData
    from infgen.utils.func import wrap_angle, angle_between_2d_vectors
    src = open(os.path.join(shims.REFERENCE_ROOT, 'infgen', 'model', 'infgen.py')).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'InfGen')
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == '_fetch_enterings')
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {'torch': torch, 'np': np, 'HeteroData': HeteroData, 'wrap_angle': wrap_angle,
          'angle_between_2d_vectors': angle_between_2d_vectors, 'os': os}
    exec(compile(mod, 'infgen/model/infgen.py:_fetch_enterings', 'exec'), ns)
    return ns['_fetch_enterings']


def main():
    shims.install()
    from torch_geometric.data import HeteroData
    from infgen.datasets.preprocess import TokenProcessor
    from infgen.modules.attr_tokenizer import Attr_Tokenizer
    fetch = reference_fetch_enterings()
    for name in (sys.argv[1:] or list(PREP_CASES)):
        raw, pt_pos, cfg, spec = build_prep_case(name)
        tp = TokenProcessor(token_size=2048, predict_motion=True, predict_state=True, predict_map=False,
                            state_token=dict(cfg.state_token), pl2seed_radius=cfg.pl2seed_radius)
        data = HeteroData()
        data['agent'] = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in raw.items()}
        data = tp._tokenize_agent(data)
        ag = data['agent']
        tok = Attr_Tokenizer(grid_range=cfg.grid_range, grid_interval=cfg.grid_interval, radius=cfg.pl2seed_radius,
                             angle_interval=cfg.angle_interval)
        stub = types.SimpleNamespace(attr_tokenizer=tok, enter_state=int(cfg.state_token['enter']),
                                     invalid_state=int(cfg.state_token['invalid']), pl2seed_radius=cfg.pl2seed_radius,
                                     predict_occ=True)
        A = ag['token_idx'].shape[0]
        P = pt_pos.shape[0]
        data['agent']['av_index'] = raw['av_idx'].clone()
        data['agent']['batch'] = torch.zeros(A, dtype=torch.long)
        data['pt_token'] = {'token_idx': torch.zeros(P, dtype=torch.long), 'position': pt_pos.clone(),
                            'batch': torch.zeros(P, dtype=torch.long)}
        data.num_graphs = 1
        data = fetch(stub, data)
        ag = data['agent']
        save = {k: ag[k].numpy() for k in ('token_idx', 'state_idx', 'token_contour', 'token_pos', 'token_heading',
                                            'agent_valid_mask', 'raw_agent_valid_mask', 'shape', 'grid_token_idx',
                                            'grid_offset_xy', 'heading_token_idx', 'pos_xy', 'heading_theta', 'sort_indices',
                                            'inrange_mask', 'bos_mask', 'pt_grid_token_idx')}
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f'case_prep_{name}.npz')
        np.savez_compressed(path, **save)
        st = ag['state_idx']
        print(f'{name}: A={A} T={st.shape[1]} P={P}; enter {(st == 2).sum().item()} exit {(st == 3).sum().item()} invalid '
              f'{(st == 0).sum().item()}; {os.path.getsize(path) / 1e3:.0f} KB')


if __name__ == '__main__':
    main()
