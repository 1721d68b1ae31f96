import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests.golden.cases import build_prep_case
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.agent_decoder import B200AgentDecoder
from infgen_b200.scene_prep import B200ScenePrep
dec = B200AgentDecoder(make_state_dict(0), DecoderConfig(), device=0)
prep = B200ScenePrep(dec)
raw, pt_pos, cfg, spec = build_prep_case('a64')
data = {'agent': {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in raw.items()}, 'pt_token': {'position': pt_pos.clone()}}
got = prep.tokenize(data)['agent']
gold = np.load('/root/repo/tests/golden/case_prep_a64.npz')
for k in gold.files:
    g = got[k].numpy(); w = gold[k]
    if g.dtype.kind in 'iub' or w.dtype.kind in 'iub':
        bad = np.argwhere(g.astype(np.int64) != w.astype(np.int64))
        if len(bad): print(k, len(bad), bad[:6].tolist(), [ (g[tuple(b)], w[tuple(b)]) for b in bad[:6]])
    else:
        d = np.abs(g - w); print(k, 'maxdiff', d.max())
a = 0
from infgen_b200.synth import load_vocab
bad = np.argwhere(got['token_idx'].numpy() != gold['token_idx'])
for (a, c) in bad[:4]:
    print('agent', a, 'col', c, 'type', int(raw['type'][a]), 'state', gold['state_idx'][a], 'valid5', raw['valid_mask'][a, ::5].int().tolist())
    print(' gpu tok', got['token_idx'][a].tolist()); print(' ref tok', gold['token_idx'][a].tolist())
