"""Convert the reference motion-token vocabulary to a pickle-free .npz (run once, in the build container).

Source: /root/reference/infgen/tokens/agent_vocab_555_s2.pkl  ->  `token_all[{veh,ped,cyc}]` float32 [2048,6,4,2]
(6 sub-steps x 4 box corners x (x,y), agent-local frame; reference `infgen/datasets/preprocess.py:302-311`).
This is the only data file in the reference tree that pins hot-path results; it is
data, not source, and is stored bit-exactly.
"""
import pickle
import sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/infgen/tokens/agent_vocab_555_s2.pkl'
dst = sys.argv[2] if len(sys.argv) > 2 else __file__.replace('convert_vocab.py', 'agent_vocab_555_s2.npz')
tok = pickle.load(open(src, 'rb'))['token_all']
out = {k: np.ascontiguousarray(np.asarray(tok[k], dtype=np.float32)) for k in ('veh', 'ped', 'cyc')}
for k, v in out.items():
    assert v.shape == (2048, 6, 4, 2), (k, v.shape)
np.savez_compressed(dst, **out)
print('wrote', dst, {k: v.shape for k, v in out.items()})
