"""Phase breakdown (clock64 stamps of CTA 0, last launch) of the insertion stage's k_layer launches:
    python tools/ins_layer_phases.py <class>      13 = seed query, 14 = new rows, 12 = edge-less K|V chains"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['INFGEN_TSTAMP'] = '1'
os.environ['INFGEN_TSTAMP_CLS'] = sys.argv[1] if len(sys.argv) > 1 else '13'
os.environ.setdefault('TQDM_DISABLE', '1')
import numpy as np
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder

cfg = DecoderConfig(motion_beam_size=5, insert_beam_size=10)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=True)
scene = make_scene(13, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg)
for rep in range(2):
    dec.inference_batch([scene], [scene['map_enc']])
ts = dec.debug_read('tstamp', (512,), np.int64)
t = ts[:256]
t = t[t > 0]
d = np.diff(t)
print('class', os.environ['INFGEN_TSTAMP_CLS'], 'total cycles', int(t[-1] - t[0]))
print('setup, pre0:', d[:2].tolist())
names = ['attn', 'wait', 'agg2', 'gate', 'out', 'ffn_up', 'ffn_down', 'ln', 'pre']
body = d[2:]
for i in range(0, len(body), 9):
    print('  layer', i // 9, dict(zip(names, body[i:i + 9].tolist())))
dec.close()
