"""Phase breakdown of k_layer (clock64 stamps of CTA 0): INFGEN_TSTAMP=1 python tools/layer_phases.py [scenes]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['INFGEN_TSTAMP'] = '1'
os.environ.setdefault('TQDM_DISABLE', '1')
import numpy as np
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = DecoderConfig(motion_beam_size=5, disable_insertion=True)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=False)
scenes = [make_scene(13 + i, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg)
          for i in range(n_scenes)]
for rep in range(2):
    dec.inference_batch(scenes, [s['map_enc'] for s in scenes])
ts = dec.debug_read('tstamp', (512,), np.int64)
for name, off in (('stack / temporal+map', 0), ('agent', 256)):
    t = ts[off:off + 256]
    t = t[t > 0]
    if len(t) < 2:
        continue
    d = np.diff(t)
    print(name, 'total cycles', int(t[-1] - t[0]), 'phases', d.tolist())
dec.close()
