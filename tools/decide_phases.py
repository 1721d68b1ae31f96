"""Phase breakdown of k_seed_decide (clock64 stamps of CTA 0, last launch of a rollout): python tools/decide_phases.py [graph]"""
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

graph = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
cfg = DecoderConfig(motion_beam_size=5, insert_beam_size=10)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=graph)
scene = make_scene(13, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg)
for rep in range(2):
    dec.inference_batch([scene], [scene['map_enc']])
ts = dec.debug_read('tstamp', (512,), np.int64)
t = ts[384:384 + 32]
t = t[t > 0]
print('k_seed_decide stamps (start -> logits loaded, softmax, per-warp top-k, merge + candidate lookups, decision, '
      'row appended; later entries are stale when the last pass appended nothing):', np.diff(t).tolist(), 'summary at', int(ts[384 + 40] - t[0]))
dec.close()
