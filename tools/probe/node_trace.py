"""Debug: clock64 phase stamps of k_node (CTA 0, last launch that ran both halves), built with -DINFGEN_NODE_TRACE."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from infgen_b200 import _capi as _c0; _c0._LIB_PATH = os.path.join(ROOT, 'tools', 'probe', 'libinfgen_trace.so')
os.environ.setdefault('TQDM_DISABLE', '1')
os.environ['INFGEN_LAYER_PATH'] = 'rows'
import numpy as np
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = DecoderConfig(motion_beam_size=5, disable_insertion=True)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=False)
scenes = [make_scene(13 + i, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg) for i in range(n)]
dec.inference_batch(scenes, [s['map_enc'] for s in scenes])
buf = np.zeros(32, dtype=np.int64)
dec.lib.infgen_debug_node_trace(buf.ctypes.data_as(ctypes.c_void_p))
names = ['start', 'prologue', 'vr', 'gate', 'out', 'LN', 'ff1', 'ff2', 'x2', 'qs', 'kv', 'qr fold']
v = buf[:12] - buf[0]
print('scenes', n, 'stamps', v.tolist())
print('deltas', {names[i]: int(v[i] - v[i - 1]) for i in range(1, 12)})
dec.close()
