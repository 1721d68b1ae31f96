"""Debug: per-chunk producer-issue / consumer-acquire clocks of k_embed_column (last launch), built with -DINFGEN_WS_TRACE."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from infgen_b200 import _capi as _c0; _c0._LIB_PATH = os.path.join(ROOT, 'tools', 'probe', 'libinfgen_trace.so')
os.environ.setdefault('TQDM_DISABLE', '1')
import numpy as np
from infgen_b200 import _capi
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
cfg = DecoderConfig(motion_beam_size=5, disable_insertion=True)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=False)
scene = make_scene(13, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg)
dec.inference_batch([scene], [scene['map_enc']])
lib = dec.lib
buf = np.zeros((2, 256), dtype=np.int64); n = np.zeros(2, dtype=np.int32)
lib.infgen_debug_ws_trace(buf.ctypes.data_as(ctypes.c_void_p), n.ctypes.data_as(ctypes.c_void_p))
print('counts', n)
t0 = min(buf[0, 0], buf[1, 0])
print('issue  ', (buf[0, :min(n[0], 60)] - t0).tolist())
print('acquire', (buf[1, :min(n[1], 60)] - t0).tolist())
dec.close()
