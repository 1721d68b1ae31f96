"""Debug: clock64 timeline of k_node_tc (CTA 0 of the last launch), built with -DINFGEN_NTC_TRACE:
    nvcc ... -DINFGEN_NTC_TRACE -o tools/probe/libinfgen_trace.so infgen_b200/csrc/engine.cu
Streams: row thread 32 (a stamp when a result becomes visible = `take`, and after every published A chunk = `put`),
MMA lane 0 (after the issue of every job).  The traced launch is the full layer (post + pre) of an operator-level
rollout of a few scenes: the last launch that finishes one layer and projects the next with K|V."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from infgen_b200 import _capi as _c0; _c0._LIB_PATH = os.path.join(ROOT, 'tools', 'probe', 'libinfgen_trace.so')
os.environ.setdefault('TQDM_DISABLE', '1')
os.environ['INFGEN_LAYER_PATH'] = 'rows'
import numpy as np, torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.agent_decoder import B200AgentDecoder
from infgen_b200 import ops
cfg = DecoderConfig(motion_beam_size=5, disable_insertion=True)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=False)
from infgen_b200.synth import make_scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scenes = [make_scene(13 + i, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg) for i in range(n)]
dec.inference_batch(scenes, [s_['map_enc'] for s_ in scenes])
buf = np.zeros((2, 64), dtype=np.int64)
dec.lib.infgen_debug_ntc_trace(buf.ctypes.data_as(ctypes.c_void_p))
t0 = buf[buf > 0].min()
for s_, name in enumerate(['row thread 32', 'mma lane 0   ']):
    v = buf[s_]; v = v[v > 0] - t0
    print(name, v.tolist())
    print('   deltas    ', np.diff(v).tolist())
dec.close()
