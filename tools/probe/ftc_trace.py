"""Debug: clock64 timeline of k_fourier_tc (last launch of a single-scene rollout), built with -DINFGEN_FTC_TRACE:
    nvcc ... -DINFGEN_FTC_TRACE -o tools/probe/libinfgen_trace.so infgen_b200/csrc/engine.cu
Streams per traced block (0: first a2a tile, D = 3; 1: first temporal tile, D = 4): row thread 0, MMA lane, row thread 128."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from infgen_b200 import _capi as _c0; _c0._LIB_PATH = os.path.join(ROOT, 'tools', 'probe', 'libinfgen_trace.so')
os.environ.setdefault('TQDM_DISABLE', '1')
import numpy as np
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
cfg = DecoderConfig(motion_beam_size=5, disable_insertion=True)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=False)
scene = make_scene(13, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg)
dec.inference_batch([scene], [scene['map_enc']])
buf = np.zeros((2, 3, 64), dtype=np.int64)
dec.lib.infgen_debug_ftc_trace(buf.ctypes.data_as(ctypes.c_void_p))
for b in range(2):
    t0 = buf[b][buf[b] > 0].min()
    for s, name in enumerate(['row thread 0  ', 'mma lane      ', 'row thread 128']):
        v = buf[b, s]; v = v[v > 0] - t0
        print(f'block {b} {name}', v.tolist())
        print(f'        deltas       ', np.diff(v).tolist())
dec.close()
