"""Debug aid: inserted rows of a long-horizon (S = 60) rollout, CUDA path vs oracle insert log; then the heading-stage
edges of the insertion of iteration T_STOP.   python tools/debug_long_insertion.py [T_STOP]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')
import numpy as np
import torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
from infgen_b200.host import prepare_scene, HostBatch
from oracle.agent_decoder_oracle import rollout

T_STOP = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = DecoderConfig(motion_beam_size=1, insert_beam_size=1, disable_insertion=False, debug_force_enter=True,
                    num_recurrent_steps_val=300)
sd = make_state_dict(2)
scene = make_scene(41, num_agents=8, num_map_tokens=256, num_steps=91, ragged=0.3, ego_index=1, cfg=cfg)
o = rollout(scene, sd, cfg, seed=2024, scene_id=0, debug_force_enter=True, collect_trace=True, max_iters=T_STOP + 1)
dec = B200AgentDecoder(sd, cfg, use_cuda_graph=False, seed=2024)
sh = prepare_scene(scene, scene['map_enc'], cfg)
hb = HostBatch([sh], cfg, [0])
dec.load(hb, [sh]); dec.prefill(); dec.step(T_STOP + 1); dec.synchronize()
R, cap, T = hb.R, hb.cap, hb.T
lg = [l for l in o['insert_log'] if l['t'] == T_STOP][-1]
n = o['trace'][T_STOP]['n_rows']
r = n - 1
cur = cfg.hist_cols - 1 + T_STOP
print('rows', n, 'new row', r, 'cur', cur, 'cap', cap)
ha_cnt = dec.debug_read('ha_cnt', (R,), np.int32); ha_src = dec.debug_read('ha_src', (24,), np.int32)
ha_raw = dec.debug_read('ha_raw', (24, 3), np.float32)
hp_cnt = dec.debug_read('hp_cnt', (R,), np.int32); hp_src = dec.debug_read('hp_src', (128,), np.int32)
hp_raw = dec.debug_read('hp_raw', (128, 3), np.float32)
print('agent nbrs gpu', ha_cnt[r], ha_src[:ha_cnt[r]].tolist())
print('agent nbrs ora', len(lg['ha_src']), lg['ha_src'].tolist())
k = min(ha_cnt[r], len(lg['ha_src']))
print('max|ha_raw diff|', np.abs(ha_raw[:k] - lg['ha_raw'].numpy()[:k]).max(axis=0) if k else None)
print('gpu raw', ha_raw[:k].round(4).tolist()); print('ora raw', lg['ha_raw'].numpy()[:k].round(4).tolist())
print('map nbrs gpu', hp_cnt[r], 'ora', len(lg['hp_src']), 'same', np.array_equal(hp_src[:hp_cnt[r]], lg['hp_src'].numpy()))
k = min(hp_cnt[r], len(lg['hp_src']))
print('max|hp_raw diff|', np.abs(hp_raw[:k] - lg['hp_raw'].numpy()[:k]).max(axis=0) if k else None)
pos = dec.debug_read('pos', (R, T, 2), np.float32); head = dec.debug_read('head', (R, T), np.float32)
print('pos gpu', pos[r, cur], 'ora', lg['pos'].numpy(), 'offset ora', lg['offset'].numpy())
xs = dec.debug_read('x_seed', (4, 128), np.float32)[0]
print('query out max diff', np.abs(xs - lg['q'].numpy()).max(), 'as_cnt', dec.debug_read('as_cnt', (1,), np.int32), 'ora', len(lg['as_src']))
ps = dec.debug_read('pred_shape', (R, 3), np.float32)[r]
print('shape gpu', ps, 'ora', lg['shape'].numpy())
x = dec.debug_read('x', (R, 128), np.float32)[r]
print('final feature of the new row max diff', np.abs(x - lg['feat_in'].numpy()).max())
dec.close()
