"""Debug aid: the seed query of pass PASS of iteration T_STOP, CUDA path vs oracle (buffers read after a debug stop).
python tools/debug_seed_query.py [T_STOP] [PASS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')
T_STOP = int(sys.argv[1]) if len(sys.argv) > 1 else 16
PASS = int(sys.argv[2]) if len(sys.argv) > 2 else 0
os.environ['INFGEN_DEBUG_STOP_PASS'] = f'{T_STOP}:{PASS}'
import numpy as np
import torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
from infgen_b200.host import prepare_scene, HostBatch
from oracle.agent_decoder_oracle import rollout

cfg = DecoderConfig(motion_beam_size=1, insert_beam_size=1, disable_insertion=False, debug_force_enter=True,
                    num_recurrent_steps_val=300)
sd = make_state_dict(2)
scene = make_scene(41, num_agents=8, num_map_tokens=256, num_steps=91, ragged=0.3, ego_index=1, cfg=cfg)
o = rollout(scene, sd, cfg, seed=2024, scene_id=0, debug_force_enter=True, collect_trace=True, max_iters=T_STOP + 1)
pl = [p for p in o['pass_log'] if p['t'] == T_STOP and p['pass'] == PASS][0]
dec = B200AgentDecoder(sd, cfg, use_cuda_graph=False, seed=2024)
sh = prepare_scene(scene, scene['map_enc'], cfg)
hb = HostBatch([sh], cfg, [0])
dec.load(hb, [sh]); dec.prefill()
try:
    dec.step(T_STOP + 1)
except RuntimeError as ex:
    print('stopped:', ex)
R, cap, T = hb.R, hb.cap, hb.T
G = 1961
n = pl['x_sa_in'].shape[0]
print('rows', n)
xs = dec.debug_read('x_seed', (4, 128), np.float32)[0]
print('query out max diff', np.abs(xs - pl['q'].numpy()).max())
occ = dec.debug_read('occ', (G,), np.float32)
print('occ equal', np.array_equal(occ, pl['occ'].numpy().astype(np.float32)), 'cells', np.nonzero(occ)[0].tolist(), np.nonzero(pl['occ'].numpy())[0].tolist())
print('occ_emb max diff', np.abs(dec.debug_read('occ_emb', (128,), np.float32) - pl['occ_emb'].numpy()).max())
as_cnt = int(dec.debug_read('as_cnt', (1,), np.int32)[0]); stride = min(cap, 300)
as_src = dec.debug_read('as_src', (stride,), np.int32)[:as_cnt]; as_raw = dec.debug_read('as_raw', (stride, 3), np.float32)[:as_cnt]
print('agent->seed gpu', as_src.tolist(), 'ora', pl['as_src'].tolist())
if as_cnt == len(pl['as_src']) and as_cnt:
    print('  raw max diff', np.abs(as_raw - pl['as_raw'].numpy()).max(axis=0))
ps_cnt = int(dec.debug_read('ps_cnt', (1,), np.int32)[0])
ps_src = dec.debug_read('ps_src', (2048,), np.int32)[:ps_cnt]; ps_raw = dec.debug_read('ps_raw', (2048, 3), np.float32)[:ps_cnt]
print('map->seed gpu', ps_cnt, 'ora', len(pl['ps_src']), 'same src', np.array_equal(ps_src, pl['ps_src'].numpy()))
if ps_cnt == len(pl['ps_src']):
    d = np.abs(ps_raw - pl['ps_raw'].numpy())
    print('  raw max diff', d.max(axis=0), 'rows with diff > 1e-3', np.nonzero(d.max(axis=1) > 1e-3)[0].tolist())
    for i in np.nonzero(d.max(axis=1) > 1e-3)[0][:5]:
        print('   ', i, ps_raw[i], pl['ps_raw'].numpy()[i])
x_sa = dec.debug_read('x', (R, 128), np.float32)[:n]
print('agent features in max diff per row', np.abs(x_sa - pl['x_sa_in'].numpy()).max(axis=1).round(5).tolist())
dec.close()
