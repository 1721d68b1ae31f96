"""Debug aid: per-iteration, per-row difference between the CUDA path (trace taps) and the oracle for a rollout with a
live insertion stage.   python tools/debug_insertion.py [insert_beam_size]"""
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
from oracle.agent_decoder_oracle import rollout

beam = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cfg = DecoderConfig(motion_beam_size=1, insert_beam_size=beam, disable_insertion=False, debug_force_enter=True,
                    insert_row_reserve=128)
sd = make_state_dict(2)
scene = make_scene(21, num_agents=12, num_map_tokens=512, num_steps=91, ragged=0.3, ego_index=2, cfg=cfg)
o = rollout(scene, sd, cfg, seed=2024, scene_id=0, debug_force_enter=True, collect_trace=True)
dec = B200AgentDecoder(sd, cfg, use_cuda_graph=False, trace=True, seed=2024)
got = dec.inference(scene, scene['map_enc'])
tr = dec.trace_arrays()
for t, w in enumerate(o['trace']):
    n = w['n_rows']
    hin = np.abs(tr['head_in'][t, :n] - w['head_in'].numpy()).max(axis=1)
    lay0 = np.abs(tr['layer_out'][t, 0, :n] - w['layer_out'][0].numpy()).max(axis=1)
    tok_g = tr['token_logits'][t, :n].argmax(-1)
    tok_w = w['token_logits'].argmax(-1).numpy()
    bad = np.nonzero(hin > 1e-3)[0]
    print(f't={t} rows={n} new={w["n_new"]} max|head_in|={hin.max():.2e} max|layer0|={lay0.max():.2e} '
          f'argmax differs at rows {np.nonzero(tok_g != tok_w)[0].tolist()} bad rows {bad.tolist()}')
    for r in bad[:4]:
        print(f'    row {r}: head_in {hin[r]:.3e} layer0 {lay0[r]:.3e}')
dec.close()

# ---- second pass: stop after iteration T_STOP and compare the edge lists / layer outputs of chosen rows
T_STOP = int(os.environ.get('T_STOP', '5'))
from infgen_b200.host import prepare_scene, HostBatch
dec = B200AgentDecoder(sd, cfg, use_cuda_graph=False, trace=True, seed=2024)
sh = prepare_scene(scene, scene['map_enc'], cfg)
hb = HostBatch([sh], cfg, [0])
dec.load(hb, [sh]); dec.prefill(); dec.step(T_STOP + 1)
dec.synchronize()
w = o['trace'][T_STOP]
n = w['n_rows']; R = hb.R; cap = hb.cap
tr = dec.trace_arrays()
print(f'--- iteration {T_STOP}: rows {n}, cap {cap}')
for i in range(6):
    d = np.abs(tr['layer_out'][T_STOP, i, :n] - w['layer_out'][i].numpy()).max(axis=1)
    print(f'layer {i}: bad rows', [(int(r), float(f'{d[r]:.2e}')) for r in np.nonzero(d > 1e-3)[0]])
a_cnt = dec.debug_read('a_cnt', (R,), np.int32); m_cnt = dec.debug_read('m_cnt', (R,), np.int32)
t_cnt = dec.debug_read('t_cnt', (R,), np.int32)
a_src = dec.debug_read('a_src', (R * cap,), np.int32).reshape(R, cap)
a_raw = dec.debug_read('a_raw', (R * cap * 3,), np.float32).reshape(R, cap, 3)
m_raw = dec.debug_read('m_raw', (R * 5 * 3,), np.float32).reshape(R, 5, 3)
ea, em, et = w['edges_a'], w['edges_m'], w['edges_t']
cur = w['cur']
for r in range(max(0, n - 12), n):
    sel = (ea['dst'] == cur * n + r)
    src_w = (ea['src'][sel] - cur * n).numpy(); raw_w = ea['raw'][sel].numpy()
    ok_src = a_cnt[r] == len(src_w) and np.array_equal(a_src[r, :a_cnt[r]], src_w)
    draw = np.abs(a_raw[r, :len(src_w)] - raw_w).max() if ok_src and len(src_w) else -1
    selm = (em['dst'] == cur * n + r)
    drawm = np.abs(m_raw[r, :int(selm.sum())] - em['raw'][selm].numpy()).max() if int(selm.sum()) == m_cnt[r] and m_cnt[r] else -1
    selt = (et['dst'] == r * hb.T + cur)
    print(f'row {r}: a2a cnt {a_cnt[r]} vs {len(src_w)} src_ok {ok_src} max|raw| {draw:.2e} | map cnt {m_cnt[r]} vs {int(selm.sum())} '
          f'max|raw| {drawm:.2e} | temporal cnt {t_cnt[r]} vs {int(selt.sum())}')
dec.close()
