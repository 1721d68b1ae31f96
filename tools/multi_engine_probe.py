"""Does splitting a batch over several engines (own stream + own iteration graph each) overlap their latency-bound insertion
chains?   python tools/multi_engine_probe.py [scenes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')
import torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
from infgen_b200.host import prepare_scene, HostBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = DecoderConfig(motion_beam_size=5, insert_beam_size=10)
sd = make_state_dict(0)
datas = [make_scene(13 + i, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg) for i in range(n)]
scenes = [prepare_scene(d, d['map_enc'], cfg) for d in datas]
for K in (2, 4):
    decs = [B200AgentDecoder(sd, cfg, seed=2024) for _ in range(K)]
    groups = [list(range(k * n // K, (k + 1) * n // K)) for k in range(K)]
    hbs = [HostBatch([scenes[i] for i in g], cfg, g, row_capacity=224) for g in groups]
    def run():
        for d, hb in zip(decs, hbs):
            d.load(hb)
            d.rollout()
        for d in decs:
            d.read()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        run()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    rows = sum(int(hb.out_n_rows.sum()) for hb in hbs)
    print(f'{K} engine(s) x {n // K} scenes: {min(ts) * 1e3:.1f} ms per {n}-scene rollout (load + rollout + read, host buffers); '
          f'rows at the end {rows}', flush=True)
    for d in decs:
        d.close()
