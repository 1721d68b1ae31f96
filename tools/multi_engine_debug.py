import os, sys, time
sys.path.insert(0, '/root/repo')
os.environ.setdefault('TQDM_DISABLE', '1')
import torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
from infgen_b200.host import prepare_scene, HostBatch, DeviceBatch
n, K = 32, 4
cfg = DecoderConfig(motion_beam_size=5, insert_beam_size=10)
sd = make_state_dict(0)
datas = [make_scene(13 + i, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg) for i in range(n)]
scenes = [prepare_scene(d, d['map_enc'], cfg) for d in datas]
dev = torch.device('cuda', 0)
decs = [B200AgentDecoder(sd, cfg, seed=2024, scenes_per_engine=0) for _ in range(K)]
groups = [list(range(k * n // K, (k + 1) * n // K)) for k in range(K)]
hbs = [HostBatch([scenes[i] for i in g], cfg, g, row_capacity=224) for g in groups]
hosts = [[scenes[i] for i in g] for g in groups]
def timeit(name, run, sync_each=False):
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); run(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f'{name}: {min(ts)*1e3:.1f} ms', flush=True)
def run_host():
    for d, hb in zip(decs, hbs): d.load(hb); d.rollout()
    for d in decs: d.read()
timeit('host batches, own streams', run_host)
dbs = [DeviceBatch(hb, dev) for hb in hbs]
torch.cuda.synchronize()
def run_dev():
    for d, db, h in zip(decs, dbs, hosts): d.load(db, h); d.rollout(); d.read()
timeit('device batches, own streams', run_dev)
def run_dev2():
    for d, db, h in zip(decs, dbs, hosts): d.load(db, h); d.rollout()
    for d in decs: d.read()
timeit('device batches, own streams, reads last', run_dev2)
streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
for d, st in zip(decs, streams): d.set_stream(st.cuda_stream)
timeit('device batches, torch streams', run_dev)
timeit('host batches, torch streams', run_host)
