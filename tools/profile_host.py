"""Host-side profile of the public call on a batch: cProfile over B200AgentDecoder.inference_batch (32 scenes, insertion on).
    python tools/profile_host.py [scenes]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')
import torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = DecoderConfig(motion_beam_size=5, insert_beam_size=10)
dec = B200AgentDecoder(make_state_dict(0), cfg, seed=2024)
scenes = [make_scene(13 + i, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg) for i in range(n)]
maps = [s['map_enc'] for s in scenes]
for _ in range(3):
    dec.inference_batch(scenes, maps, scene_ids=list(range(n)))
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    dec.inference_batch(scenes, maps, scene_ids=list(range(n)))
pr.disable()
print(f'{(time.perf_counter() - t0) / 5 * 1e3:.1f} ms per call')
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
dec.close()
