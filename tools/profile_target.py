"""Short workload for ncu: two closed-loop rollouts of the headline scene (64 agents, 2048 map tokens, 16 iterations).
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python tools/profile_target.py [scenes] [graph] [insertion]
The second rollout is the steady-state one (weights L2-resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')
import torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
graph = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
insertion = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
cfg = DecoderConfig(motion_beam_size=5, insert_beam_size=10, disable_insertion=not insertion)
# one engine whatever the batch size: the captures document the kernels at the batch size asked for
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=graph, scenes_per_engine=0)
scenes = [make_scene(13 + i, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg)
          for i in range(n_scenes)]
# the second rollout is bracketed by cudaProfilerStart/Stop: run ncu with `--profile-from-start off` to capture exactly it
for rep in range(2):
    l0 = dec.kernel_launches()
    if rep == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    dec.inference_batch(scenes, [s['map_enc'] for s in scenes])
    if rep == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print(f'rollout {rep}: {dec.kernel_launches() - l0} launches')
dec.close()
