"""Short workload for ncu: closed-loop rollouts of the headline scene with the insertion stage enabled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('TQDM_DISABLE', '1')
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene
from infgen_b200.agent_decoder import B200AgentDecoder
force = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
cfg = DecoderConfig(motion_beam_size=5, insert_beam_size=1, disable_insertion=False, debug_force_enter=force)
dec = B200AgentDecoder(make_state_dict(0), cfg, use_cuda_graph=False)
scene = make_scene(13, num_agents=64, num_map_tokens=2048, num_steps=91, ragged=0.0, ego_index=5, cfg=cfg)
for rep in range(2):
    l0 = dec.kernel_launches()
    out = dec.inference_batch([scene], [scene['map_enc']])
    print(f'rollout {rep}: {dec.kernel_launches() - l0} launches, rows {out[0]["pos_a"].shape[0]}')
dec.close()
