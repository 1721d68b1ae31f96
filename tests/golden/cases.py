"""Seeded definitions of the golden cases (inputs are regenerated, only reference outputs are stored)."""
import torch
from infgen_b200.config import DecoderConfig
from infgen_b200.weights import make_state_dict
from infgen_b200.synth import make_scene

CASES = {
    # BASELINE.json configs[0]: 8 agents, 11 greedy iterations. The reference cannot run a horizon shorter than the
    # scene (agent_decoder.py:1638 only pads), so the scene itself is 66 raw steps (T = 13 columns).
    'cfg0_a8': dict(scene_seed=11, agents=8, map_tokens=512, steps=66, ragged=0.0, ego=0, weight_seed=0,
                    disable_insertion=True, full_logits=True),
    'ragged_a24': dict(scene_seed=12, agents=24, map_tokens=768, steps=91, ragged=0.6, ego=3, weight_seed=1,
                       disable_insertion=False, no_insert_bias=True),
    'std_a64': dict(scene_seed=13, agents=64, map_tokens=2048, steps=91, ragged=0.2, ego=5, weight_seed=0,
                    disable_insertion=True),
    # insertion stage (agent_decoder.py:1744-2114) live: the reference's DEBUG=1 switch (:1888-1889) forces the seed
    # head to 'enter' (random-init weights never insert otherwise); 12 agents grow to 27 rows over 16 iterations
    'insert_a12': dict(scene_seed=21, agents=12, map_tokens=512, steps=91, ragged=0.3, ego=2, weight_seed=2,
                       disable_insertion=False, debug_force_enter=True),
}


def build_case(name: str):
    spec = CASES[name]
    cfg = DecoderConfig(motion_beam_size=1, insert_beam_size=1, disable_insertion=spec['disable_insertion'],
                        debug_force_enter=bool(spec.get('debug_force_enter', False)))
    sd = make_state_dict(spec['weight_seed'])
    if spec.get('no_insert_bias'):
        # make the seed-state head answer 'invalid' for every query: the insertion stage then runs in the reference
        # but inserts nobody, leaving a motion stage whose *state head is live* (agents can turn invalid / exit)
        sd['seed_state_predict_head.mlp.3.bias'] = torch.tensor([30.0, -30.0])
    scene = make_scene(spec['scene_seed'], num_agents=spec['agents'], num_map_tokens=spec['map_tokens'],
                       num_steps=spec['steps'], ragged=spec['ragged'], ego_index=spec['ego'], cfg=cfg)
    return scene, sd, cfg, spec
