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


# --- row f2: per-scene preparation (TokenProcessor._tokenize_agent + InfGen._fetch_enterings) ------------------------------
PREP_CASES = {
    'a16': dict(scene_seed=31, agents=16, map_tokens=512, ragged=0.5, ego=2),
    'a64': dict(scene_seed=13, agents=64, map_tokens=2048, ragged=0.3, ego=5),
}


def build_prep_case(name: str):
    """Raw 10 Hz tracks of a synthetic scene (the inputs of TokenProcessor._tokenize_agent, preprocess.py:364-373) with
    the irregularities real tracks have: first valid steps that are not multiples of 5 (extrapolation, :324-343), a
    heading flip (clean_heading, :315-322), a track too short for any token, gaps."""
    import numpy as np
    spec = PREP_CASES[name]
    cfg = DecoderConfig()
    scene = make_scene(spec['scene_seed'], num_agents=spec['agents'], num_map_tokens=spec['map_tokens'], num_steps=91,
                       ragged=spec['ragged'], ego_index=spec['ego'], cfg=cfg)
    ag = scene['agent']
    rng = np.random.default_rng(1000 + spec['scene_seed'])
    valid = ag['valid_mask'].clone()
    heading = ag['heading'].clone()
    A = valid.shape[0]
    for a in range(A):
        if a == spec['ego']:
            continue
        u = rng.uniform()
        if u < 0.25:                                   # late start at an arbitrary raw step
            valid[a, :int(rng.integers(1, 60))] = False
        elif u < 0.35:                                 # early end
            valid[a, int(rng.integers(20, 88)):] = False
        elif u < 0.40:                                 # a gap in the middle
            g0 = int(rng.integers(15, 60))
            valid[a, g0:g0 + int(rng.integers(3, 14))] = False
        elif u < 0.45:                                 # three valid steps only
            valid[a] = False
            valid[a, 41:44] = True
        if rng.uniform() < 0.15:                       # heading flip by ~pi for a few steps
            h0 = int(rng.integers(5, 80))
            heading[a, h0:h0 + 4] += 3.0
    raw = {'valid_mask': valid, 'heading': heading, 'position': ag['position'].clone(), 'velocity': ag['velocity'].clone(),
           'type': ag['type'].clone(), 'category': torch.zeros(A, dtype=torch.uint8), 'shape': ag['shape'].clone(),
           'av_idx': torch.tensor([spec['ego']], dtype=torch.long)}
    return raw, scene['pt_token']['position'].clone(), cfg, spec


# --- row f2, map side: InfGen.match_token_map / sample_pt_pred over a tokenized synthetic raw map --------------------------------
MAPMATCH_CASES = {
    'p24': dict(seed=1, polygons=24, mask_seed=7),
    'p96': dict(seed=2, polygons=96, mask_seed=8),
}


# --- rows a15 / f4: teacher-forced InfGenAgentDecoder.forward, motion branch -------------------------------------------------
FWD_CASES = {
    'a24': dict(scene_seed=12, agents=24, map_tokens=768, ragged=0.6, ego=3, weight_seed=1),
    'a64': dict(scene_seed=13, agents=64, map_tokens=2048, ragged=0.2, ego=5, weight_seed=0),
}


def build_fwd_case(name: str):
    spec = FWD_CASES[name]
    cfg = DecoderConfig(motion_beam_size=1, insert_beam_size=1, disable_insertion=False)
    sd = make_state_dict(spec['weight_seed'])
    scene = make_scene(spec['scene_seed'], num_agents=spec['agents'], num_map_tokens=spec['map_tokens'], num_steps=91,
                       ragged=spec['ragged'], ego_index=spec['ego'], cfg=cfg)
    return scene, sd, cfg, spec
