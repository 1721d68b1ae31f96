"""Run the UNMODIFIED reference `InfGenAgentDecoder.inference` on CPU (build container only).

TEST INFRASTRUCTURE ONLY - used by tests/golden/make_golden.py to generate golden vectors and by
tests/test_oracle_vs_reference.py (skipped where /root/reference is absent, i.e. on the GPU box).
The reference is imported from /root/reference through `oracle.shims`; nothing is copied.
"""
import os
from typing import Dict, List, Optional
import torch

from . import shims


def reference_available() -> bool:
    return os.path.isdir(os.path.join(shims.REFERENCE_ROOT, 'infgen', 'modules'))


def build_reference_decoder(state_dict: Dict[str, torch.Tensor], cfg):
    """Instantiate the reference decoder with the constructor arguments `InfGen.__init__` passes
    (infgen/model/infgen.py:86-134, configs/ours_standard.yaml) and load `state_dict` strictly."""
    shims.install()
    from infgen.modules.attr_tokenizer import Attr_Tokenizer
    from infgen.modules.agent_decoder import InfGenAgentDecoder
    tok = Attr_Tokenizer(grid_range=cfg.grid_range, grid_interval=cfg.grid_interval, radius=cfg.pl2seed_radius,
                         angle_interval=cfg.angle_interval)
    dec = InfGenAgentDecoder(
        dataset='waymo', input_dim=2, hidden_dim=128, num_historical_steps=cfg.num_historical_steps,
        time_span=cfg.time_span, pl2a_radius=cfg.pl2a_radius, pl2seed_radius=cfg.pl2seed_radius,
        a2a_radius=cfg.a2a_radius, a2sa_radius=cfg.a2sa_radius, pl2sa_radius=cfg.pl2sa_radius, num_freq_bands=64,
        num_layers=6, num_heads=8, head_dim=16, dropout=0.1, token_size=2048, attr_tokenizer=tok,
        predict_motion=True, predict_state=True, predict_map=False, predict_occ=True,
        state_token=dict(cfg.state_token), use_grid_token=True, use_head_token=True,
        use_state_token=cfg.use_state_token, disable_insertion=cfg.disable_insertion, seed_size=1, buffer_size=128,
        num_recurrent_steps_val=cfg.num_recurrent_steps_val,
        loss_weight={'state_cls_loss': 10, 'pos_cls_loss': 1, 'head_cls_loss': 1, 'shape_reg_loss': .2,
                     'seed_state_weight': [0.1, 0.9], 'seed_type_weight': [0.8, 0.1, 0.1]})
    # `attr_tokenizer.{grid,dist,dir}` are registered buffers derived from the config, not learned weights
    full = {k: v for k, v in dec.state_dict().items() if k.startswith('attr_tokenizer.')}
    full.update(state_dict)
    dec.load_state_dict(full, strict=True)
    dec.motion_beam_size = cfg.motion_beam_size
    dec.insert_beam_size = cfg.insert_beam_size
    dec.eval()
    return dec


def to_hetero(scene: Dict):
    """Wrap a synth scene into the HeteroData stand-in, adding the per-agent `token_traj_all` expansion."""
    shims.install()
    from torch_geometric.data import HeteroData
    from infgen_b200.synth import expand_token_traj_all
    d = HeteroData()
    for k, v in scene.items():
        if k == 'map_enc':
            continue
        if k == 'num_graphs':
            continue
        d[k] = dict(v) if isinstance(v, dict) else v
    d['agent']['token_traj_all'] = expand_token_traj_all(scene)
    return d.clone()


@torch.no_grad()
def run_reference(scene: Dict, state_dict: Dict[str, torch.Tensor], cfg, capture: bool = True,
                  decoder=None, forced_tokens: Optional[torch.Tensor] = None) -> Dict:
    """Returns {'out': reference output dict, 'trace': per-iteration captures}.

    trace['head_in'][t]      [A_t,128]  input of token_predict_head at iteration t (= last layer feature @cur)
    trace['token_logits'][t] [A_t,2048]
    trace['state_logits'][t] [A_t,3]
    """
    dec = decoder or build_reference_decoder(state_dict, cfg)
    trace: Dict[str, List[torch.Tensor]] = {'head_in': [], 'token_logits': [], 'state_logits': []}
    hooks = []
    if capture:
        hooks.append(dec.token_predict_head.register_forward_hook(
            lambda m, i, o: (trace['head_in'].append(i[0].detach().clone()),
                             trace['token_logits'].append(o.detach().clone())) and None))
        hooks.append(dec.state_predict_head.register_forward_hook(
            lambda m, i, o: trace['state_logits'].append(o.detach().clone())))
    data = to_hetero(scene)
    map_enc = {'x_pt': scene['map_enc']['x_pt'].clone()}
    dec.num_recurrent_steps_val = cfg.num_recurrent_steps_val      # undo the -1 -> 80 mutation (agent_decoder.py:1633)
    try:
        out = dec.inference(data, map_enc)
    finally:
        for h in hooks:
            h.remove()
    return {'out': out, 'trace': trace}
