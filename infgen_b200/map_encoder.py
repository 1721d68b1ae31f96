"""`B200MapEncoder`: drop-in for the reference `InfGenMapDecoder.forward` (infgen/modules/map_decoder.py:70-130) -
SURVEY.md section 8f row f1, the module `InfGenDecoder.inference` runs right before the agent decode
(infgen/modules/infgen_decoder.py:123-130).

    enc = B200MapEncoder.from_state_dict(map_encoder.state_dict(), traj_src)      # own engine
    out = enc.forward(data)                                                        # same dict keys as the reference

or share one engine with the agent decoder (`B200AgentDecoder(agent_sd, map_state_dict=map_sd)`), in which case `x_pt`
never leaves HBM between the two calls (`infgen_load_scenes` with `x_pt == NULL`).  The radius graph, the relative
embedding and the three pt2pt AttentionLayers run in libinfgen_b200.so (map.cuh + k_fourier_tc + k_attn / k_node); this
module only gathers the per-token fields the reference reads.  No CPU fallback.
"""
import ctypes as C
import os
from typing import Dict, Optional, Sequence
import numpy as np
import torch

from . import _capi

MAP_TOKEN_SIZE, MAP_TOKEN_DIM = 1024, 22


def load_map_vocab() -> torch.Tensor:
    """`map_token['traj_src']` [1024, 11, 2] of the reference's map vocabulary (infgen/tokens/map_traj_token5.pkl, converted
    to tokens/map_traj_token5.npz by tools/convert_tokens.py)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tokens', 'map_traj_token5.npz')
    return torch.from_numpy(np.load(path)['traj_src'])


def map_token_fields(data: Dict) -> Dict[str, np.ndarray]:
    """The per-token inputs of `InfGenMapDecoder.forward` (map_decoder.py:75-90) as contiguous numpy arrays."""
    pt = data['pt_token']
    f = lambda t: np.ascontiguousarray(torch.as_tensor(t).detach().cpu().numpy())
    if 'light_type' in pt:                                     # already gathered per token
        light = f(pt['light_type'])
    else:
        token2pl = data[('pt_token', 'to', 'map_polygon')]['edge_index']
        light = f(torch.as_tensor(data['map_polygon']['light_type'])[torch.as_tensor(token2pl)[1]])
    return {
        'pos': np.ascontiguousarray(f(pt['position'])[:, :2], dtype=np.float32),
        'ori': f(pt['orientation']).astype(np.float32), 'type': f(pt['type']).astype(np.int32),
        'pl_type': f(pt['pl_type']).astype(np.int32), 'light_type': light.astype(np.int32),
        'token_idx': f(pt['token_idx']).astype(np.int32),
    }


def encode_on(engine, lib, fields: Sequence[Dict[str, np.ndarray]], pl2pl_radius: float = 10.0,
              want_x: bool = True, want_logits: bool = False):
    """`infgen_map_encode` of the concatenated tokens of several scenes; returns (x_pt [P,128] or None, logits or None)."""
    ptr = np.zeros(len(fields) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([f['pos'].shape[0] for f in fields])
    cat = {k: np.ascontiguousarray(np.concatenate([f[k] for f in fields])) for k in fields[0]}
    P = int(ptr[-1])
    mb = _capi.MapBatch(n_scenes=len(fields), pt_ptr=_capi.i32p(ptr), pt_pos=_capi.f32p(cat['pos']), pt_ori=_capi.f32p(cat['ori']),
                        type=_capi.i32p(cat['type']), pl_type=_capi.i32p(cat['pl_type']), light_type=_capi.i32p(cat['light_type']),
                        token_idx=_capi.i32p(cat['token_idx']), pl2pl_radius=float(pl2pl_radius))
    x = np.empty((P, 128), dtype=np.float32) if want_x else None
    lg = np.empty((P, MAP_TOKEN_SIZE), dtype=np.float32) if want_logits else None
    _capi.check(lib.infgen_map_encode(engine, C.byref(mb), _capi.HOST, _capi.f32p(x), _capi.f32p(lg)))
    return x, lg, ptr


class B200MapEncoder:
    """Stand-alone map encoder (an engine that holds only the map weights) or a view on a `B200AgentDecoder` engine built
    with `map_state_dict`."""

    def __init__(self, map_state_dict: Optional[Dict[str, torch.Tensor]] = None, traj_src: Optional[torch.Tensor] = None,
                 pl2pl_radius: float = 10.0, device: int = 0, share=None):
        self.pl2pl_radius = pl2pl_radius
        if share is not None:
            self._dec, self._own = share, False
        else:
            from .agent_decoder import B200AgentDecoder
            from .config import DecoderConfig
            sd = {k[len('encoder.map_encoder.'):] if k.startswith('encoder.map_encoder.') else k: v
                  for k, v in map_state_dict.items()}
            self._dec = B200AgentDecoder(None, DecoderConfig(disable_insertion=True), device=device, map_state_dict=sd,
                                         map_traj_src=traj_src)
            self._own = True

    @classmethod
    def from_state_dict(cls, map_state_dict, traj_src=None, **kw):
        return cls(map_state_dict, traj_src, **kw)

    @classmethod
    def from_reference(cls, map_encoder, **kw):
        """Build from a live reference `InfGenMapDecoder` (weights stay owned by the module)."""
        return cls(map_encoder.state_dict(), torch.as_tensor(map_encoder.map_token['traj_src']), **kw)

    def close(self):
        if self._own:
            self._dec.close()

    def forward(self, data: Dict) -> Dict[str, torch.Tensor]:
        """`InfGenMapDecoder.forward(data)`: same keys, dtypes and shapes (map_decoder.py:124-130)."""
        pt = data['pt_token']
        x, lg, _ = encode_on(self._dec._h, self._dec.lib, [map_token_fields(data)], self.pl2pl_radius, True, True)
        pred = torch.as_tensor(pt['pt_pred_mask']).bool().cpu()
        logits = torch.from_numpy(lg)[pred]
        return {
            'x_pt': torch.from_numpy(x),
            'map_next_token_idx': torch.topk(torch.softmax(logits, dim=-1), k=10, dim=-1)[1],
            'map_next_token_prob': logits,
            'map_next_token_idx_gt': torch.as_tensor(pt['token_idx']).cpu()[torch.as_tensor(pt['pt_target_mask']).bool().cpu()],
            'map_next_token_eval_mask': pred[pred],
        }

    __call__ = forward


def install_map(map_encoder, **kw) -> B200MapEncoder:
    """Replace `map_encoder.forward` (reference InfGenMapDecoder) by the B200 path, in place."""
    import types
    enc = B200MapEncoder.from_reference(map_encoder, **kw)

    def _forward(self, data):
        dev = torch.as_tensor(data['pt_token']['position']).device
        return {k: v.to(dev) for k, v in enc.forward(data).items()}
    map_encoder.forward = types.MethodType(_forward, map_encoder)
    map_encoder._b200 = enc
    return enc
