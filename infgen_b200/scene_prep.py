"""Host mirror of the reference's per-scene preparation of the agent stream, running on the
engine's GPU through the C ABI `infgen_prepare_scene`:

    TokenProcessor._tokenize_agent   /root/reference/infgen/datasets/preprocess.py:364-550
    InfGen._fetch_enterings          /root/reference/infgen/model/infgen.py:1008-1090

`B200ScenePrep.tokenize(data)` takes the nested dict the reference's `TokenProcessor.forward` takes (`data['agent']` with
valid_mask / heading / position / velocity / type / shape / av_idx, `data['pt_token']['position']`) and writes the same keys
the two reference functions write, with the reference's dtypes (LongTensors, bool masks).  There is no CPU fallback.
"""
import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _capi


class B200ScenePrep:
    def __init__(self, decoder):
        """decoder: a B200AgentDecoder (its engine owns the vocabulary and the position grid on the device)."""
        self.dec = decoder
        self.lib = decoder.lib
        self.shift = 5

    def tokenize(self, data: Dict) -> Dict:
        ag = data['agent']
        valid = ag['valid_mask'].to(torch.uint8).contiguous()
        heading = ag['heading'].float().contiguous()
        position = ag['position'].float().contiguous()
        velocity = ag['velocity'].float().contiguous()
        a_type = ag['type'].to(torch.uint8).contiguous()
        av = ag['av_idx'] if 'av_idx' in ag else ag['av_index']
        av = int(av.reshape(-1)[0]) if isinstance(av, torch.Tensor) else int(av)
        pt = data['pt_token']['position'].float().contiguous() if 'pt_token' in data else torch.zeros(0, 3)
        if pt.shape[-1] == 2:
            pt = torch.cat([pt, torch.zeros(pt.shape[0], 1)], -1).contiguous()
        A, N = valid.shape
        T, P = N // self.shift, int(pt.shape[0])
        o = {
            'token_idx': torch.empty(A, T, dtype=torch.long), 'state_idx': torch.empty(A, T, dtype=torch.long),
            'token_contour': torch.empty(A, T, 4, 2), 'token_pos': torch.empty(A, T, 2), 'token_heading': torch.empty(A, T),
            'raw_agent_valid_mask': torch.empty(A, T, dtype=torch.uint8), 'agent_valid_mask': torch.empty(A, T, dtype=torch.uint8),
            'grid_token_idx': torch.empty(A, T, dtype=torch.long), 'grid_offset_xy': torch.empty(A, T, 2),
            'heading_token_idx': torch.empty(A, T, dtype=torch.long), 'pos_xy': torch.empty(A, T, 2),
            'heading_theta': torch.empty(A, T), 'sort_indices': torch.empty(A, T, dtype=torch.long),
            'inrange_mask': torch.empty(A, T, dtype=torch.uint8), 'bos_mask': torch.empty(A, T, dtype=torch.uint8),
            'pt_grid_token_idx': torch.empty(T, P, dtype=torch.long),
        }
        pin = _capi.PrepIn(n_agents=A, n_steps=N, av_index=av, n_pt=P, valid_mask=_capi.u8p(valid),
                           heading=_capi.f32p(heading), position=_capi.f32p(position), velocity=_capi.f32p(velocity),
                           type=_capi.u8p(a_type), pt_position=_capi.f32p(pt) if P else None)
        pout = _capi.PrepOut()
        for k, t in o.items():
            if t.dtype == torch.long:
                setattr(pout, k, _capi.i64p(t))
            elif t.dtype == torch.uint8:
                setattr(pout, k, _capi.u8p(t))
            else:
                setattr(pout, k, _capi.f32p(t))
        _capi.check(self.lib.infgen_prepare_scene(self.dec._h, C.byref(pin), C.byref(pout)))
        for k in ('raw_agent_valid_mask', 'agent_valid_mask', 'inrange_mask', 'bos_mask'):
            o[k] = o[k].bool()
        # reset agent shapes to the first fully specified one (preprocess.py:521-524) - host side, three floats per agent
        shape = ag['shape'].clone()
        nz = torch.all(shape != 0., dim=-1)
        first = torch.argmax(nz.long(), dim=1)
        if not bool(nz.any(dim=1).all()):
            raise ValueError('Found invalid shape values.')
        o['shape'] = shape[torch.arange(A), first][:, None, :].expand(-1, shape.shape[1], -1).contiguous()
        ag.update(o)
        ag['av_index'] = ag.get('av_idx', ag.get('av_index'))
        return data
