"""Host mirror of the reference's per-scene preparation of the agent stream, running on the
engine's GPU through the C ABI `infgen_prepare_scene`:

    TokenProcessor._tokenize_agent   /root/reference/infgen/datasets/preprocess.py:364-550
    InfGen._fetch_enterings          /root/reference/infgen/model/infgen.py:1008-1090
    InfGen.match_token_map           /root/reference/infgen/model/infgen.py:918-984   (`infgen_match_map_tokens`)
    InfGen.sample_pt_pred            /root/reference/infgen/model/infgen.py:986-1006  (host: a torch.randperm over a few
                                     hundred mask slots - the reference's own RNG stream, so equal seeds give equal masks)

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

    def match_token_map(self, data: Dict, map_token: Dict = None, want_distance: bool = False) -> Dict:
        """`InfGen.match_token_map(data)`: reads data['map_save'] (traj_pos [P,3,2], traj_theta, pl_idx_list) and
        data['pt_token']['side'], writes data['pt_token'] (traj_mask, position, orientation, height, token_idx) and the
        pt_token -> map_polygon edges, as the reference does.  map_token: the reference's `self.map_token` dict (only
        'sample_pt' is read); default = the shipped map vocabulary sampled at its first / middle / last point."""
        ms = data['map_save']
        traj_pos = ms['traj_pos'].to(torch.float).contiguous()
        traj_theta = ms['traj_theta'].to(torch.float).contiguous()
        pl_idx = ms['pl_idx_list']
        side = data['pt_token']['side'].to(torch.uint8).contiguous()
        if map_token is not None and 'sample_pt' in map_token:
            sample_pt = map_token['sample_pt'].to(torch.float).contiguous()
        else:
            from .map_encoder import load_map_vocab
            src = load_map_vocab()
            sample_pt = src[:, torch.linspace(0, src.shape[1] - 1, steps=3).long()].contiguous()
        P, V = int(traj_pos.shape[0]), int(sample_pt.shape[0])
        if traj_pos.shape[1:] != (3, 2) or sample_pt.shape[1:] != (3, 2):
            raise ValueError('match_token_map: polylines and vocabulary entries are three 2-D points each')
        polygons, rank = torch.unique(pl_idx, sorted=True, return_inverse=True)
        rank = rank.to(torch.int32).contiguous()
        NP = int(polygons.numel())
        token_idx = torch.empty(P, dtype=torch.long)
        position, orientation = torch.empty(P, 3), torch.empty(P)
        counts = torch.empty(NP, 3, dtype=torch.int32)
        best = torch.empty(P) if want_distance else None
        pin = _capi.MapMatchIn(n_tokens=P, n_vocab=V, n_polygons=NP, traj_pos=_capi.f32p(traj_pos),
                               traj_theta=_capi.f32p(traj_theta), pl_rank=_capi.i32p(rank), side=_capi.u8p(side),
                               sample_pt=_capi.f32p(sample_pt))
        pout = _capi.MapMatchOut(token_idx=_capi.i64p(token_idx), position=_capi.f32p(position),
                                 orientation=_capi.f32p(orientation), side_counts=_capi.i32p(counts),
                                 best_distance=_capi.f32p(best) if want_distance else None)
        _capi.check(self.lib.infgen_match_map_tokens(self.dec._h, C.byref(pin), C.byref(pout)))
        longest = int(counts.max())
        traj_mask = torch.arange(longest)[None, None, :] < counts[:, :, None]
        pt = data['pt_token']
        pt['traj_mask'] = traj_mask
        pt['position'] = position
        pt['orientation'] = orientation
        pt['height'] = position[:, -1]
        pt['token_idx'] = token_idx
        data[('pt_token', 'to', 'map_polygon')] = {'edge_index': torch.stack([torch.arange(P), pl_idx.long()])}
        if want_distance:
            pt['match_distance'] = best
        return data

    @staticmethod
    def sample_pt_pred(data: Dict) -> Dict:
        """`InfGen.sample_pt_pred(data)`: the random third of the slots of every (polygon, side) row that the map head has
        to predict.  Same torch calls in the same order as the reference (one torch.randperm), so a caller that seeds torch
        like the reference gets the reference's masks."""
        traj_mask = data['pt_token']['traj_mask']
        n_pl, n_side, L = traj_mask.shape
        raw = torch.arange(1, L).repeat(n_pl, n_side, 1)
        k = (L - 1) // 3
        masked = raw.view(-1)[torch.randperm(raw.numel())[:n_pl * n_side * k].reshape(n_pl, n_side, k)]
        masked = torch.sort(masked, -1)[0]
        valid = traj_mask.clone()
        valid.scatter_(2, masked, False)
        pred = traj_mask.clone()
        pred.scatter_(2, masked, False)
        keep = torch.ones_like(pred)
        keep.scatter_(2, masked - 1, False)
        pred.masked_fill_(keep, False)
        pred = pred * torch.roll(traj_mask, shifts=-1, dims=2)
        target = torch.roll(pred, shifts=1, dims=2)
        pt = data['pt_token']
        pt['pt_valid_mask'], pt['pt_pred_mask'], pt['pt_target_mask'] = valid[traj_mask], pred[traj_mask], target[traj_mask]
        return data
