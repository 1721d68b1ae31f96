"""CPU restatement of the reference map encoder `InfGenMapDecoder.forward` (infgen/modules/map_decoder.py:70-130) -
SURVEY.md section 8f row f1, the next row after the decode path.

TEST INFRASTRUCTURE ONLY (same rules as agent_decoder_oracle.py): only tests/, __graft_entry__.smoke() and the
cpu_baseline leg of bench.py may import it.  Pinned: tests/test_oracle_map_vs_golden.py checks it against golden vectors
written by the UNMODIFIED reference (tests/golden/make_golden_map.py, run in the build container through oracle/shims).

Third-party semantics (torch_cluster 1.6.3 `radius_graph`, not vendored in the reference; defined by oracle/shims/cluster.py):
strict `dist^2 < r^2`, per target the first `max_num_neighbors + 1` candidates by ascending index (the target itself
included) with the self loop dropped afterwards, edges ordered by target then source.
"""
from typing import Dict, Tuple
import torch

from .agent_decoder_oracle import (fourier_embedding, mlp_embedding, mlp_layer, attention_layer, wrap_angle, angle_between)


def radius_graph_first_k(pos: torch.Tensor, r: float, max_num_neighbors: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """map_decoder.py:91-93 `radius_graph(x=pos_pt[:, :2], r, loop=False, max_num_neighbors=100)` -> (src, dst)."""
    d = pos[:, None, :] - pos[None, :, :]                       # [target, candidate]
    within = (d * d).sum(-1) < float(r) * float(r)
    keep = within & (within.cumsum(dim=1) <= max_num_neighbors + 1)
    nz = keep.nonzero()
    dst, src = nz[:, 0], nz[:, 1]
    m = src != dst
    return src[m], dst[m]


def map_encode(W: Dict[str, torch.Tensor], pt: Dict[str, torch.Tensor], traj_src: torch.Tensor, pl2pl_radius: float = 10.0,
               max_num_neighbors: int = 100) -> Dict[str, torch.Tensor]:
    """pt: position [P,3], orientation [P], type [P], pl_type [P], token_idx [P], light_type [P] (the polygon's light type
    gathered per token, map_decoder.py:85-86), pt_pred_mask [P] bool.  traj_src [1024,11,2]."""
    pos = pt['position'][:, :2].contiguous()
    ori = pt['orientation'].contiguous()
    ov = torch.stack([ori.cos(), ori.sin()], dim=-1)
    tok_tab = mlp_embedding(W, 'token_emb', traj_src.reshape(traj_src.shape[0], -1).float())       # :79-81
    x = tok_tab[pt['token_idx'].long()]
    cat = torch.stack([W['type_pt_emb.weight'][pt['type'].long()], W['polygon_type_emb.weight'][pt['pl_type'].long()],
                       W['light_pl_emb.weight'][pt['light_type'].long()]]).sum(dim=0)                # :87-90
    x = x + cat
    src, dst = radius_graph_first_k(pos, pl2pl_radius, max_num_neighbors)
    rel = pos[src] - pos[dst]
    rel_o = wrap_angle(ori[src] - ori[dst])
    r = torch.stack([torch.norm(rel, p=2, dim=-1), angle_between(ov[dst], rel), rel_o], dim=-1)      # :98-104
    r = fourier_embedding(W, 'r_pt2pt_emb', r)
    for i in range(3):
        x = attention_layer(W, f'pt2pt_layers.{i}', x, x, r, src, dst, bipartite=False)
    logits = mlp_layer(W, 'token_predict_head', x[pt['pt_pred_mask']])
    return {'x_pt': x, 'map_next_token_prob': logits, 'map_next_token_idx': torch.softmax(logits, dim=-1).topk(10, dim=-1)[1],
            'edge_src': src, 'edge_dst': dst}
