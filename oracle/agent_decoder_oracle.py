"""CPU oracle: a plain-torch fp32 restatement of the reference closed-loop decode (motion stage).

TEST INFRASTRUCTURE - NOT A PRODUCT PATH.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
leg may import this module; `infgen_b200/` never does and fails loudly when its CUDA library is missing.

What it restates (reference file:line, all under /root/reference):
  * `InfGenAgentDecoder.inference`          infgen/modules/agent_decoder.py:1605-2389  (motion stage; the insertion
                                            stage :1773-2114 is restated in `oracle/insertion_oracle.py` when present)
  * `_agent_token_embedding`                agent_decoder.py:332-406
  * `_build_vector_a`                       agent_decoder.py:426-447
  * `_build_agent_feature` + re-embed       agent_decoder.py:449-509, 2265-2287
  * `_build_temporal_edge`                  agent_decoder.py:540-610
  * `_build_interaction_edge`               agent_decoder.py:612-681
  * `_build_map2agent_edge`                 agent_decoder.py:683-758
  * `AttentionLayer`                        infgen/modules/layers.py:16-113
  * `FourierEmbedding/MLPEmbedding/MLPLayer` layers.py:116-215
  * `Attr_Tokenizer.encode_pos`             infgen/modules/attr_tokenizer.py:77-89
  * `angle_between_2d_vectors`, `wrap_angle` infgen/utils/func.py:30-34, 58-62
Third-party semantics (not vendored in the reference; pinned versions torch_cluster 1.6.3, torch_geometric 2.5.3,
environment.yml:287-289) are restated as: radius = strict `<`, first-k-by-index; PyG softmax with +1e-16;
propagate = index_add in edge order.

It is deliberately LITERAL about the reference's cost profile: every iteration pushes all A*T rows through the 18
attention layers and re-projects K/V of all T tiled map copies, exactly as the reference does, so timing it gives
an honest "port" CPU baseline.  The KV-cache the CUDA path uses is therefore *checked* by parity, not assumed.

Pinning: tests/test_oracle_vs_golden.py checks this file against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py, run in the build container under oracle/shims), and
tests/test_oracle_vs_reference.py does it live when /root/reference is present.
"""
import math
from typing import Callable, Dict, List, Optional
import torch
import torch.nn.functional as F

INVALID, VALID, ENTER, EXIT = 0, 1, 2, 3
SEED_TYPE = 3
H, NH, HD = 128, 8, 16


# ------------------------------------------------------------------------------------------------------------
# small ops
# ------------------------------------------------------------------------------------------------------------
def ln(W, p, x):
    return F.layer_norm(x, (x.shape[-1],), W[p + '.weight'], W[p + '.bias'], 1e-5)


def lin(W, p, x):
    return F.linear(x, W[p + '.weight'], W.get(p + '.bias'))


def wrap_angle(a: torch.Tensor) -> torch.Tensor:
    """func.py:58-62 (python-style modulo, constants rounded to fp32 by torch's scalar promotion)."""
    return -math.pi + (a + math.pi) % (2 * math.pi)


def angle_between(ctr: torch.Tensor, nbr: torch.Tensor) -> torch.Tensor:
    """func.py:30-34."""
    return torch.atan2(ctr[..., 0] * nbr[..., 1] - ctr[..., 1] * nbr[..., 0],
                       (ctr[..., :2] * nbr[..., :2]).sum(-1))


def fourier_embedding(W, p, x, cat: Optional[torch.Tensor] = None):
    """layers.py:142-160. x [N,D]; `cat` is the already-summed categorical embedding [N,128] (or None)."""
    f = x.unsqueeze(-1) * W[p + '.freqs.weight'] * 2 * math.pi
    f = torch.cat([f.cos(), f.sin(), x.unsqueeze(-1)], dim=-1)            # [N,D,129]
    acc = None
    for d in range(x.shape[1]):
        h = lin(W, f'{p}.mlps.{d}.0', f[:, d])
        h = torch.relu(ln(W, f'{p}.mlps.{d}.1', h))
        h = lin(W, f'{p}.mlps.{d}.3', h)
        acc = h if acc is None else acc + h
    if cat is not None:
        acc = acc + cat
    return lin(W, p + '.to_out.2', torch.relu(ln(W, p + '.to_out.0', acc)))


def mlp_embedding(W, p, x):
    """layers.py:170-177, 189."""
    h = torch.relu(ln(W, p + '.mlp.1', lin(W, p + '.mlp.0', x)))
    h = torch.relu(ln(W, p + '.mlp.4', lin(W, p + '.mlp.3', h)))
    return lin(W, p + '.mlp.6', h)


def mlp_layer(W, p, x):
    """layers.py:206-215."""
    return lin(W, p + '.mlp.3', torch.relu(ln(W, p + '.mlp.1', lin(W, p + '.mlp.0', x))))


def attention_layer(W, p, x_src, x_dst, r, src, dst, bipartite: bool):
    """layers.py:61-113 on flat node arrays. src/dst: int64 [E] (edge j -> i). Returns new x_dst [N,128]."""
    xs = ln(W, p + '.attn_prenorm_x_src', x_src)
    xd = ln(W, p + '.attn_prenorm_x_dst', x_dst) if bipartite else xs
    q = lin(W, p + '.to_q', xd).view(-1, NH, HD)
    k = lin(W, p + '.to_k', xs).view(-1, NH, HD)
    v = lin(W, p + '.to_v', xs).view(-1, NH, HD)
    kj, vj = k[src], v[src]
    if r is not None:
        rn = ln(W, p + '.attn_prenorm_r', r)
        kj = kj + lin(W, p + '.to_k_r', rn).view(-1, NH, HD)
        vj = vj + lin(W, p + '.to_v_r', rn).view(-1, NH, HD)
    sim = (q[dst] * kj).sum(-1) * (HD ** -0.5)                              # [E,8]
    n = x_dst.shape[0]
    idx = dst[:, None].expand_as(sim)
    smax = sim.new_full((n, NH), float('-inf')).scatter_reduce(0, idx, sim, 'amax', include_self=True)
    e = (sim - smax[dst]).exp()
    den = sim.new_zeros((n, NH)).scatter_add_(0, idx, e) + 1e-16
    alpha = e / den[dst]
    agg = x_dst.new_zeros((n, NH, HD)).index_add_(0, dst, vj * alpha.unsqueeze(-1)).view(n, NH * HD)
    g = torch.sigmoid(lin(W, p + '.to_g', torch.cat([agg, xd], dim=-1)))
    u = agg + g * (lin(W, p + '.to_s', xd) - agg)
    x = x_dst + ln(W, p + '.attn_postnorm', lin(W, p + '.to_out', u))
    ff = lin(W, p + '.ff_mlp.3', torch.relu(lin(W, p + '.ff_mlp.0', ln(W, p + '.ff_prenorm', x))))
    return x + ln(W, p + '.ff_postnorm', ff)


# ------------------------------------------------------------------------------------------------------------
# grid tokenizer (attr_tokenizer.py:24-43, 77-89)
# ------------------------------------------------------------------------------------------------------------
def build_grid(grid_range=150.0, grid_interval=3.0, radius=75.0) -> torch.Tensor:
    n = int(grid_range / grid_interval) + 1
    ax = torch.linspace(0, n - 1, steps=n)
    gx, gy = torch.meshgrid(ax, ax, indexing='xy')
    g = torch.stack([gx.flatten(), gy.flatten()], -1).reshape(n, n, 2).flip(dims=[0]).reshape(-1, 2)
    g = (g - n // 2) * grid_interval
    d = (g ** 2).sum(-1).sqrt()
    return g[(d <= radius)].contiguous()


def encode_pos(grid, x, y, theta_y):
    """Nearest grid cell of x[N,2] in the frame of ego (y[1,2], heading theta_y[1]); attr_tokenizer.py:77-89."""
    rel = x - y
    th = -(theta_y - math.pi / 2).expand(x.shape[0])
    c, s = th.cos(), th.sin()
    rot = torch.zeros(x.shape[0], 2, 2)
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1] = c, s, -s, c
    rel = torch.bmm(rel[:, None], rot)[:, 0]
    dist = ((rel[:, None] - grid[None]) ** 2).sum(-1).sqrt()
    return torch.argmin(dist, dim=-1)


# ------------------------------------------------------------------------------------------------------------
# sampler shared with the CUDA path (reference uses torch.multinomial: agent_decoder.py:2194 - same distribution)
# ------------------------------------------------------------------------------------------------------------
def uniform01(seed: int, scene: int, row: int, it: int) -> float:
    """Counter-based uniform in [0,1): lowbias32 mix of (seed, scene, row, iteration). Mirrored in csrc/sampler.cuh."""
    m = 0xFFFFFFFF
    x = (seed * 0x9E3779B1 + scene * 0x85EBCA77 + row * 0xC2B2AE3D + it * 0x27D4EB2F + 0x165667B1) & m
    x ^= x >> 16
    x = (x * 0x7FEB352D) & m
    x ^= x >> 15
    x = (x * 0x846CA68B) & m
    x ^= x >> 16
    return (x >> 8) * (1.0 / 16777216.0)


def sample_topk(logits: torch.Tensor, k: int, seed: int, scene: int, it: int) -> torch.Tensor:
    """softmax -> top-k -> draw proportional to the k probabilities (agent_decoder.py:2162-2163, 2194-2195).
    Inverse-CDF over the descending top-k list with the counter-based uniform above; k == 1 is greedy argmax."""
    prob = torch.softmax(logits, dim=-1)
    top_p, top_i = torch.topk(prob, k=k, dim=-1)
    if k == 1:
        return top_i[:, 0]
    out = torch.empty(logits.shape[0], dtype=torch.long)
    tp = top_p.to(torch.float32)
    for a in range(logits.shape[0]):
        total = torch.tensor(0.0)
        for j in range(k):
            total = total + tp[a, j]
        thr = torch.tensor(uniform01(seed, scene, a, it), dtype=torch.float32) * total
        c = torch.tensor(0.0)
        pick = k - 1
        for j in range(k):
            c = c + tp[a, j]
            if bool(thr < c):
                pick = j
                break
        out[a] = top_i[a, pick]
    return out


# ------------------------------------------------------------------------------------------------------------
# feature builders
# ------------------------------------------------------------------------------------------------------------
def build_vector_a(pos, head, state):
    """agent_decoder.py:426-447 (the `==` no-op at :444 is kept as a no-op)."""
    mv = torch.cat([pos.new_zeros(pos.shape[0], 1, 2), pos[:, 1:] - pos[:, :-1]], dim=1)
    inv = state == INVALID
    mv[inv] = -2.0
    prev = state.roll(shifts=1, dims=1)
    last_inv = (prev == INVALID) & ~inv
    last_inv[:, 0] = state[:, 0] == ENTER
    mv[last_inv] = 1.0
    last_val = (prev != INVALID) & inv
    last_val[:, 0] = False
    mv[last_val] = -1.0
    hv = torch.stack([head.cos(), head.sin()], dim=-1)
    return mv, hv


def embed_columns(W, tok_emb, mv, hv, cat, state, grid_emb):
    """agent_decoder.py:480-507 / 2271-2286: [.., 512] -> fusion_emb -> [..,128]. All inputs are [A,T,*]."""
    A, T = state.shape
    feat = torch.stack([mv.norm(p=2, dim=-1), angle_between(hv, mv)], dim=-1)
    x_a = fourier_embedding(W, 'x_a_emb', feat.view(-1, 2), cat.reshape(-1, H)).view(A, T, H)
    s_a = W['state_a_emb.weight'][state.reshape(-1)].view(A, T, H)
    return mlp_embedding(W, 'fusion_emb', torch.cat([tok_emb, x_a, s_a, grid_emb], dim=-1))


# ------------------------------------------------------------------------------------------------------------
# edge builders (dst restricted to the inference column(s))
# ------------------------------------------------------------------------------------------------------------
def temporal_edges(W, cfg, pos, head, state, hv, temporal_mask, inference_mask):
    """agent_decoder.py:540-610. Flat index = a*T + c (agent-major). Returns src, dst, r_raw[E,4], r_emb[E,128]."""
    A, T = state.shape
    hist = temporal_mask.clone()
    inf = inference_mask.clone()
    is_bos = state == ENTER
    bos = torch.where(is_bos.any(1), is_bos.long().argmax(1), torch.tensor(0))
    col = torch.arange(T)[None].expand(A, T)
    hist[col < bos[:, None]] = False
    q = cfg.num_seed_feature                                  # "last num_graphs*10 rows" quirk (:553-556)
    hist[-q:] = False
    inf[-q:] = False
    is_bos2 = is_bos.clone()
    is_bos2[-q:] = False
    bos2 = torch.where(is_bos2.any(1), is_bos2.long().argmax(1), torch.tensor(0))
    start = torch.clamp(bos2 - cfg.time_span / cfg.shift + 1, min=0)
    hist[~(col >= start[:, None])] = False
    m = hist.unsqueeze(2) & inf.unsqueeze(1)                  # [A, src col, dst col]
    nz = m.nonzero()
    src, dst = nz[:, 0] * T + nz[:, 1], nz[:, 0] * T + nz[:, 2]
    keep = (dst - src > 0) & (dst - src <= cfg.time_span / cfg.shift)
    src, dst = src[keep], dst[keep]
    p, h, hvf, inv = pos.reshape(-1, 2), head.reshape(-1), hv.reshape(-1, 2), (state == INVALID).reshape(-1)
    rp = p[src] - p[dst]
    rh = wrap_angle(h[src] - h[dst])
    rp[inv[src] & ~inv[dst]] = -1.0
    rp[~inv[src] & inv[dst]] = 1.0
    rh[inv[src] & ~inv[dst]] = -1.0
    rp[inv[src] & inv[dst]] = -2.0
    rh[inv[src] & inv[dst]] = -2.0
    raw = torch.stack([rp.norm(p=2, dim=-1), angle_between(hvf[dst], rp), rh, (src - dst).float()], dim=-1)
    return src, dst, raw, fourier_embedding(W, 'r_t_emb', raw)


def interaction_edges(W, cfg, pos, head, state, hv, interact_mask, inference_mask):
    """agent_decoder.py:612-659 (+ torch_cluster.radius_graph, PyG subgraph). Flat index = c*A + a (step-major)."""
    A, T = state.shape
    node_ok = (interact_mask & inference_mask).t().reshape(-1)
    ps, hs, hvs = pos.transpose(0, 1).reshape(-1, 2), head.t().reshape(-1), hv.transpose(0, 1).reshape(-1, 2)
    inv = (state == INVALID).t().reshape(-1)
    srcs, dsts = [], []
    for c in inference_mask.any(0).nonzero().flatten().tolist():
        pc = pos[:, c]
        d = pc[:, None] - pc[None]                              # [dst, src]
        within = (d * d).sum(-1) < cfg.a2a_radius ** 2
        rank = within.cumsum(1)
        within &= rank <= cfg.max_a2a_neighbors + 1
        within.fill_diagonal_(False)
        nz = within.nonzero()
        dsts.append(c * A + nz[:, 0])
        srcs.append(c * A + nz[:, 1])
    src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.long)
    dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.long)
    keep = node_ok[src] & node_ok[dst]
    src, dst = src[keep], dst[keep]
    rp = ps[src] - ps[dst]
    rh = wrap_angle(hs[src] - hs[dst])
    rp[inv[src] & ~inv[dst]] = -1.0
    rp[~inv[src] & inv[dst]] = 1.0
    rh[inv[src] & ~inv[dst]] = -1.0
    rp[inv[src] & inv[dst]] = -2.0
    rh[inv[src] & inv[dst]] = -2.0
    raw = torch.stack([rp.norm(p=2, dim=-1), angle_between(hvs[dst], rp), rh], dim=-1)
    return src, dst, raw, fourier_embedding(W, 'r_a2a_emb', raw)


def map_edges(W, cfg, pos, head, state, hv, interact_mask, inference_mask, pt_pos, pt_ori):
    """agent_decoder.py:683-729 (+ torch_cluster.radius). src flat = c*P + p, dst flat = c*A + a."""
    A, T = state.shape
    P = pt_pos.shape[0]
    node_ok = (interact_mask & inference_mask).t().reshape(-1)
    ps, hs, hvs = pos.transpose(0, 1).reshape(-1, 2), head.t().reshape(-1), hv.transpose(0, 1).reshape(-1, 2)
    inv = (state == INVALID).t().reshape(-1)
    srcs, dsts = [], []
    for c in inference_mask.any(0).nonzero().flatten().tolist():
        d = pos[:, c][:, None] - pt_pos[None, :, :2]            # [agent, map token]
        within = (d * d).sum(-1) < cfg.pl2a_radius ** 2
        within &= within.cumsum(1) <= cfg.max_pl2a_neighbors
        nz = within.nonzero()
        dsts.append(c * A + nz[:, 0])
        srcs.append(c * P + nz[:, 1])
    src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.long)
    dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.long)
    keep = node_ok[dst]
    src, dst = src[keep], dst[keep]
    rp = pt_pos[src % P, :2] - ps[dst]
    ro = wrap_angle(pt_ori[src % P] - hs[dst])
    rp[inv[dst]] = 1.0
    ro[inv[dst]] = 1.0
    raw = torch.stack([rp.norm(p=2, dim=-1), angle_between(hvs[dst], rp), ro], dim=-1)
    return src, dst, raw, fourier_embedding(W, 'r_pt2a_emb', raw)


# ------------------------------------------------------------------------------------------------------------
# insertion stage helpers (agent_decoder.py:449-509, 760-904, attr_tokenizer.py:91-110)
# ------------------------------------------------------------------------------------------------------------
def decode_pos(grid, index, y, theta_y):
    """attr_tokenizer.py:91-99: grid cell -> world position in the frame of ego (y [2], heading theta_y scalar)."""
    cx = grid[index.long()]                                    # [N,2]
    th = (theta_y - math.pi / 2).expand(cx.shape[0])
    c, s = th.cos(), th.sin()
    rot = torch.zeros(cx.shape[0], 2, 2)
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1] = c, s, -s, c
    cx = torch.bmm(cx[:, None], rot)[:, 0]
    return (cx + y).float()


def decode_heading(index, angle_interval=3.0):
    """attr_tokenizer.py:106-110."""
    ang = index * angle_interval - 180
    return (ang / 360 * (2 * math.pi)).float()


def seed_query_feature(W, T, grid_tab, G):
    """`_build_agent_feature(num_step, device, None, None, state_index=invalid, n=1)` (agent_decoder.py:449-509) as the
    insertion stage calls it (:1817-1821): no-token embedding, centre-cell grid embedding, all-invalid motion,
    seed type / 0.1 shape, invalid state."""
    tok = W['no_token_emb.weight'][0][None, None].repeat(1, T, 1)
    gemb = grid_tab[G // 2][None, None].repeat(1, T, 1)
    mv, hv = build_vector_a(torch.zeros(1, T, 2), torch.zeros(1, T), torch.full((1, T), INVALID))
    cat = W['type_a_emb.weight'][SEED_TYPE][None] + mlp_embedding(W, 'shape_emb', torch.full((1, 3), 0.1))
    state = torch.full((1, T), INVALID)
    return embed_columns(W, tok, mv, hv, cat.expand(T, H).reshape(1, T, H), state, gemb)


def a2sa_edges(W, emb, pos_p, head_p, hv_p, mask_a, mask_sa, cur, r, max_nb):
    """`_build_a2sa_edge` (agent_decoder.py:760-849) with the inference masks restricted to column `cur`.
    pos_p/head_p/hv_p: [N,T,*] (N rows incl. the query row(s)); mask_a, mask_sa: [N] booleans AT column cur.
    Flat index = c*N + a.  Returns src, dst (flat), r embedding."""
    N, T = head_p.shape
    pc = pos_p[:, cur]
    q_rows = mask_sa.nonzero().flatten()
    srcs, dsts = [], []
    for qi in q_rows.tolist():
        d = pc - pc[qi]
        within = (d * d).sum(-1) < r * r
        within &= within.cumsum(0) <= max_nb                   # torch_cluster.radius: first max_num_neighbors by index
        cand = within.nonzero().flatten()
        cand = cand[~mask_sa[cand] & mask_a[cand]]
        srcs.append(cur * N + cand)
        dsts.append(torch.full_like(cand, cur * N + qi))
    src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.long)
    dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.long)
    ps, hs, hvs = pos_p.transpose(0, 1).reshape(-1, 2), head_p.t().reshape(-1), hv_p.transpose(0, 1).reshape(-1, 2)
    rp = ps[src] - ps[dst]
    rh = wrap_angle(hs[src] - hs[dst])
    raw = torch.stack([rp.norm(p=2, dim=-1), angle_between(hvs[dst], rp), rh], dim=-1)
    return src, dst, raw, fourier_embedding(W, emb, raw)


def map2sa_edges(W, emb, pos_p, head_p, hv_p, mask_sa, cur, r, max_nb, pt_pos, pt_ori):
    """`_build_map2sa_edge` (agent_decoder.py:851-904). src flat = c*P + p, dst flat = c*N + a."""
    N, T = head_p.shape
    P = pt_pos.shape[0]
    q_rows = mask_sa.nonzero().flatten()
    srcs, dsts = [], []
    for qi in q_rows.tolist():
        d = pt_pos[:, :2] - pos_p[qi, cur]
        within = (d * d).sum(-1) < r * r
        within &= within.cumsum(0) <= max_nb
        cand = within.nonzero().flatten()
        srcs.append(cur * P + cand)
        dsts.append(torch.full_like(cand, cur * N + qi))
    src = torch.cat(srcs) if srcs else torch.zeros(0, dtype=torch.long)
    dst = torch.cat(dsts) if dsts else torch.zeros(0, dtype=torch.long)
    ps, hs, hvs = pos_p.transpose(0, 1).reshape(-1, 2), head_p.t().reshape(-1), hv_p.transpose(0, 1).reshape(-1, 2)
    rp = pt_pos[src % P, :2] - ps[dst]
    ro = wrap_angle(pt_ori[src % P] - hs[dst])
    raw = torch.stack([rp.norm(p=2, dim=-1), angle_between(hvs[dst], rp), ro], dim=-1)
    return src, dst, raw, fourier_embedding(W, emb, raw)


# ------------------------------------------------------------------------------------------------------------
# the rollout
# ------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def rollout(scene: Dict, W: Dict[str, torch.Tensor], cfg, seed: int = 2024, scene_id: int = 0,
            forced_tokens: Optional[torch.Tensor] = None, forced_states: Optional[torch.Tensor] = None,
            collect_trace: bool = False, max_iters: Optional[int] = None,
            assume_no_insertion: bool = False, debug_force_enter: bool = False) -> Dict:
    """Closed-loop decode of one scene (agent_decoder.py:1605-2389): insertion stage :1744-2114 (when enabled) and
    motion stage :2116-2301.  debug_force_enter mirrors the reference's `DEBUG=1` switch (:1888-1889), which makes
    the seed head answer 'enter' - the only way to exercise insertion with random-init weights.

    forced_tokens / forced_states: optional [A, S] int64 teacher-forcing overrides of the sampled motion token /
    predicted state per iteration (used to compare implementations step by step without divergence).
    """
    # assume_no_insertion: the caller knows (from a golden log / a biased seed head) that the insertion stage inserts
    # nobody; it then has no effect on the motion stage (whose state head stays live) and is skipped.
    run_insertion = not cfg.disable_insertion and not assume_no_insertion
    ag = scene['agent']
    HC = cfg.hist_cols
    nh = cfg.num_historical_steps
    filt = ag['state_idx'][:, HC - 1] != INVALID
    eval_mask = ag['valid_mask'][filt, nh - 1]
    agent_id = ag['id'][filt].clone()
    valid = ag['raw_agent_valid_mask'][filt].clone()
    pos = ag['token_pos'][filt].clone()
    token = ag['token_idx'][filt].clone()
    state = ag['state_idx'][filt].clone()
    head = ag['token_heading'][filt].clone()
    shape_a = ag['shape'][filt].clone()
    type_a = ag['type'][filt].clone().long()
    grid_a = ag['grid_token_idx'][filt].clone()
    gt_traj = ag['position'][filt, nh:, :2].contiguous()
    vocab = torch.stack([ag['trajectory_token_veh'], ag['trajectory_token_ped'], ag['trajectory_token_cyc']])
    pt_pos, pt_ori = scene['pt_token']['position'], scene['pt_token']['orientation']
    x_pt = scene['map_enc']['x_pt']
    P = pt_pos.shape[0]

    n_rec = cfg.num_recurrent_steps_val
    if n_rec == -1:
        n_rec = ag['position'].shape[1] - nh
    A, T0 = state.shape
    T = (n_rec + nh) // cfg.shift
    if T < T0:
        raise ValueError('horizon shorter than the scene is unsupported by the reference (agent_decoder.py:1638)')
    if T > T0:
        padn = T - T0
        valid = torch.cat([valid, torch.ones(A, padn, dtype=torch.bool)], 1)
        pos = torch.cat([pos, torch.zeros(A, padn, 2)], 1)
        token = torch.cat([token, torch.full((A, padn), -1)], 1)
        state = torch.cat([state, torch.zeros(A, padn, dtype=torch.long)], 1)
        head = torch.cat([head, torch.zeros(A, padn)], 1)
        grid_a = torch.cat([grid_a, torch.full((A, padn), -1)], 1)
    av0 = int(ag['av_index'][0])
    av = av0 - int((~filt[:av0]).sum())

    pos[:, HC:], head[:, HC:], token[:, HC:], state[:, HC:], grid_a[:, HC:] = 0, 0, -1, 0, -1
    valid[:, HC:] = True
    valid[~eval_mask] = False
    hist_token = ag['token_idx'][filt]
    hist_state = ag['state_idx'][filt]

    # ---- embedding tables (agent_decoder.py:347-373) --------------------------------------------------------
    tok_tab = []
    for ti, nm in enumerate(('veh', 'ped', 'cyc')):
        e = mlp_embedding(W, f'token_emb_{nm}', vocab[ti][:, -1].flatten(1, 2))
        tok_tab.append(torch.cat([e, W['bos_token_emb.weight'], W['no_token_emb.weight']]))   # [-2]=bos, [-1]=none
    tok_tab = torch.stack(tok_tab)                                                            # [3,2050,128]
    grid = build_grid(cfg.grid_range, cfg.grid_interval, cfg.pl2seed_radius)
    grid_tab = torch.cat([mlp_embedding(W, 'token_emb_grid', grid), W['invalid_offset_token_emb.weight']])
    seed_type_emb = W['type_a_emb.weight'][SEED_TYPE]
    inv_shape_emb = mlp_embedding(W, 'shape_emb', torch.full((1, 3), 0.1))[0]

    tok_emb = torch.zeros(A, T, H)
    for ti in range(3):
        m = type_a == ti
        tok_emb[m] = tok_tab[ti][token[m]]
    is_inv = state == INVALID
    types_at = type_a[:, None].repeat(1, T)
    types_at[is_inv] = SEED_TYPE
    shapes_at = shape_a[:, nh - 1][:, None].repeat(1, T, 1)
    shapes_at[is_inv] = 0.1
    type_emb = W['type_a_emb.weight'][types_at]                                               # [A,T,128]
    shape_emb = mlp_embedding(W, 'shape_emb', shapes_at.reshape(-1, 3)).view(A, T, H)
    mv, hv = build_vector_a(pos, head, state)
    feat0 = embed_columns(W, tok_emb, mv, hv, type_emb + shape_emb, state, grid_tab[grid_a])   # raw_feat_a

    # ---- masks (agent_decoder.py:1695-1719) -------------------------------------------------------------------
    mask = valid.clone()
    is_bos, is_eos = state == ENTER, state == EXIT
    bos = torch.where(is_bos.any(1), is_bos.long().argmax(1), torch.tensor(0))
    eos = torch.where(is_eos.any(1), is_eos.long().argmax(1), torch.tensor(T - 1))
    col = torch.arange(T)[None].expand(A, T)
    motion_mask = (col > bos[:, None]) & (col <= eos[:, None])
    motion_mask[:, nh // cfg.shift:] = False
    temporal_mask = torch.ones_like(mask)
    temporal_mask[motion_mask] = mask[motion_mask]
    interact_mask = torch.ones_like(mask)
    non_motion = ~motion_mask
    non_motion[:, nh // cfg.shift:] = False
    interact_mask[non_motion] = False
    interact_mask[state == ENTER] = True
    interact_mask[av] = True
    temporal_mask[:, HC:] = True
    interact_mask[:, HC:] = True

    S = n_rec // cfg.shift
    A0 = A
    G = grid.shape[0]
    pred_traj = torch.zeros(A, n_rec, 2)
    pred_head = torch.zeros(A, n_rec)
    pred_state = torch.zeros(A, n_rec)
    pred_type = type_a.clone()
    pred_shape = shape_a[:, HC - 1]
    vtype = type_a.clone()                                                # vocabulary (box-track table) of each row
    max_agent_id = int(agent_id.max())
    next_token_list = [hist_token[:, i:i + 1] for i in range(HC)]
    next_state_list = [hist_state[:, i:i + 1] for i in range(HC)]
    seed_state_prob, seed_pos_prob, seed_agent_occ, seed_pt_occ, seed_occ_gt = [], [], [], [], []
    insert_log: List[Dict] = []
    pass_log: List[Dict] = []
    cache: Dict[int, torch.Tensor] = {}
    trace: List[Dict] = []
    insert_limit = 10

    n_iter = S if max_iters is None else min(S, max_iters)
    for t in range(n_iter):
        cur, nxt = HC - 1 + t, HC + t
        A = pos.shape[0]
        # ================= 1. insertion stage (agent_decoder.py:1744-2114) ====================================
        st_prob = torch.zeros(11, 1)
        pos_prob, ag_occ, pt_occ, occ_gt = (torch.zeros(11, 1, G) for _ in range(4))
        n_new = 0
        p_pass = 0
        while run_insertion:
            p_pass += 1
            if t == 0 or p_pass - 1 >= insert_limit:
                break
            A = pos.shape[0]
            # the query ("seed") row is a copy of the ego row (`_pad_feat`, :511-526)
            pos_p = torch.cat([pos, pos[av:av + 1]])
            head_p = torch.cat([head, head[av:av + 1]])
            hv_p = torch.cat([hv, hv[av:av + 1]])
            is_seed = torch.zeros(A + 1, dtype=torch.bool)
            is_seed[-1] = True
            inter_p = torch.cat([interact_mask[:, cur], torch.ones(1, dtype=torch.bool)])
            feat_seed = seed_query_feature(W, T, grid_tab, G)
            raw_feat = feat0.clone()
            feat = torch.cat([feat0, feat_seed])
            e_a2s = a2sa_edges(W, 'r_a2sa_emb', pos_p, head_p, hv_p, inter_p, is_seed, cur, cfg.pl2seed_radius, 300)
            e_p2s = map2sa_edges(W, 'r_pt2sa_emb', pos_p, head_p, hv_p, is_seed, cur, cfg.pl2seed_radius, 2048,
                                 pt_pos, pt_ori)
            occ = torch.zeros(G, dtype=torch.long)
            g_cur = grid_a[:, cur]
            occ[g_cur[g_cur != -1]] = 1
            occ_emb = mlp_layer(W, 'seed_agent_occ_embed', occ.reshape(1, G).float())          # [1,128]
            occ_src = torch.zeros(1, dtype=torch.long)
            occ_dst = torch.tensor([cur * (A + 1) + A])
            x_pt_t = x_pt.repeat(T, 1)
            for i in range(3):
                x = feat.transpose(0, 1).reshape(-1, H)
                x = attention_layer(W, f'occ2sa_attn_layers.{i}', occ_emb, x, None, occ_src, occ_dst, True)
                x = attention_layer(W, f'pt2sa_attn_layers.{i}', x_pt_t, x, e_p2s[3], e_p2s[0], e_p2s[1], True)
                x = attention_layer(W, f'a2sa_attn_layers.{i}', x, x, e_a2s[3], e_a2s[0], e_a2s[1], False)
                feat = x.view(T, A + 1, H).transpose(0, 1)
            q = feat[-1:, cur]                                                               # [1,128]
            if collect_trace:                                      # seed-query taps of every pass (debug tools)
                pass_log.append({'t': t, 'pass': p_pass - 1, 'q': q[0].clone(), 'occ': occ.clone(), 'occ_emb': occ_emb[0].clone(),
                                 'as_src': (e_a2s[0] % (A + 1)).clone(), 'as_raw': e_a2s[2].clone(),
                                 'ps_src': (e_p2s[0] % P).clone(), 'ps_raw': e_p2s[2].clone(),
                                 'x_sa_in': feat0[:, cur].clone()})
            ego_pos, ego_head = pos[av, cur], head[av, cur]
            g_ag = mlp_layer(W, 'grid_agent_occ_head', q)
            g_pt = mlp_layer(W, 'grid_pt_occ_head', q)
            s_prob = mlp_layer(W, 'seed_state_predict_head', q)
            s_idx = s_prob.softmax(-1).argmax(-1, keepdim=True)
            s_idx = torch.where(s_idx == 1, torch.tensor(ENTER), torch.tensor(INVALID))
            if debug_force_enter:
                s_idx = torch.full_like(s_idx, ENTER)
            ty_idx = mlp_layer(W, 'seed_type_predict_head', q).softmax(-1).argmax(-1, keepdim=True)
            shp = mlp_layer(W, 'seed_shape_predict_head', q)
            p_soft = torch.softmax(mlp_layer(W, 'seed_pos_rel_token_predict_head', q), dim=-1)
            top_p, top_i = torch.topk(p_soft, k=cfg.insert_beam_size, dim=-1)
            if cfg.insert_beam_size > 1:
                # reference: torch.multinomial(topk_prob, 1) (:1900); here the counter-based draw shared with the
                # CUDA path, keyed by (scene, pass, iteration) - same distribution
                tp = top_p[0].to(torch.float32)
                total = torch.tensor(0.0)
                for j in range(cfg.insert_beam_size):
                    total = total + tp[j]
                thr = torch.tensor(uniform01(seed ^ 0x5EED, scene_id, p_pass - 1, t), dtype=torch.float32) * total
                c_acc, pick = torch.tensor(0.0), cfg.insert_beam_size - 1
                for j in range(cfg.insert_beam_size):
                    c_acc = c_acc + tp[j]
                    if bool(thr < c_acc):
                        pick = j
                        break
                cell = top_i[:, pick:pick + 1]
            else:
                cell = top_i[:, :1]
            new_pos = decode_pos(grid, cell[..., 0], ego_pos, ego_head)
            if bool(occ[cell[0, 0]]):                                                        # overlap filter (:1906-1909)
                feat0 = raw_feat
                continue
            if bool((s_idx == INVALID).all()) or n_new + 1 > insert_limit:
                feat0 = raw_feat
                break
            n_new += 1
            # ---- 1.5 append the new agent (:1923-1995) -------------------------------------------------------
            agent_id = torch.cat([agent_id, torch.tensor([max_agent_id + 1], dtype=agent_id.dtype)])
            max_agent_id += 1
            mask = torch.cat([mask, torch.ones(1, T, dtype=torch.bool)])
            temporal_mask = torch.cat([temporal_mask, torch.ones(1, T, dtype=torch.bool)])
            interact_mask = torch.cat([interact_mask, torch.ones(1, T, dtype=torch.bool)])
            n_pos, n_head = torch.zeros(1, T, 2), torch.zeros(1, T)
            n_grid = torch.full((1, T), -1, dtype=grid_a.dtype)
            n_state = torch.full((1, T), INVALID, dtype=state.dtype)
            n_shape = torch.full((1, T, 3), 0.1)
            n_type = torch.full((1, T), SEED_TYPE, dtype=torch.long)
            n_pos[:, cur] = new_pos
            n_head[:, cur] = ego_head                                                        # dummy value
            n_grid[:, cur] = cell[0, 0]
            n_type[:, cur:] = ty_idx[0, 0]
            n_shape[:, cur:] = shp[:, None]
            n_state[:, cur] = s_idx[0, 0]
            pos, head, grid_a, state = torch.cat([pos, n_pos]), torch.cat([head, n_head]), \
                torch.cat([grid_a, n_grid]), torch.cat([state, n_state])
            pred_type = torch.cat([pred_type, n_type[:, cur].to(pred_type.dtype)])
            pred_shape = torch.cat([pred_shape, n_shape[:, cur]])
            vtype = torch.cat([vtype, ty_idx[0].to(vtype.dtype)])
            mask[-1:, :cur + 1] = False
            interact_mask[-1:, :cur] = False
            npt, nph, nps = torch.zeros(1, n_rec, 2), torch.zeros(1, n_rec), torch.zeros(1, n_rec)
            npt[:, (t - 1) * 5:t * 5] = n_pos[:, cur, None].repeat(1, 5, 1)
            nph[:, (t - 1) * 5:t * 5] = n_head[:, cur, None].repeat(1, 5)
            nps[:, (t - 1) * 5:t * 5] = s_idx.repeat(1, 5).float()
            pred_traj, pred_head, pred_state = torch.cat([pred_traj, npt]), torch.cat([pred_head, nph]), \
                torch.cat([pred_state, nps])
            n_tok = W['no_token_emb.weight'][0][None, None].repeat(1, T, 1)
            n_tok[:, cur] = W['bos_token_emb.weight'][0]
            tok_emb = torch.cat([tok_emb, n_tok])
            n_type_emb = W['type_a_emb.weight'][n_type]                                      # [1,T,128]
            n_shape_emb = mlp_embedding(W, 'shape_emb', n_shape.reshape(-1, 3)).view(1, T, H)
            type_emb, shape_emb = torch.cat([type_emb, n_type_emb]), torch.cat([shape_emb, n_shape_emb])
            # ---- 2. heading of the new agent (:2003-2074) ----------------------------------------------------
            mv_s, hv_s = build_vector_a(pos[-1:], head[-1:], state[-1:])
            assert bool((mv_s[:, :cur] == -2.0).all()) and bool((mv_s[:, cur] == 1.0).all())
            mv_s[:, cur + 2:] = 0.0
            hv_s[:, cur + 2:] = 0.0
            mv, hv = torch.cat([mv, mv_s]), torch.cat([hv, hv_s])
            n_gemb = grid_tab[n_grid]
            feat_s = embed_columns(W, n_tok, mv_s, hv_s, n_type_emb + n_shape_emb, n_state, n_gemb)
            feat = torch.cat([raw_feat, feat_s])
            is_new = torch.zeros(A + 1, dtype=torch.bool)
            is_new[-1] = True
            e_a = a2sa_edges(W, 'r_a2a_emb', pos, head, hv, interact_mask[:, cur], is_new, cur, cfg.a2sa_radius, 24)
            e_p = map2sa_edges(W, 'r_pt2a_emb', pos, head, hv, is_new, cur, cfg.pl2sa_radius, 128, pt_pos, pt_ori)
            for i in range(3):
                x = feat.transpose(0, 1).reshape(-1, H)
                x = attention_layer(W, f'pt2a_attn_layers.{i}', x_pt_t, x, e_p[3], e_p[0], e_p[1], True)
                x = attention_layer(W, f'a2a_attn_layers.{i}', x, x, e_a[3], e_a[0], e_a[1], False)
                feat = x.view(T, A + 1, H).transpose(0, 1)
            hq = feat[-1:, cur]
            h_idx = mlp_layer(W, 'seed_heading_rel_token_predict_head', hq).softmax(-1).argmax(-1, keepdim=True)
            new_head = wrap_angle(decode_heading(h_idx, cfg.angle_interval) + ego_head)
            head[-1:, cur] = new_head[:, 0]
            off = torch.tanh(mlp_layer(W, 'seed_offset_xy_predict_head', hq)) * 2
            pos[-1:, cur] += off
            mv_s, hv_s = build_vector_a(pos[-1:], head[-1:], state[-1:])
            mv_s[:, cur + 2:] = 0.0
            hv_s[:, cur + 2:] = 0.0
            mv[-1:] = mv_s
            hv[-n_new:] = hv_s                                 # sic (:2083): every agent inserted this iteration
            feat_s = embed_columns(W, n_tok, mv_s, hv_s, n_type_emb + n_shape_emb, state[-1:], n_gemb)
            feat0 = torch.cat([raw_feat, feat_s])
            occ_gt[n_new], ag_occ[n_new], pt_occ[n_new], pos_prob[n_new] = occ.float(), g_ag, g_pt, p_soft
            st_prob[n_new] = s_prob.softmax(-1)[:, -1]
            insert_log.append({'t': t, 'cell': int(cell[0, 0]), 'type': int(ty_idx[0, 0]), 'pos': pos[-1, cur].clone(),
                               'head': head[-1, cur].clone(), 'shape': shp[0].clone(), 'heading_token': int(h_idx[0, 0]),
                               # heading-stage taps (debug tools): agent / map neighbours of the new row, head input
                               'ha_src': (e_a[0] % (A + 1)).clone(), 'ha_raw': e_a[2].clone(), 'hp_src': (e_p[0] % P).clone(),
                               'hp_raw': e_p[2].clone(), 'hq': hq[0].clone(), 'offset': off[0].clone(), 'q': q[0].clone(),
                               'as_src': (e_a2s[0] % (A + 1)).clone(), 'feat_in': feat_s[0, cur].clone()})
        seed_state_prob.append(st_prob)
        seed_pos_prob.append(pos_prob)
        seed_agent_occ.append(ag_occ)
        seed_pt_occ.append(pt_occ)
        seed_occ_gt.append(occ_gt)
        next_state_list[-1] = torch.cat([next_state_list[-1], torch.full((n_new, 1), ENTER, dtype=torch.long)])

        # ================= 3. motion stage (agent_decoder.py:2116-2301) =======================================
        A = pos.shape[0]
        x_pt_tiled = x_pt.repeat(T, 1)                                        # [T*P,128], index c*P+p (:2143)
        # the t==0 mask built at :1760-1764 is overwritten at :2119-2121 before the motion edges are built, so the
        # only destination column is `cur` in every iteration (column 0 runs through the layers edge-less at t=0)
        inf_mask = torch.zeros_like(temporal_mask)
        inf_mask[:, cur] = True
        e_t = temporal_edges(W, cfg, pos, head, state, hv, temporal_mask, inf_mask)
        e_a = interaction_edges(W, cfg, pos, head, state, hv, interact_mask, inf_mask)
        e_m = map_edges(W, cfg, pos, head, state, hv, interact_mask, inf_mask, pt_pos, pt_ori)

        feat = feat0
        layer_out = []
        for i in range(6):
            if i in cache:
                feat = cache[i]
            x = attention_layer(W, f't_attn_layers.{i}', feat.reshape(-1, H), feat.reshape(-1, H),
                                e_t[3], e_t[0], e_t[1], False)
            x = x.view(A, T, H).transpose(0, 1).reshape(-1, H)                               # step-major
            x = attention_layer(W, f'pt2a_attn_layers.{i}', x_pt_tiled, x, e_m[3], e_m[0], e_m[1], True)
            x = attention_layer(W, f'a2a_attn_layers.{i}', x, x, e_a[3], e_a[0], e_a[1], False)
            feat = x.view(T, A, H).transpose(0, 1)
            if t == 0:
                cache[i + 1] = feat.clone()
            else:
                n = cache[i + 1].shape[0]
                cache[i + 1][:n, cur] = feat[:n, cur]
                if A > n:                                       # rows inserted this iteration: all their columns (:2156-2158)
                    cache[i + 1] = torch.cat([cache[i + 1], feat[n:]])
            if collect_trace:
                layer_out.append(feat[:, cur].clone())

        hin = feat[:, cur]
        logits = mlp_layer(W, 'token_predict_head', hin)
        s_logits = mlp_layer(W, 'state_predict_head', hin)
        nstate = s_logits.softmax(-1).argmax(-1)
        nstate[nstate == 2] = EXIT
        nstate[av] = VALID
        if not cfg.use_state_token:
            nstate[nstate == EXIT] = VALID
        if cfg.disable_insertion:
            nstate[:] = VALID
        ntoken = sample_topk(logits, cfg.motion_beam_size, seed, scene_id, t)
        if forced_tokens is not None:
            ntoken = forced_tokens[:A, t].clone()
        if forced_states is not None:
            nstate = forced_states[:A, t].clone()

        # ---- advance (agent_decoder.py:2175-2239) -----------------------------------------------------------
        box = vocab[vtype.clamp(max=2), ntoken]                                             # [A,6,4,2]
        th = head[:, cur]
        c, s = th.cos(), th.sin()
        rot = torch.zeros(A, 2, 2)
        rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1] = c, s, -s, c
        world = torch.bmm(box.view(A, 24, 2), rot).view(A, 6, 4, 2) + pos[:, None, None, cur]
        d = world[:, 1:, 0] - world[:, 1:, 3]
        pred_traj[:, t * 5:(t + 1) * 5] = world[:, 1:].mean(dim=2)
        pred_head[:, t * 5:(t + 1) * 5] = torch.atan2(d[..., 1], d[..., 0])
        pred_state[:, t * 5:(t + 1) * 5] = nstate[:, None].float()
        pos[:, nxt] = world[:, -1].mean(dim=1)
        d = world[:, -1, 0] - world[:, -1, 3]
        theta = torch.atan2(d[:, 1], d[:, 0])
        head[:, nxt] = theta
        state[:, nxt] = nstate
        grid_a[:, nxt] = encode_pos(grid, pos[:, nxt], pos[av:av + 1, nxt], theta[av:av + 1])
        inv_n = nstate == INVALID
        ntoken = ntoken.clone()
        ntoken[inv_n] = -1
        pos[inv_n, nxt] = 0.0
        head[inv_n, nxt] = 0.0
        grid_a[inv_n, nxt] = -1
        mask[inv_n, nxt] = False
        interact_mask[inv_n, nxt] = False
        tok_emb[inv_n, nxt] = W['no_token_emb.weight'][0]
        type_emb[inv_n, nxt] = seed_type_emb
        shape_emb[inv_n, nxt] = inv_shape_emb
        for ti in range(3):
            m = vtype == ti
            tok_emb[m, nxt] = tok_tab[ti][ntoken[m]]

        # ---- re-embed (agent_decoder.py:2265-2287) ------------------------------------------------------------
        mv, hv = build_vector_a(pos, head, state)
        mv[:, nxt + 1:] = 0.0
        hv[:, nxt + 1:] = 0.0
        feat0 = embed_columns(W, tok_emb, mv, hv, type_emb + shape_emb, state, grid_tab[grid_a])
        next_token_list.append(ntoken[:, None])
        next_state_list.append(nstate[:, None])
        if collect_trace:
            def _dst_only(e, div):
                return {'src': e[0].clone(), 'dst': e[1].clone(), 'raw': e[2].clone(), 'emb': e[3].clone()}
            trace.append({'t': t, 'cur': cur, 'n_rows': A, 'n_new': n_new, 'edges_t': _dst_only(e_t, T),
                          'edges_a': _dst_only(e_a, A),
                          'edges_m': _dst_only(e_m, A), 'layer_out': torch.stack(layer_out), 'head_in': hin.clone(),
                          'token_logits': logits.clone(), 'state_logits': s_logits.clone(),
                          'token': ntoken.clone(), 'state': nstate.clone(), 'pos_next': pos[:, nxt].clone(),
                          'head_next': head[:, nxt].clone(), 'grid_next': grid_a[:, nxt].clone(),
                          'feat_next': feat0[:, nxt].clone()})

    # ---- outputs (agent_decoder.py:2303-2389) ---------------------------------------------------------------
    A = pos.shape[0]
    for i in range(len(next_token_list)):
        k = A - next_token_list[i].shape[0]
        next_token_list[i] = torch.cat([next_token_list[i], torch.full((k, 1), -1, dtype=torch.long)]).long()
        k = A - next_state_list[i].shape[0]
        next_state_list[i] = torch.cat([next_state_list[i], torch.zeros(k, 1, dtype=torch.long)]).long()
    pred_traj = torch.cat([torch.zeros(A, nh, 2), pred_traj], 1)
    pred_head = torch.cat([torch.zeros(A, nh), pred_head], 1)
    pred_state = torch.cat([torch.zeros(A, nh), pred_state], 1)
    pred_traj[:A0, 0] = ag['position'][filt, 0, :2]
    pred_head[:A0, 0] = ag['heading'][filt, 0]
    pred_state[:A0, 1:nh] = hist_state[:, :HC].repeat_interleave(cfg.shift, dim=1).float()
    htok = hist_token[:, :HC].clone()
    htok[htok < 0] = 0
    hbox = vocab[type_a.clamp(max=2)[:, None].expand(A0, HC), htok]                         # [A0,HC,6,4,2]
    th = head[:A0, 0]
    c, s = th.cos(), th.sin()
    rot = torch.zeros(A0, 2, 2)
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1] = c, s, -s, c
    hworld = torch.bmm(hbox.reshape(A0, HC * 24, 2), rot).view(A0, HC, 6, 4, 2) + pos[:A0, 0][:, None, None, None]
    pred_traj[:A0, 1:nh] = hworld[:, :, 1:].mean(dim=3).reshape(A0, -1, 2)
    d = hworld[:, :, 1:, 0] - hworld[:, :, 1:, 3]
    pred_head[:A0, 1:nh] = torch.atan2(d[..., 1], d[..., 0]).reshape(A0, -1)
    pred_valid = (pred_state != INVALID) & (pred_state != ENTER)
    eval_shape = torch.zeros_like(pred_shape)
    for ti, key in enumerate(('vehicle', 'pedstrain', 'cyclist')):
        from infgen_b200.config import AGENT_SHAPE
        eval_shape[vtype == ti] = torch.tensor(AGENT_SHAPE[key])
    out = {
        'ego_index': av, 'agent_id': agent_id, 'valid_mask': valid, 'pos_a': pos, 'head_a': head, 'gt_traj': gt_traj,
        'pred_traj': pred_traj, 'pred_head': pred_head, 'pred_type': pred_type, 'pred_state': pred_state,
        'pred_z': torch.zeros_like(pred_traj[..., 0]), 'pred_shape': pred_shape, 'eval_shape': eval_shape,
        'pred_valid': pred_valid,
        'next_token_idx': torch.cat(next_token_list, dim=-1), 'next_state_idx': torch.cat(next_state_list, dim=-1),
        'next_state_prob_seed': torch.cat(seed_state_prob, dim=1) if seed_state_prob else None,
        'next_pos_rel_prob_seed': torch.cat(seed_pos_prob, dim=1) if seed_pos_prob else None,
        'grid_agent_occ_seed': torch.cat(seed_agent_occ, dim=1) if seed_agent_occ else None,
        'grid_pt_occ_seed': torch.cat(seed_pt_occ, dim=1) if seed_pt_occ else None,
        'grid_agent_occ_gt_seed': torch.cat(seed_occ_gt, dim=1) if seed_occ_gt else None,
    }
    return {'out': out, 'trace': trace, 'insert_log': insert_log, 'pass_log': pass_log}
