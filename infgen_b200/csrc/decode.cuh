// Decode-loop kernels: everything of `InfGenAgentDecoder.inference` (agent_decoder.py:1740-2301) that is not an
// nn.Module call - edge construction, sampling, token->pose advance, grid tokenisation, next-column inputs.
#pragma once
#include "common.cuh"
#include "ops.cuh"

namespace infgen {

constexpr int ST_INVALID = 0, ST_VALID = 1, ST_ENTER = 2, ST_EXIT = 3;
constexpr int RING = 16;           // temporal K/V ring depth (>= window + 1, power of two)

struct DecState {
    int n_scenes, cap, T, S, HC, W;
    int q_rows;                    // num_seed_feature: tail rows without temporal edges (agent_decoder.py:553-556)
    int G, V;                      // grid cells, motion-token vocabulary
    int max_m;                     // max_pl2a_neighbors
    int max_a;                     // max_a2a_neighbors (agent_decoder.py:633)
    int a_stride;                  // agent<->agent slots per row: min(cap, max_a + 1)
    float r_m2, r_a2;              // squared radii
    int use_state_token, disable_insertion, beam;
    int teacher_forced;            // InfGenAgentDecoder.forward: a temporal destination must itself be in the history mask
                                   // (mask_t = hist x hist, agent_decoder.py:577-579; inference: hist x inference_mask)
    unsigned seed;
    const int *n_rows, *ego_row, *scene_id;
    int *col, *iter;               // device scalars: current column / iteration
    float *pos, *head;             // [R][T][2], [R][T]
    int *state, *token, *grid;     // [R][T]
    uint8_t *interact, *tsrc;      // [R][T]
    const int *type;               // [R] (rows appended by the insertion stage get their predicted type)
    const int *ins_col;            // [R] insertion column of appended rows, -1 for the scene's own agents (NULL: none)
    int *hv_src;                   // [R] row whose heading gives this row's heading VECTOR during the current iteration,
                                   // -1 = its own (agent_decoder.py:2083 overwrites the vectors of every agent inserted
                                   // in an iteration with the newest one's until they are rebuilt at :2265)
    const int *pt_ptr;
    const float *pt_pos, *pt_ori;
    const float *grid_cells;       // [G][2]
    const float *vocab;            // [3][V][6][4][2]
    // edges
    int *t_cnt, *t_src; float *t_raw;                    // [R], [R*W], [R*W][4]
    int *m_cnt, *m_src; float *m_raw;                    // [R], [R*max_m], [R*max_m][3]
    int *a_cnt, *a_start, *a_total, *a_src; float *a_raw; // [R], [R], [n_scenes], [R*a_stride], [..][3]
    // next-column embedding inputs
    float *xa_raw;                 // [R][2]
    int *tok_row, *state_idx, *grid_row, *cat_idx;       // [R]
    // sampler inputs / outputs
    const float *part_v; const int *part_i; const float *part_m, *part_s; const float *state_logits;
    const int *forced_tok, *forced_state;                // [R][S] or NULL
    float *pred_traj, *pred_head, *pred_state;           // [R][5S][2], [R][5S], [R][5S]
    int *next_token, *next_state;                        // [R][T]
};

__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

// ---------------------------------------------------------------------------------------------------------------
// edges whose destination is column `cur` (agent_decoder.py:540-610, 612-681, 683-758 with the inference masks of
// :2119-2121).  One warp per (row, edge type).  Semantics of the third-party calls (oracle/shims): radius = strict `<`,
// the first max_num_neighbors sources by ascending index; edges ordered by destination then source.
// Slots: temporal r*W + k, map r*max_m + k, agent r*a_stride + k (k-th neighbour by ascending source row).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_edge_build(const DecState s, int col_add) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * NWARP + warp;
    const int r = gw / 3, kind = gw - 3 * r;
    const int R = s.n_scenes * s.cap;
    if (r >= R) return;
    const int b = r / s.cap, i = r - b * s.cap, n = s.n_rows[b];
    if (i >= n) return;
    const int col = *s.col + col_add, T = s.T;
    const int r0 = b * s.cap;
    const float px = s.pos[((size_t)r * T + col) * 2], py = s.pos[((size_t)r * T + col) * 2 + 1];
    const float hd = s.head[(size_t)r * T + col];
    const float hd_v = (s.hv_src && s.hv_src[r] >= 0) ? s.head[(size_t)s.hv_src[r] * T + col] : hd;
    const float hx = cosf(hd_v), hy = sinf(hd_v);
    const bool inv_d = s.state[(size_t)r * T + col] == ST_INVALID;
    const bool inter = s.interact[(size_t)r * T + col] != 0;
    if (kind == 0) {
        // ---- temporal: (r, c) -> (r, col), 0 < col - c <= W ------------------------------------------------------
        int cnt = 0;
        if (i < n - s.q_rows && (!s.teacher_forced || s.tsrc[(size_t)r * T + col] != 0)) {
            const int c = col - s.W + lane;
            const bool ok = lane < s.W && c >= 0 && s.tsrc[(size_t)r * T + c] != 0;
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int slot = r * s.W + __popc(mask & lanemask_lt());
                const bool inv_s = s.state[(size_t)r * T + c] == ST_INVALID;
                float rx = __fsub_rn(s.pos[((size_t)r * T + c) * 2], px);
                float ry = __fsub_rn(s.pos[((size_t)r * T + c) * 2 + 1], py);
                float rh = wrap_angle(__fsub_rn(s.head[(size_t)r * T + c], hd));
                if (inv_s && !inv_d) { rx = -1.f; ry = -1.f; rh = -1.f; }       // :595-601 sentinels
                if (!inv_s && inv_d) { rx = 1.f; ry = 1.f; }
                if (inv_s && inv_d) { rx = -2.f; ry = -2.f; rh = -2.f; }
                s.t_src[slot] = r * RING + (c & (RING - 1));
                float4 raw = make_float4(norm2(rx, ry), angle_between(hx, hy, rx, ry), rh, (float)(c - col));
                st4(s.t_raw + (size_t)slot * 4, raw);
            }
            cnt = __popc(mask);
        }
        if (lane == 0) s.t_cnt[r] = cnt;
    } else if (kind == 1) {
        // ---- map -> agent: first max_m tokens within the radius -------------------------------------------------
        const int pt0 = s.pt_ptr[b], pt1 = s.pt_ptr[b + 1];
        int cnt = 0;
        if (inter) {
            for (int p0 = pt0; p0 < pt1 && cnt < s.max_m; p0 += 32) {
                const int p = p0 + lane;
                float dx = 0.f, dy = 0.f;
                bool ok = false;
                if (p < pt1) {
                    dx = __fsub_rn(px, s.pt_pos[(size_t)p * 2]);
                    dy = __fsub_rn(py, s.pt_pos[(size_t)p * 2 + 1]);
                    ok = dist2(dx, dy) < s.r_m2;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                const int rank = cnt + __popc(mask & lanemask_lt());
                if (ok && rank < s.max_m) {
                    const int slot = r * s.max_m + rank;
                    float rx = -dx, ry = -dy;
                    float ro = wrap_angle(__fsub_rn(s.pt_ori[p], hd));
                    if (inv_d) { rx = 1.f; ry = 1.f; ro = 1.f; }                    // :722-723
                    s.m_src[slot] = p;
                    s.m_raw[(size_t)slot * 3 + 0] = norm2(rx, ry);
                    s.m_raw[(size_t)slot * 3 + 1] = angle_between(hx, hy, rx, ry);
                    s.m_raw[(size_t)slot * 3 + 2] = ro;
                }
                cnt = min(s.max_m, cnt + __popc(mask));
            }
        }
        if (lane == 0) s.m_cnt[r] = cnt;
    } else {
        // ---- agent <-> agent (:632-634): radius_graph over ALL rows of the scene = the first max_a + 1 rows within the
        //      radius by ascending index, the row itself included, then minus the self loop (torch_cluster.radius_graph
        //      queries max_num_neighbors + 1 and drops loops afterwards); `subgraph` then keeps the edges whose two ends
        //      interact at this column.  Rows that do not interact therefore still count towards the limit.
        int cnt = 0, seen = 0;
        const int base = r * s.a_stride;
        if (inter) {
            for (int j0 = 0; j0 < n && seen <= s.max_a; j0 += 32) {
                const int j = j0 + lane;
                const int rj = r0 + j;
                bool within = false;
                float rx = 0.f, ry = 0.f;
                if (j < n) {
                    rx = __fsub_rn(s.pos[((size_t)rj * T + col) * 2], px);
                    ry = __fsub_rn(s.pos[((size_t)rj * T + col) * 2 + 1], py);
                    // radius_graph tests |p_i - p_j|^2 < r^2 with the difference taken as (dst - src)
                    const float dx = __fsub_rn(px, s.pos[((size_t)rj * T + col) * 2]);
                    const float dy = __fsub_rn(py, s.pos[((size_t)rj * T + col) * 2 + 1]);
                    within = dist2(dx, dy) < s.r_a2;
                }
                const unsigned wm = __ballot_sync(0xffffffffu, within);
                const bool in_first = within && (seen + __popc(wm & lanemask_lt())) <= s.max_a;
                const bool ok = in_first && j != i && s.interact[(size_t)rj * T + col] != 0;
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const int slot = base + cnt + __popc(mask & lanemask_lt());
                    const bool inv_s = s.state[(size_t)rj * T + col] == ST_INVALID;
                    float rh = wrap_angle(__fsub_rn(s.head[(size_t)rj * T + col], hd));
                    if (inv_s && !inv_d) { rx = -1.f; ry = -1.f; rh = -1.f; }               // :647-653
                    if (!inv_s && inv_d) { rx = 1.f; ry = 1.f; }
                    if (inv_s && inv_d) { rx = -2.f; ry = -2.f; rh = -2.f; }
                    s.a_src[slot] = rj;
                    s.a_raw[(size_t)slot * 3 + 0] = norm2(rx, ry);
                    s.a_raw[(size_t)slot * 3 + 1] = angle_between(hx, hy, rx, ry);
                    s.a_raw[(size_t)slot * 3 + 2] = rh;
                }
                cnt += __popc(mask);
                seen += __popc(wm);
            }
        }
        if (lane == 0) { s.a_cnt[r] = cnt; s.a_start[r] = base; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// inputs of the column embedding (agent_decoder.py:426-447 _build_vector_a, :449-509 / :2265-2287) of one row
// ---------------------------------------------------------------------------------------------------------------
struct EmbedIn {
    float xa0, xa1;                // |motion vector|, angle(heading, motion vector)
    int tok_row, state_idx, grid_row, cat_idx;
};
__device__ __forceinline__ EmbedIn embed_inputs_row(const DecState &s, int r, int col) {
    const int T = s.T, R = s.n_scenes * s.cap;
    const int st = s.state[(size_t)r * T + col];
    const bool inv = st == ST_INVALID;
    float mx = 0.f, my = 0.f;
    bool last_inv, last_val;
    if (col == 0) {
        last_inv = st == ST_ENTER;
        last_val = false;
    } else {
        mx = __fsub_rn(s.pos[((size_t)r * T + col) * 2], s.pos[((size_t)r * T + col - 1) * 2]);
        my = __fsub_rn(s.pos[((size_t)r * T + col) * 2 + 1], s.pos[((size_t)r * T + col - 1) * 2 + 1]);
        const int pst = s.state[(size_t)r * T + col - 1];
        last_inv = pst == ST_INVALID && !inv;
        last_val = pst != ST_INVALID && inv;
    }
    if (inv) { mx = -2.f; my = -2.f; }
    if (last_inv) { mx = 1.f; my = 1.f; }
    if (last_val) { mx = -1.f; my = -1.f; }
    const float hd = s.head[(size_t)r * T + col];
    EmbedIn o;
    o.xa0 = norm2(mx, my);
    o.xa1 = angle_between(cosf(hd), sinf(hd), mx, my);
    const int tok = s.token[(size_t)r * T + col];
    o.tok_row = s.type[r] * (s.V + 2) + (tok < 0 ? s.V + 2 + tok : tok);   // [-2] = BOS row, [-1] = no-token row
    o.state_idx = st;
    const int g = s.grid[(size_t)r * T + col];
    o.grid_row = g < 0 ? s.G : g;                                            // [-1] = invalid-offset row
    // row R = seed type + 0.1 shape.  The reference builds the categorical embeddings once, while every future
    // column is still 'invalid' (agent_decoder.py:1653-1657, 458-470), and later only rewrites them for steps that
    // turn invalid (:2235-2239): every generated column therefore carries the seed/0.1 row, valid or not.
    // Rows appended by the insertion stage carry their predicted type / shape from the insertion column on (:1953-1954).
    const bool inserted = s.ins_col != nullptr && s.ins_col[r] >= 0;
    o.cat_idx = (inv || (col >= s.HC && !inserted)) ? R : r;
    return o;
}

// counter-based uniform in [0,1): mirrors oracle.agent_decoder_oracle.uniform01
__device__ __forceinline__ float uniform01(unsigned seed, unsigned scene, unsigned row, unsigned it) {
    unsigned x = seed * 0x9E3779B1u + scene * 0x85EBCA77u + row * 0xC2B2AE3Du + it * 0x27D4EB2Fu + 0x165667B1u;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------------------------------
// sampling + state update + token->pose advance + grid token (agent_decoder.py:2160-2262).  One warp per row; every
// warp first re-derives the ego row's new pose (the grid token is ego-relative, attr_tokenizer.py:77-89).
// ---------------------------------------------------------------------------------------------------------------
struct AdvOut {
    int tok, st;
    float lx, ly, lh;              // pose at column nxt
};
// warp-cooperative; all lanes return the same values.  write: store the row's results.
__device__ __forceinline__ AdvOut advance_row(const DecState &s, int b, int i, int col, int t, bool write) {
    const int lane = threadIdx.x & 31;
    const int T = s.T, nxt = col + 1;
    const int r = b * s.cap + i;
    // ---- merge the per-slice candidates: global max, softmax denominator, top-KTOP ------------------------------
    const float pm = lane < NSLICE ? s.part_m[(size_t)r * NSLICE + lane] : -INFINITY;
    const float gmax = warp_max(pm);
    const float pd = lane < NSLICE ? s.part_s[(size_t)r * NSLICE + lane] * expf(pm - gmax) : 0.f;
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < NSLICE; ++k) den += __shfl_sync(0xffffffffu, pd, k);      // slice order, as a serial loop
    constexpr int NC = NSLICE * KTOP;                                           // 40 candidates, <= 2 per lane
    float v0 = s.part_v[(size_t)r * NC + lane];
    int i0 = s.part_i[(size_t)r * NC + lane];
    float v1 = -INFINITY;
    int i1 = 0x7fffffff;
    if (lane + 32 < NC) { v1 = s.part_v[(size_t)r * NC + lane + 32]; i1 = s.part_i[(size_t)r * NC + lane + 32]; }
    float cv[KTOP];
    int ci[KTOP];
#pragma unroll
    for (int k = 0; k < KTOP; ++k) {
        float bv = v0; int bi = i0;
        if (v1 > bv || (v1 == bv && i1 < bi)) { bv = v1; bi = i1; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        cv[k] = bv; ci[k] = bi;
        if (i0 == bi) { v0 = -INFINITY; i0 = 0x7fffffff; }                      // token ids are unique
        if (i1 == bi) { v1 = -INFINITY; i1 = 0x7fffffff; }
    }
    int tok = ci[0];
    if (s.beam > 1) {                      // softmax -> top-k -> multinomial over the k probabilities (:2162-2163, 2194)
        float p[KTOP], total = 0.f;
        for (int k = 0; k < s.beam; ++k) { p[k] = expf(cv[k] - gmax) / den; total += p[k]; }
        const float thr = uniform01(s.seed, (unsigned)s.scene_id[b], (unsigned)i, (unsigned)t) * total;
        float c = 0.f; int pick = s.beam - 1;
        for (int k = 0; k < s.beam; ++k) { c += p[k]; if (thr < c) { pick = k; break; } }
        tok = ci[pick];
    }
    if (s.forced_tok) tok = s.forced_tok[(size_t)r * s.S + t];
    // ---- state (:2166-2173) -------------------------------------------------------------------------------------
    int st;
    {
        const float l0 = s.state_logits[(size_t)r * 4], l1 = s.state_logits[(size_t)r * 4 + 1],
                    l2 = s.state_logits[(size_t)r * 4 + 2];
        const float m = fmaxf(l0, fmaxf(l1, l2));
        const float e0 = expf(l0 - m), e1 = expf(l1 - m), e2 = expf(l2 - m);
        const float sum = e0 + e1 + e2;
        const float p0 = e0 / sum, p1 = e1 / sum, p2 = e2 / sum;
        st = 0; float bp = p0;
        if (p1 > bp) { bp = p1; st = 1; }
        if (p2 > bp) { bp = p2; st = 2; }
        if (st == 2) st = ST_EXIT;
        if (i == s.ego_row[b]) st = ST_VALID;
        if (!s.use_state_token && st == ST_EXIT) st = ST_VALID;
        if (s.disable_insertion) st = ST_VALID;
    }
    if (s.forced_state) st = s.forced_state[(size_t)r * s.S + t];
    // ---- token -> pose (:2175-2211): lane k computes sub-step k of the token's box track ------------------------
    const int ty = min(s.type[r], 2);
    const int tk = tok < 0 ? tok + s.V : tok;
    const float *box = s.vocab + ((size_t)ty * s.V + tk) * 48;
    const float px = s.pos[((size_t)r * T + col) * 2], py = s.pos[((size_t)r * T + col) * 2 + 1];
    const float th = s.head[(size_t)r * T + col];
    const float c = cosf(th), sn = sinf(th);
    float mx = 0.f, my = 0.f, hh = 0.f;
    if (lane >= 1 && lane < 6) {
        const int k = lane;
        float wx[4], wy[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float bx = box[(k * 4 + q) * 2], by = box[(k * 4 + q) * 2 + 1];
            wx[q] = __fadd_rn(__fadd_rn(__fmul_rn(bx, c), __fmul_rn(by, -sn)), px);
            wy[q] = __fadd_rn(__fadd_rn(__fmul_rn(bx, sn), __fmul_rn(by, c)), py);
        }
        mx = (((wx[0] + wx[1]) + wx[2]) + wx[3]) * 0.25f;
        my = (((wy[0] + wy[1]) + wy[2]) + wy[3]) * 0.25f;
        hh = atan2f(__fsub_rn(wy[0], wy[3]), __fsub_rn(wx[0], wx[3]));
        if (write) {
            const size_t o = (size_t)r * (5 * s.S) + t * 5 + (k - 1);
            s.pred_traj[o * 2] = mx; s.pred_traj[o * 2 + 1] = my;
            s.pred_head[o] = hh;
            s.pred_state[o] = (float)st;
        }
    }
    AdvOut out;
    out.lx = __shfl_sync(0xffffffffu, mx, 5);
    out.ly = __shfl_sync(0xffffffffu, my, 5);
    out.lh = __shfl_sync(0xffffffffu, hh, 5);
    const bool inv = st == ST_INVALID;
    if (inv) { tok = -1; out.lx = 0.f; out.ly = 0.f; out.lh = 0.f; }                      // :2221-2239
    out.tok = tok; out.st = st;
    if (write && lane == 0) {
        s.pos[((size_t)r * T + nxt) * 2] = out.lx; s.pos[((size_t)r * T + nxt) * 2 + 1] = out.ly;
        s.head[(size_t)r * T + nxt] = out.lh;
        s.state[(size_t)r * T + nxt] = st;
        s.token[(size_t)r * T + nxt] = tok;
        s.interact[(size_t)r * T + nxt] = inv ? 0 : 1;
        s.next_token[(size_t)r * T + nxt] = tok;
        s.next_state[(size_t)r * T + nxt] = st;
        if (inv) s.grid[(size_t)r * T + nxt] = -1;
    }
    return out;
}

__global__ void __launch_bounds__(NT) k_advance(const DecState s) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * NWARP + warp;
    const int R = s.n_scenes * s.cap;
    if (r >= R) return;
    const int b = r / s.cap, i = r - b * s.cap;
    if (i >= s.n_rows[b]) return;
    const int col = *s.col, t = *s.iter, T = s.T, nxt = col + 1;
    const int ego = s.ego_row[b];
    // the ego row (no stores) and this row: two independent chains of dependent loads, inlined back to back so that the
    // scheduler can overlap them (a rolled two-pass loop serialised them)
    const AdvOut eo = advance_row(s, b, ego, col, t, false);
    const AdvOut me = advance_row(s, b, i, col, t, true);
    if (me.st == ST_INVALID) return;
    // ---- ego-centric grid token of the new position (attr_tokenizer.py:77-89, agent_decoder.py:2214) ------------
    const float eth = -__fsub_rn(eo.lh, 1.5707963267948966f);
    const float ec = cosf(eth), es = sinf(eth);
    const float rx = __fsub_rn(me.lx, eo.lx), ry = __fsub_rn(me.ly, eo.ly);
    const float qx = __fadd_rn(__fmul_rn(rx, ec), __fmul_rn(ry, -es));
    const float qy = __fadd_rn(__fmul_rn(rx, es), __fmul_rn(ry, ec));
    float bd = INFINITY; int bi = 0x7fffffff;
#pragma unroll 8
    for (int g = lane; g < s.G; g += 32) {
        const float d = norm2(__fsub_rn(qx, s.grid_cells[(size_t)g * 2]), __fsub_rn(qy, s.grid_cells[(size_t)g * 2 + 1]));
        if (d < bd) { bd = d; bi = g; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) s.grid[(size_t)r * T + nxt] = bi;
}

// ===============================================================================================================
// Column embedding (agent_decoder.py:449-509 / 2265-2287): x_a_emb (Fourier, D=2, categorical seed) -> concat
// [token_emb | x_a_emb | state_emb | grid_emb] -> fusion_emb (512 -> 128 -> 128 -> 128), tiles of EM rows
// ===============================================================================================================
struct ColEmbArgs {
    RowSpace rows;
    FourierW fx;               // x_a_emb
    MlpEmbW fusion;
    DecState s;                // decode state: the embedded column is *s.col + col_add
    int col_add;
    const float *cat_tab;      // [R+1][128] type + shape embedding rows
    const float *tok_tab;      // [3][token_size+2][128]
    const float *state_tab;    // [4][128]
    const float *grid_tab;     // [grid_size+1][128]
    float *out;                // [R][128]
    float *out2, *out3;        // optional copies of the embedded rows (insertion stage: inputs of the two edge-less stacks)
};
constexpr int XLD = 516;       // leading dimension of the 512-wide fusion input
constexpr int COLEMB_SMEM_FLOATS = WS_SMEM_FLOATS + EM * XLD + EM * FLD + 2 * EM * HLD + EM * 4;
constexpr size_t COLEMB_SMEM = (size_t)COLEMB_SMEM_FLOATS * sizeof(float);

__global__ void __launch_bounds__(NT_S) k_embed_column(const ColEmbArgs a) {
    extern __shared__ __align__(16) float smem[];
    WsSmem wsm(smem);
    float *sX = smem + WS_SMEM_FLOATS;       // [EM][XLD]
    float *sF = sX + EM * XLD;               // [EM][132]
    float *sH = sF + EM * FLD;               // [EM][HLD]
    float *sA = sH + EM * HLD;               // [EM][HLD]
    float *sraw = sA + EM * HLD;             // [EM][4]
    __shared__ EmbedIn s_in[EM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ int s_row[EM];                // global row of every tile position, -1 = inactive
    if (tid < EM) {
        const int r = a.rows.tile_row(blockIdx.x, tid, EM);
        s_row[tid] = a.rows.active_row(r) ? r : -1;
    }
    __syncthreads();
    bool any = false;
    for (int m = 0; m < EM; ++m) any |= s_row[m] >= 0;
    if (!any) return;
    if (tid < EM) {                          // inputs of the embedding (also kept in global memory for the debug tools)
        const int r = s_row[tid];
        EmbedIn in;
        in.xa0 = 0.f; in.xa1 = 0.f; in.tok_row = 0; in.state_idx = 0; in.grid_row = 0; in.cat_idx = 0;
        if (r >= 0) {
            in = embed_inputs_row(a.s, r, *a.s.col + a.col_add);
            a.s.xa_raw[(size_t)r * 2] = in.xa0; a.s.xa_raw[(size_t)r * 2 + 1] = in.xa1;
            a.s.tok_row[r] = in.tok_row; a.s.state_idx[r] = in.state_idx; a.s.grid_row[r] = in.grid_row;
            a.s.cat_idx[r] = in.cat_idx;
        }
        s_in[tid] = in;
        sraw[tid * 4 + 0] = in.xa0; sraw[tid * 4 + 1] = in.xa1; sraw[tid * 4 + 2] = 0.f; sraw[tid * 4 + 3] = 0.f;
    }
    ws_init(wsm);
    if (warp == NWARP) {
        if (lane < WS_STAGES) {
            WSeg segs[WS_MAX_SEGS];
            int n = fourier_segs(a.fx, 2, segs);
            n += mlp3_segs(a.fusion, 128, segs + n);
            ws_produce(wsm, segs, n);
        }
        return;
    }
    WsCons ws(wsm);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = warp; m < EM; m += NWARP) {
        float4 c = z4, t = z4, s = z4, g = z4;
        if (s_row[m] >= 0) {
            c = ld4(a.cat_tab + (size_t)s_in[m].cat_idx * 128 + 4 * lane);
            t = ld4(a.tok_tab + (size_t)s_in[m].tok_row * 128 + 4 * lane);
            s = ld4(a.state_tab + (size_t)s_in[m].state_idx * 128 + 4 * lane);
            g = ld4(a.grid_tab + (size_t)s_in[m].grid_row * 128 + 4 * lane);
        }
        st4(sA + m * HLD + 4 * lane, c);
        st4(sX + m * XLD + 4 * lane, t);
        st4(sX + m * XLD + 256 + 4 * lane, s);
        st4(sX + m * XLD + 384 + 4 * lane, g);
    }
    bool single = s_row[0] >= 0;             // one appended row: GEMV path of the tile products
    for (int m = 1; m < EM; ++m) single = single && s_row[m] < 0;
    fourier_body<EM>(ws, a.fx, 2, sraw, sF, sH, sA, single);
    for (int m = warp; m < EM; m += NWARP) st4(sX + m * XLD + 128 + 4 * lane, ld4(sH + m * HLD + 4 * lane));
    csync();
    mlp3_body<EM>(ws, a.fusion, sX, XLD, 128, sH, sA, [&](int m, int n, float v) {
        const int r = s_row[m];
        if (r >= 0) {
            a.out[(size_t)r * 128 + n] = v;
            if (a.out2) a.out2[(size_t)r * 128 + n] = v;
            if (a.out3) a.out3[(size_t)r * 128 + n] = v;
        }
    }, single);
}

__global__ void k_next_iter(int *col, int *iter) { *col += 1; *iter += 1; }
__global__ void k_set_scalar(int *p, int v) { *p = v; }

// ---------------------------------------------------------------------------------------------------------------
// scene setup: expand the history columns into the [R][T] state arrays and seed the outputs (:1638-1657, 1721-1735)
// ---------------------------------------------------------------------------------------------------------------
struct SetupArgs {
    DecState s;
    const float *pos_hist, *head_hist;
    const int *state_hist, *token_hist, *grid_hist;
    const uint8_t *tsrc_hist, *interact_hist;
};
__global__ void k_setup_state(const SetupArgs a) {
    const DecState &s = a.s;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = s.n_scenes * s.cap, T = s.T;
    if (idx >= R * T) return;
    const int r = idx / T, c = idx % T;
    const bool act = (r % s.cap) < s.n_rows[r / s.cap];
    float px = 0.f, py = 0.f, hd = 0.f;
    int st = 0, tok = -1, g = -1;
    uint8_t ts = 1, in = 1;
    if (act && c < s.HC) {
        const size_t h = (size_t)r * s.HC + c;
        px = a.pos_hist[h * 2]; py = a.pos_hist[h * 2 + 1]; hd = a.head_hist[h];
        st = a.state_hist[h]; tok = a.token_hist[h]; g = a.grid_hist[h];
        ts = a.tsrc_hist[h]; in = a.interact_hist[h];
    }
    if (!act) { ts = 0; in = 0; }
    s.pos[(size_t)idx * 2] = px; s.pos[(size_t)idx * 2 + 1] = py;
    s.head[idx] = hd; s.state[idx] = st; s.token[idx] = tok; s.grid[idx] = g;
    s.tsrc[idx] = ts; s.interact[idx] = in;
    s.next_token[idx] = c < s.HC ? tok : -1;
    s.next_state[idx] = c < s.HC ? st : 0;
    if (c == 0) { s.t_cnt[r] = 0; s.m_cnt[r] = 0; s.a_cnt[r] = 0; s.a_start[r] = 0; }
}

// history part of pred_traj / pred_head rebuilt from the history tokens (agent_decoder.py:2311-2335); note that the
// reference rotates/translates every history token by the pose of column 0.
__global__ void k_history_traj(const DecState s, float *hist_traj, float *hist_head) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = s.n_scenes * s.cap, T = s.T;
    if (r >= R || (r % s.cap) >= s.n_rows[r / s.cap]) return;
    const float px = s.pos[(size_t)r * T * 2], py = s.pos[(size_t)r * T * 2 + 1];
    const float th = s.head[(size_t)r * T];
    const float c = cosf(th), sn = sinf(th);
    const int ty = min(s.type[r], 2);
    for (int hc = 0; hc < s.HC; ++hc) {
        int tk = s.next_token[(size_t)r * T + hc];
        if (tk < 0) tk = 0;
        const float *box = s.vocab + ((size_t)ty * s.V + tk) * 48;
        for (int k = 1; k < 6; ++k) {
            float wx[4], wy[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float bx = box[(k * 4 + q) * 2], by = box[(k * 4 + q) * 2 + 1];
                wx[q] = __fadd_rn(__fadd_rn(__fmul_rn(bx, c), __fmul_rn(by, -sn)), px);
                wy[q] = __fadd_rn(__fadd_rn(__fmul_rn(bx, sn), __fmul_rn(by, c)), py);
            }
            const size_t o = (size_t)r * (s.HC * 5) + hc * 5 + (k - 1);
            hist_traj[o * 2] = (((wx[0] + wx[1]) + wx[2]) + wx[3]) * 0.25f;
            hist_traj[o * 2 + 1] = (((wy[0] + wy[1]) + wy[2]) + wy[3]) * 0.25f;
            hist_head[o] = atan2f(__fsub_rn(wy[0], wy[3]), __fsub_rn(wx[0], wx[3]));
        }
    }
}

// categorical embedding rows: cat[r] = type_a_emb[type r] + shape_emb(shape r); cat[R] = type_a_emb[seed] + shape_emb(0.1)
__global__ void k_add_type_emb(float *cat, const float *type_emb, const int *type, int R, int seed_type) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (R + 1) * 128) return;
    const int r = idx >> 7, c = idx & 127;
    const int ty = r < R ? type[r] : seed_type;
    cat[idx] += type_emb[ty * 128 + c];
}
__global__ void k_fill_shape_rows(float *dst, const float *shape, int R) {   // [R+1][3], last row = 0.1
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (R + 1) * 3) return;
    dst[idx] = idx < R * 3 ? shape[idx] : 0.1f;
}

}  // namespace infgen
