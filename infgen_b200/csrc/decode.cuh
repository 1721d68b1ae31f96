// Decode-loop kernels: everything of `InfGenAgentDecoder.inference` (agent_decoder.py:1740-2301) that is not an
// nn.Module call - edge construction, sampling, token->pose advance, grid tokenisation, next-column inputs.
#pragma once
#include "common.cuh"
#include "ops.cuh"

namespace infgen {

constexpr int ST_INVALID = 0, ST_VALID = 1, ST_ENTER = 2, ST_EXIT = 3;
constexpr int MAX_CAP = 256;       // rows per scene supported by k_edge_build's shared bitmaps
constexpr int RING = 16;           // temporal K/V ring depth (>= window + 1, power of two)

struct DecState {
    int n_scenes, cap, T, S, HC, W;
    int q_rows;                    // num_seed_feature: tail rows without temporal edges (agent_decoder.py:553-556)
    int G, V;                      // grid cells, motion-token vocabulary
    int max_m;                     // max_pl2a_neighbors
    float r_m2, r_a2;              // squared radii
    int use_state_token, disable_insertion, beam;
    unsigned seed;
    const int *n_rows, *ego_row, *scene_id;
    int *col, *iter;               // device scalars: current column / iteration
    float *pos, *head;             // [R][T][2], [R][T]
    int *state, *token, *grid;     // [R][T]
    uint8_t *interact, *tsrc;      // [R][T]
    const int *type;               // [R]
    const int *pt_ptr;
    const float *pt_pos, *pt_ori;
    const float *grid_cells;       // [G][2]
    const float *vocab;            // [3][V][6][4][2]
    // edges
    int *t_cnt, *t_src; float *t_raw;                    // [R], [R*W], [R*W][4]
    int *m_cnt, *m_src; float *m_raw;                    // [R], [R*max_m], [R*max_m][3]
    int *a_cnt, *a_start, *a_total, *a_src; float *a_raw; // [R], [R], [n_scenes], [n_scenes*cap*cap], [..][3]
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
// :2119-2121).  One CTA per scene.  Semantics of the third-party calls (oracle/shims): radius = strict `<`, the
// first max_num_neighbors sources by ascending index; edges ordered by destination then source.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_edge_build(const DecState s) {
    __shared__ int s_acnt[MAX_CAP];
    __shared__ int s_astart[MAX_CAP];
    __shared__ unsigned s_amask[MAX_CAP][MAX_CAP / 32];
    const int b = blockIdx.x, n = s.n_rows[b], col = *s.col, T = s.T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = b * s.cap;
    const int pt0 = s.pt_ptr[b], pt1 = s.pt_ptr[b + 1];
    const int nwords = (n + 31) / 32;
    for (int i = warp; i < n; i += NWARP) {
        const int r = r0 + i;
        const float px = s.pos[((size_t)r * T + col) * 2], py = s.pos[((size_t)r * T + col) * 2 + 1];
        const float hd = s.head[(size_t)r * T + col];
        const float hx = cosf(hd), hy = sinf(hd);
        const bool inv_d = s.state[(size_t)r * T + col] == ST_INVALID;
        const bool inter = s.interact[(size_t)r * T + col] != 0;
        // ---- temporal: (r, c) -> (r, col), 0 < col - c <= W --------------------------------------------------
        {
            int cnt = 0;
            if (i < n - s.q_rows) {
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
        }
        // ---- map -> agent: first max_m tokens within the radius ---------------------------------------------
        {
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
        }
        // ---- agent <-> agent, pass 1: neighbour bitmap ------------------------------------------------------
        {
            int cnt = 0;
            for (int w = 0; w < nwords; ++w) {
                const int j = w * 32 + lane;
                bool ok = false;
                if (inter && j < n && j != i) {
                    const int rj = r0 + j;
                    if (s.interact[(size_t)rj * T + col]) {
                        const float dx = __fsub_rn(px, s.pos[((size_t)rj * T + col) * 2]);
                        const float dy = __fsub_rn(py, s.pos[((size_t)rj * T + col) * 2 + 1]);
                        ok = dist2(dx, dy) < s.r_a2;
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) s_amask[i][w] = mask;
                cnt += __popc(mask);
            }
            if (lane == 0) s_acnt[i] = cnt;
        }
    }
    __syncthreads();
    if (warp == 0) {                                   // exclusive scan of the per-row a2a degrees
        int carry = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const int v = i < n ? s_acnt[i] : 0;
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (i < n) s_astart[i] = carry + x - v;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) s.a_total[b] = carry;
    }
    __syncthreads();
    for (int i = warp; i < n; i += NWARP) {
        const int r = r0 + i;
        const int base = b * s.cap * s.cap + s_astart[i];
        if (lane == 0) { s.a_start[r] = base; s.a_cnt[r] = s_acnt[i]; }
        if (s_acnt[i] == 0) continue;
        const float px = s.pos[((size_t)r * T + col) * 2], py = s.pos[((size_t)r * T + col) * 2 + 1];
        const float hd = s.head[(size_t)r * T + col];
        const float hx = cosf(hd), hy = sinf(hd);
        const bool inv_d = s.state[(size_t)r * T + col] == ST_INVALID;
        int run = 0;
        for (int w = 0; w < nwords; ++w) {
            const unsigned mask = s_amask[i][w];
            if (mask & (1u << lane)) {
                const int rj = r0 + w * 32 + lane;
                const int slot = base + run + __popc(mask & lanemask_lt());
                const bool inv_s = s.state[(size_t)rj * T + col] == ST_INVALID;
                float rx = __fsub_rn(s.pos[((size_t)rj * T + col) * 2], px);
                float ry = __fsub_rn(s.pos[((size_t)rj * T + col) * 2 + 1], py);
                float rh = wrap_angle(__fsub_rn(s.head[(size_t)rj * T + col], hd));
                if (inv_s && !inv_d) { rx = -1.f; ry = -1.f; rh = -1.f; }               // :647-653
                if (!inv_s && inv_d) { rx = 1.f; ry = 1.f; }
                if (inv_s && inv_d) { rx = -2.f; ry = -2.f; rh = -2.f; }
                s.a_src[slot] = rj;
                s.a_raw[(size_t)slot * 3 + 0] = norm2(rx, ry);
                s.a_raw[(size_t)slot * 3 + 1] = angle_between(hx, hy, rx, ry);
                s.a_raw[(size_t)slot * 3 + 2] = rh;
            }
            run += __popc(mask);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// inputs of the column embedding (agent_decoder.py:426-447 _build_vector_a, :449-509 / :2265-2287)
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_embed_inputs(const DecState s, int col_add) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = s.n_scenes * s.cap;
    if (r >= R || (r % s.cap) >= s.n_rows[r / s.cap]) return;
    const int col = *s.col + col_add, T = s.T;
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
    s.xa_raw[(size_t)r * 2] = norm2(mx, my);
    s.xa_raw[(size_t)r * 2 + 1] = angle_between(cosf(hd), sinf(hd), mx, my);
    const int tok = s.token[(size_t)r * T + col];
    s.tok_row[r] = s.type[r] * (s.V + 2) + (tok < 0 ? s.V + 2 + tok : tok);   // [-2] = BOS row, [-1] = no-token row
    s.state_idx[r] = st;
    const int g = s.grid[(size_t)r * T + col];
    s.grid_row[r] = g < 0 ? s.G : g;                                            // [-1] = invalid-offset row
    // row R = seed type + 0.1 shape.  The reference builds the categorical embeddings once, while every future
    // column is still 'invalid' (agent_decoder.py:1653-1657, 458-470), and later only rewrites them for steps that
    // turn invalid (:2235-2239): every generated column therefore carries the seed/0.1 row, valid or not.
    s.cat_idx[r] = (inv || col >= s.HC) ? R : r;
}

// counter-based uniform in [0,1): mirrors oracle.agent_decoder_oracle.uniform01
__device__ __forceinline__ float uniform01(unsigned seed, unsigned scene, unsigned row, unsigned it) {
    unsigned x = seed * 0x9E3779B1u + scene * 0x85EBCA77u + row * 0xC2B2AE3Du + it * 0x27D4EB2Fu + 0x165667B1u;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------------------------------
// sampling + state update + token->pose advance + grid token (agent_decoder.py:2160-2262). One CTA per scene.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_advance(const DecState s) {
    const int b = blockIdx.x, n = s.n_rows[b], col = *s.col, t = *s.iter, T = s.T, nxt = col + 1;
    const int r0 = b * s.cap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ego = s.ego_row[b];
    for (int i = threadIdx.x; i < n; i += NT) {
        const int r = r0 + i;
        // ---- merge the per-slice candidates: global max, softmax denominator, top-KTOP ----------------------
        float gmax = -INFINITY;
        for (int k = 0; k < NSLICE; ++k) gmax = fmaxf(gmax, s.part_m[(size_t)r * NSLICE + k]);
        float den = 0.f;
        for (int k = 0; k < NSLICE; ++k)
            den += s.part_s[(size_t)r * NSLICE + k] * expf(s.part_m[(size_t)r * NSLICE + k] - gmax);
        float cv[KTOP]; int ci[KTOP];
        unsigned long long taken = 0ull;
        for (int k = 0; k < KTOP; ++k) {
            float bv = -INFINITY; int bi = 0x7fffffff, bj = -1;
            for (int j = 0; j < NSLICE * KTOP; ++j) {
                if (taken >> j & 1ull) continue;
                const float v = s.part_v[(size_t)r * NSLICE * KTOP + j];
                const int id = s.part_i[(size_t)r * NSLICE * KTOP + j];
                if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; bj = j; }
            }
            taken |= 1ull << bj;
            cv[k] = bv; ci[k] = bi;
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
        // ---- state (:2166-2173) -------------------------------------------------------------------------------
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
            if (i == ego) st = ST_VALID;
            if (!s.use_state_token && st == ST_EXIT) st = ST_VALID;
            if (s.disable_insertion) st = ST_VALID;
        }
        if (s.forced_state) st = s.forced_state[(size_t)r * s.S + t];
        // ---- token -> pose (:2175-2211) -----------------------------------------------------------------------
        const int ty = min(s.type[r], 2);
        const int tk = tok < 0 ? tok + s.V : tok;
        const float *box = s.vocab + ((size_t)ty * s.V + tk) * 48;
        const float px = s.pos[((size_t)r * T + col) * 2], py = s.pos[((size_t)r * T + col) * 2 + 1];
        const float th = s.head[(size_t)r * T + col];
        const float c = cosf(th), sn = sinf(th);
        float lx = 0.f, ly = 0.f, lh = 0.f;
        for (int k = 1; k < 6; ++k) {
            float wx[4], wy[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float bx = box[(k * 4 + q) * 2], by = box[(k * 4 + q) * 2 + 1];
                wx[q] = __fadd_rn(__fadd_rn(__fmul_rn(bx, c), __fmul_rn(by, -sn)), px);
                wy[q] = __fadd_rn(__fadd_rn(__fmul_rn(bx, sn), __fmul_rn(by, c)), py);
            }
            const float mx = (((wx[0] + wx[1]) + wx[2]) + wx[3]) * 0.25f;
            const float my = (((wy[0] + wy[1]) + wy[2]) + wy[3]) * 0.25f;
            const float hh = atan2f(__fsub_rn(wy[0], wy[3]), __fsub_rn(wx[0], wx[3]));
            const size_t o = (size_t)r * (5 * s.S) + t * 5 + (k - 1);
            s.pred_traj[o * 2] = mx; s.pred_traj[o * 2 + 1] = my;
            s.pred_head[o] = hh;
            s.pred_state[o] = (float)st;
            lx = mx; ly = my; lh = hh;
        }
        const bool inv = st == ST_INVALID;
        if (inv) { tok = -1; lx = 0.f; ly = 0.f; lh = 0.f; }                              // :2221-2239
        s.pos[((size_t)r * T + nxt) * 2] = lx; s.pos[((size_t)r * T + nxt) * 2 + 1] = ly;
        s.head[(size_t)r * T + nxt] = lh;
        s.state[(size_t)r * T + nxt] = st;
        s.token[(size_t)r * T + nxt] = tok;
        s.interact[(size_t)r * T + nxt] = inv ? 0 : 1;
        s.next_token[(size_t)r * T + nxt] = tok;
        s.next_state[(size_t)r * T + nxt] = st;
        if (inv) s.grid[(size_t)r * T + nxt] = -1;
    }
    __syncthreads();
    // ---- ego-centric grid token of the new position (attr_tokenizer.py:77-89, agent_decoder.py:2214) -----------
    const int re = r0 + ego;
    const float ex = s.pos[((size_t)re * T + nxt) * 2], ey = s.pos[((size_t)re * T + nxt) * 2 + 1];
    const float eth = -__fsub_rn(s.head[(size_t)re * T + nxt], 1.5707963267948966f);
    const float ec = cosf(eth), es = sinf(eth);
    for (int i = warp; i < n; i += NWARP) {
        const int r = r0 + i;
        if (s.state[(size_t)r * T + nxt] == ST_INVALID) continue;
        const float rx = __fsub_rn(s.pos[((size_t)r * T + nxt) * 2], ex);
        const float ry = __fsub_rn(s.pos[((size_t)r * T + nxt) * 2 + 1], ey);
        const float qx = __fadd_rn(__fmul_rn(rx, ec), __fmul_rn(ry, -es));
        const float qy = __fadd_rn(__fmul_rn(rx, es), __fmul_rn(ry, ec));
        float bd = INFINITY; int bi = 0x7fffffff;
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
