// Per-scene preparation of the agent stream on the device:
//
//   k_tokenize_agents    `TokenProcessor._tokenize_agent`  reference infgen/datasets/preprocess.py:364-550
//                        (clean_heading :315-322, _extrapolate_agent_to_prev_token_step :324-343, _match_agent_token
//                        :552-660, cal_polygon_contour :24-55): one CTA per agent; the closed-loop match walks the 18
//                        token steps in order (each step's frame is the previous step's MATCHED box), every step an argmin
//                        over the 2048 vocabulary boxes of the summed corner distances.
//   k_fetch_enterings    `InfGen._fetch_enterings`          reference infgen/model/infgen.py:1008-1090: ego-centric grid
//                        token / offset of every agent and map token per column (Attr_Tokenizer.encode_pos,
//                        attr_tokenizer.py:77-89: argmin over 1961 cells, one warp per point), heading tokens, and the
//                        per-column order of the entering agents by bearing (`sort_indices`).
//
// Integer outputs are bit-exact against the oracle / the reference goldens as long as no argmin is decided by the last
// ulp (the arithmetic below mirrors the reference's operation order with explicit _rn intrinsics, no FMA contraction).
#pragma once
#include "common.cuh"

namespace infgen {

constexpr int PREP_NT = 256;
constexpr int PREP_MAX_STEPS = 512;           // raw steps per track held in shared memory

struct TokenizeArgs {
    int A, N, T;                              // agents, raw steps, token steps = N / 5
    int V;                                    // vocabulary size (2048)
    int predict_state;
    const unsigned char *valid;               // [A][N]
    const float *heading;                     // [A][N]
    const float *pos;                         // [A][N][3]
    const float *vel;                         // [A][N][2]
    const unsigned char *type;                // [A]
    const float *vocab;                       // [3][V][6][4][2]
    // outputs
    long long *token_idx, *state_idx;         // [A][T]
    float *contour;                           // [A][T][4][2]
    float *token_pos, *token_heading;         // [A][T][2], [A][T]
    unsigned char *raw_valid, *agent_valid;   // [A][T]
};

__global__ void __launch_bounds__(PREP_NT) k_tokenize_agents(const TokenizeArgs a) {
    __shared__ float s_h[PREP_MAX_STEPS], s_px[PREP_MAX_STEPS], s_py[PREP_MAX_STEPS];
    __shared__ unsigned char s_v[PREP_MAX_STEPS];
    __shared__ float s_red_d[PREP_NT / 32];
    __shared__ int s_red_i[PREP_NT / 32];
    __shared__ float s_con[8];
    __shared__ int s_best;
    __shared__ int s_tok[128], s_state[128];
    __shared__ unsigned char s_tv[128];
    const int ag = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N, T = a.T;
    for (int i = tid; i < N; i += PREP_NT) {
        s_h[i] = a.heading[(size_t)ag * N + i];
        s_px[i] = a.pos[((size_t)ag * N + i) * 3];
        s_py[i] = a.pos[((size_t)ag * N + i) * 3 + 1];
        s_v[i] = a.valid[(size_t)ag * N + i];
    }
    __syncthreads();
    const int ty = a.type[ag];
    // _get_agent_shape (:345-353): (width, length) by type
    const float width = ty == 0 ? 2.0f : 1.0f, length = ty == 0 ? 4.8f : (ty == 1 ? 2.0f : 1.0f);
    if (tid == 0) {
        // clean_heading (:315-322) - pairs of the ORIGINAL valid mask
        for (int i = 0; i + 1 < N; ++i) {
            const float d = fabsf(wrap_angle(__fsub_rn(s_h[i], s_h[i + 1])));
            if (d > 1.5f && s_v[i] && s_v[i + 1]) s_h[i + 1] = s_h[i];
        }
        // _extrapolate_agent_to_prev_token_step (:324-343); torch.max(valid, dim=1).indices = first True (0 if none)
        int t = 0;
        for (int i = 0; i < N; ++i) if (s_v[i]) { t = i; break; }
        int n = t % 5;
        if (t == 10 && !s_v[5]) n = 5;
        if (n > 0) {
            const float vx = a.vel[((size_t)ag * N + t) * 2], vy = a.vel[((size_t)ag * N + t) * 2 + 1];
            const float dx = __fmul_rn(vx, 0.1f), dy = __fmul_rn(vy, 0.1f);
            for (int j = 0; j < n; ++j) {
                s_v[t - j - 1] = 1;
                s_h[t - j - 1] = s_h[t];
                s_px[t - j - 1] = __fsub_rn(s_px[t - j], dx);
                s_py[t - j - 1] = __fsub_rn(s_py[t - j], dy);
            }
        }
    }
    __syncthreads();
    // ---- _match_agent_token (:552-660) ------------------------------------------------------------------------------------
    const float *voc = a.vocab + (size_t)ty * a.V * 48 + 40;      // last sub-step box of token k: voc + 48 k, [4][2]
    float prev_h = s_h[0], prev_x = s_px[0], prev_y = s_py[0];
    for (int c = 0; c < T; ++c) {
        const int i = 5 * (c + 1);
        const bool ok = s_v[i - 5] && s_v[i];
        const float cs = cosf(prev_h), sn = sinf(prev_h);
        // cal_polygon_contour (:24-55) of the raw box at step i
        const float hc = __fmul_rn(0.5f, cosf(s_h[i])), hs = __fmul_rn(0.5f, sinf(s_h[i]));
        const float lc = __fmul_rn(length, hc), ls = __fmul_rn(length, hs), wc = __fmul_rn(width, hc), ws = __fmul_rn(width, hs);
        const float x = s_px[i], y = s_py[i];
        const float cx[4] = {__fsub_rn(__fadd_rn(x, lc), ws), __fadd_rn(__fadd_rn(x, lc), ws), __fadd_rn(__fsub_rn(x, lc), ws),
                             __fsub_rn(__fsub_rn(x, lc), ws)};
        const float cy[4] = {__fadd_rn(__fadd_rn(y, ls), wc), __fsub_rn(__fadd_rn(y, ls), wc), __fsub_rn(__fsub_rn(y, ls), wc),
                             __fadd_rn(__fsub_rn(y, ls), wc)};
        float bd = INFINITY;
        int bi = 0x7fffffff;
        for (int k = tid; k < a.V; k += PREP_NT) {
            const float4 b0 = ldg4(voc + (size_t)k * 48), b1 = ldg4(voc + (size_t)k * 48 + 4);
            const float bx[4] = {b0.x, b0.z, b1.x, b1.z}, by[4] = {b0.y, b0.w, b1.y, b1.w};
            float d = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // row vector x [[c, s], [-s, c]] + prev_pos.  Rounding as torch's CPU bmm / norm produce it (K = 2 GEMM:
                // acc = a0 b0, acc = fma(a1, b1, acc); norm: sqrt(fma(y, y, x x)) - determined by experiment)
                const float wx = __fadd_rn(__fmaf_rn(by[q], -sn, __fmul_rn(bx[q], cs)), prev_x);
                const float wy = __fadd_rn(__fmaf_rn(by[q], cs, __fmul_rn(bx[q], sn)), prev_y);
                const float ex = __fsub_rn(wx, cx[q]), ey = __fsub_rn(wy, cy[q]);
                d = __fadd_rn(d, sqrtf(__fmaf_rn(ey, ey, __fmul_rn(ex, ex))));
            }
            if (d < bd) { bd = d; bi = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { s_red_d[warp] = bd; s_red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float d = s_red_d[0];
            int k = s_red_i[0];
            for (int w = 1; w < PREP_NT / 32; ++w)
                if (s_red_d[w] < d || (s_red_d[w] == d && s_red_i[w] < k)) { d = s_red_d[w]; k = s_red_i[w]; }
            s_best = k;
            const float *b = voc + (size_t)k * 48;
            for (int q = 0; q < 4; ++q) {
                s_con[2 * q] = __fadd_rn(__fmaf_rn(b[2 * q + 1], -sn, __fmul_rn(b[2 * q], cs)), prev_x);
                s_con[2 * q + 1] = __fadd_rn(__fmaf_rn(b[2 * q + 1], cs, __fmul_rn(b[2 * q], sn)), prev_y);
            }
            s_tok[c] = k;
        }
        __syncthreads();
        // matched box -> outputs; frame of the next step
        const float mx = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(s_con[0], s_con[2]), s_con[4]), s_con[6]), 0.25f);
        const float my = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(s_con[1], s_con[3]), s_con[5]), s_con[7]), 0.25f);
        const float mh = atan2f(__fsub_rn(s_con[1], s_con[7]), __fsub_rn(s_con[0], s_con[6]));
        if (tid < 8) a.contour[(((size_t)ag * T + c) * 4) * 2 + tid] = s_con[tid];
        if (tid == 0) {
            a.token_pos[((size_t)ag * T + c) * 2] = mx;
            a.token_pos[((size_t)ag * T + c) * 2 + 1] = my;
            a.token_heading[(size_t)ag * T + c] = mh;
            s_tv[c] = s_v[i - 5] && s_v[i];
        }
        prev_h = ok ? mh : s_h[i];
        prev_x = ok ? mx : s_px[i];
        prev_y = ok ? my : s_py[i];
        __syncthreads();
    }
    // ---- states (:438-459) -------------------------------------------------------------------------------------------------
    if (tid == 0) {
        int bos = 0, eos = T - 1;
        for (int c = 0; c < T; ++c) if (s_tv[c]) { bos = c; break; }
        for (int c = T - 1; c >= 0; --c) if (s_tv[c]) { eos = c; break; }
        for (int c = 0; c < T; ++c) {
            int st = 1;
            if (c == bos) st = 2;
            if (c == eos) st = 3;
            if (c < bos || c > eos) st = 0;
            s_state[c] = st;
        }
        if (s_state[T - 1] == 3) s_state[T - 1] = 1;
        for (int c = 0; c < T; ++c) {
            const int st = s_state[c];
            const size_t o = (size_t)ag * T + c;
            unsigned char tv = s_tv[c];
            if (st == 2) tv = 0;
            if (st == 0) { a.token_pos[o * 2] = 0.f; a.token_pos[o * 2 + 1] = 0.f; a.token_heading[o] = 0.f; }
            if (st == 2) { a.token_pos[o * 2] = s_px[5 * (c + 1)]; a.token_pos[o * 2 + 1] = s_py[5 * (c + 1)]; }
            a.token_idx[o] = st == 0 ? -1 : (st == 2 ? -2 : s_tok[c]);
            a.state_idx[o] = st;
            a.raw_valid[o] = tv;
            a.agent_valid[o] = a.predict_state ? 1 : tv;
        }
    }
}

struct EnterArgs {
    int A, T, P, G;
    int av;                                   // ego row
    float radius, angle_interval;
    const float *token_pos, *token_heading;   // [A][T][2], [A][T]
    const long long *state_idx;               // [A][T]
    const float *pt_pos;                      // [P][3]
    const float *cells;                       // [G][2]
    long long *grid_idx;                      // [A][T]
    float *grid_off, *pos_xy;                 // [A][T][2]
    long long *head_tok;                      // [A][T]
    float *head_theta;                        // [A][T]
    long long *sort_idx;                      // [A][T]
    unsigned char *inrange, *bos;             // [A][T]
    long long *pt_grid;                       // [T][P]
    float *bearing;                           // scratch [T][A]
};

// Attr_Tokenizer.encode_pos (attr_tokenizer.py:77-89) of one point by one warp: nearest cell (first minimum) and the
// rotated ego-centric coordinates
__device__ __forceinline__ int encode_pos_warp(const float *__restrict__ cells, int G, float px, float py, float ex, float ey,
                                               float eh, float &qx, float &qy) {
    const int lane = threadIdx.x & 31;
    const float th = -__fsub_rn(eh, 1.5707963267948966f);
    const float c = cosf(th), s = sinf(th);
    const float rx = __fsub_rn(px, ex), ry = __fsub_rn(py, ey);
    qx = __fadd_rn(__fmul_rn(rx, c), __fmul_rn(ry, -s));
    qy = __fadd_rn(__fmul_rn(rx, s), __fmul_rn(ry, c));
    float bd = INFINITY;
    int bi = 0x7fffffff;
    for (int g = lane; g < G; g += 32) {
        const float d = norm2(__fsub_rn(qx, cells[(size_t)g * 2]), __fsub_rn(qy, cells[(size_t)g * 2 + 1]));
        if (d < bd) { bd = d; bi = g; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    return bi;
}

// one warp per (column t, point): points 0..A-1 are the agents, A..A+P-1 the map tokens.  grid = (ceil((A+P)/8), T)
__global__ void __launch_bounds__(256) k_fetch_enterings(const EnterArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.y, T = a.T;
    const int p = blockIdx.x * 8 + warp;
    if (p >= a.A + a.P) return;
    const float ex = a.token_pos[((size_t)a.av * T + t) * 2], ey = a.token_pos[((size_t)a.av * T + t) * 2 + 1];
    const float eh = a.token_heading[(size_t)a.av * T + t];
    if (p < a.A) {
        const size_t o = (size_t)p * T + t;
        const float px = a.token_pos[o * 2], py = a.token_pos[o * 2 + 1], ph = a.token_heading[o];
        const int st = (int)a.state_idx[o];
        const float dx = __fsub_rn(px, ex), dy = __fsub_rn(py, ey);
        const bool inr = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) <= a.radius;
        float qx = 0.f, qy = 0.f;
        int g = -1;
        const bool m = st != 0 && inr;
        if (m) g = encode_pos_warp(a.cells, a.G, px, py, ex, ey, eh, qx, qy);
        if (lane == 0) {
            a.grid_idx[o] = g;
            a.grid_off[o * 2] = m ? __fsub_rn(qx, a.cells[(size_t)g * 2]) : 0.f;
            a.grid_off[o * 2 + 1] = m ? __fsub_rn(qy, a.cells[(size_t)g * 2 + 1]) : 0.f;
            a.pos_xy[o * 2] = m ? dx : 0.f;
            a.pos_xy[o * 2 + 1] = m ? dy : 0.f;
            a.inrange[o] = inr;
            a.bos[o] = st == 2;
            // heading token (attr_tokenizer.py:101-104) and wrapped relative heading
            const float rel = __fsub_rn(ph, eh);
            const float w = wrap_angle(rel);
            const float deg = __fmul_rn(__fdiv_rn(__fadd_rn(w, 3.14159265358979323846f), 6.28318530717958647692f), 360.0f);
            a.head_tok[o] = (long long)floorf(__fdiv_rn(deg, a.angle_interval));
            a.head_theta[o] = w;
            // bearing of the agent in the ego frame (sort key of the entering agents, infgen.py:1056-1062)
            const float b = angle_between(cosf(eh), sinf(eh), dx, dy);
            a.bearing[(size_t)t * a.A + p] = (st == 2 && inr) ? b : INFINITY;
        }
    } else {
        const int q = p - a.A;
        const float px = a.pt_pos[(size_t)q * 3], py = a.pt_pos[(size_t)q * 3 + 1];
        const float dx = __fsub_rn(px, ex), dy = __fsub_rn(py, ey);
        const bool inr = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) <= a.radius;
        float qx, qy;
        int g = -1;
        if (inr) g = encode_pos_warp(a.cells, a.G, px, py, ex, ey, eh, qx, qy);
        if (lane == 0) a.pt_grid[(size_t)t * a.P + q] = g;
    }
}

// sort_indices[:, t]: the entering in-range agents of column t by ascending bearing (ties: lower row first), every other
// slot holds the ego row (infgen.py:1060-1063).  grid = T, one thread per agent (rank by counting)
__global__ void k_sort_enterings(const EnterArgs a) {
    const int t = blockIdx.x;
    for (int i = threadIdx.x; i < a.A; i += blockDim.x) a.sort_idx[(size_t)i * a.T + t] = a.av;
    __syncthreads();
    for (int i = threadIdx.x; i < a.A; i += blockDim.x) {
        const float d = a.bearing[(size_t)t * a.A + i];
        if (isinf(d)) continue;
        int rank = 0;
        for (int j = 0; j < a.A; ++j) {
            const float e = a.bearing[(size_t)t * a.A + j];
            if (e < d || (e == d && j < i)) ++rank;
        }
        a.sort_idx[(size_t)rank * a.T + t] = i;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Map side of the preparation: `InfGen.match_token_map` (infgen/model/infgen.py:918-984).  One warp per 5 m map polyline
// (three points): the points are moved into the frame of the first one (:927-935), the lanes stride over the map
// vocabulary (three sample points per entry, shared memory) and keep the entry with the smallest summed squared distance
// (:936-937; ties to the lower index as torch.argmin); lane 0 also counts the polyline for its (polygon, side) row of
// the scene's [polygon, side, slot] mask (:955-971).
// ---------------------------------------------------------------------------------------------------------------
struct MapMatchArgs {
    int P, V;
    const float *traj_pos;         // [P][3][2]
    const float *traj_theta;       // [P]
    const int *pl_rank;            // [P] row of the owning polygon among the scene's sorted distinct polygons
    const unsigned char *side;     // [P] 0 left, 1 right, 2 centre
    const float *sample_pt;        // [V][3][2]
    long long *token_idx;          // [P]
    float *position;               // [P][3]
    float *orientation;            // [P]
    float *best;                   // [P] distance of the match (tests: margin of near ties)
    int *counts;                   // [polygons][3], zeroed
};
constexpr int MATCH_NT = 256;
__global__ void __launch_bounds__(MATCH_NT) k_match_map_tokens(const MapMatchArgs a) {
    extern __shared__ __align__(16) float s_tok[];                 // [V][6]
    for (int i = threadIdx.x; i < a.V * 6; i += MATCH_NT) s_tok[i] = a.sample_pt[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * (MATCH_NT / 32) + warp;
    if (t >= a.P) return;
    const float *tp = a.traj_pos + (size_t)t * 6;
    const float th = a.traj_theta[t];
    const float c = cosf(th), s = sinf(th);
    float lx[3], ly[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {                                   // [dx, dy] x [[c, -s], [s, c]]
        const float dx = __fsub_rn(tp[2 * j], tp[0]), dy = __fsub_rn(tp[2 * j + 1], tp[1]);
        lx[j] = __fmaf_rn(dy, s, __fmul_rn(dx, c));                 // torch's K = 2 bmm on the CPU: fma(a1, b1, a0 * b0)
        ly[j] = __fmaf_rn(dy, c, __fmul_rn(dx, -s));
    }
    float bd = INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < a.V; v += 32) {
        const float *q = s_tok + v * 6;
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float ex = __fsub_rn(q[2 * j], lx[j]), ey = __fsub_rn(q[2 * j + 1], ly[j]);
            d = __fadd_rn(d, __fmul_rn(ex, ex));
            d = __fadd_rn(d, __fmul_rn(ey, ey));
        }
        if (d < bd) { bd = d; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) {
        a.token_idx[t] = bi;
        a.position[(size_t)t * 3] = tp[0]; a.position[(size_t)t * 3 + 1] = tp[1]; a.position[(size_t)t * 3 + 2] = 0.f;
        a.orientation[t] = th;
        if (a.best) a.best[t] = bd;
        atomicAdd(a.counts + a.pl_rank[t] * 3 + min((int)a.side[t], 2), 1);
    }
}

}  // namespace infgen
