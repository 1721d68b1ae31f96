// Insertion stage of `InfGenAgentDecoder.inference` (agent_decoder.py:1744-2114): the kernels that are not
// AttentionLayer / FourierEmbedding / MLPLayer calls (those reuse k_layer / k_fourier / k_mlp_layer).
//
// Per decode iteration t > 0 and per pass p (up to insert_limit = 10, :1737-1738) a "seed" query row - a copy of the ego
// row at column cur (`_pad_feat`, :511-526) - attends to an occupancy node, to the map tokens within pl2seed_radius and
// to the agents within pl2seed_radius through 3 x {occ2sa, pt2sa, a2sa} layers (:1861-1871); heads on the result decide
// whether / where / what to insert (:1884-1911).  A new agent then gets its heading and a position offset from
// 3 x {pt2a, a2a} layers over its own 10 m neighbourhood (:2037-2074).
//
// KV-cache formulation (checked by parity, like the motion stage): every agent row runs through the seed stack and the
// heading stack WITHOUT edges in the reference (only the query row has incoming edges), so their K/V rows per layer
// depend only on the column's input feature and are computed once per iteration (and once more for each inserted row).
#pragma once
#include "common.cuh"
#include "ops.cuh"
#include "decode.cuh"

namespace infgen {

constexpr int SEED_MAP_MAX = 2048;     // max_num_neighbors of the pl2seed radius query (:1847)
constexpr int SEED_AGENT_MAX = 300;    // (:1838)
constexpr int NEW_MAP_MAX = 128;       // (:2035)
constexpr int NEW_AGENT_MAX = 24;      // (:2029)
constexpr int INSERT_LIMIT = 10;       // (:1738)
constexpr int SEED_SLOTS = INSERT_LIMIT + 1;
constexpr int SEED_ROW_STRIDE = 4;     // the query row of scene b is row 4b of its own row space: alone in its tile, so that
                                       // all warps of the CTA share its (up to 2048) edges (InsState::seed_stride = 4).
                                       // Batches of more than one wave of clusters pack the query rows instead
                                       // (seed_stride = 1: four scenes per tile, two warps per row): 32 scenes are 8 clusters
                                       // in one wave, not 32 clusters in three

struct InsState {
    int seed_stride;                   // rows between the query rows of consecutive scenes: SEED_ROW_STRIDE or 1
    int beam;                          // insert_beam_size
    int force_enter;                   // the reference's DEBUG=1 switch (:1888-1889)
    unsigned seed;
    float r_seed2, r_new_a2, r_new_m2; // squared radii: pl2seed (75 m), a2sa (10 m), pl2sa (10 m)
    float angle_interval;
    // per scene
    int *active;                       // still inserting in this iteration
    int *n_new;                        // agents inserted in this iteration
    int *pass;                         // passes done in this iteration
    int *new_row;                      // row (within the scene) appended by the last pass, -1 if none
    int *row_lo;                       // first row the "new rows only" launches process (= new_row or n_rows)
    int *flags;                        // [2]: any scene still active, any scene appended a row (loop control)
    int *stat;                         // [2] passes / heading stages executed by graph replays (launch accounting)
    int *done_ctr;                     // CTAs of k_seed_decide that have finished (the last one summarises the pass)
    int *ha_lo;                        // [ns] first row the heading-stack K/V pass of this iteration still has to cover:
                                       //      0 until the first heading stage of the iteration ran, then "none"
    int as_stride;                     // agent -> seed slots per scene: min(cap, SEED_AGENT_MAX)
    int *as_seen;                      // [ns] rows within the seed radius so far (the neighbour limit counts them all)
    int *as_new_list, *as_new_n;       // slots appended by the last heading stage (their relative embedding is due)
    int *new_list, *n_new_list;        // global rows appended by the last pass, compact (row-list launches of the heading stage)
    int *prev_list, *n_prev_list;      // the same for the pass before (rows whose heading-stack K|V rows are still due)
    // loop control without the host: condition handles of the CUDA-graph WHILE (another pass) / IF (a row was appended)
    // nodes, set from k_ins_begin / k_seed_decide when the iteration is replayed as a graph (use_cond)
    int use_cond;
    cudaGraphConditionalHandle h_pass, h_new;
    // seed query edges: one destination (the seed row) per scene
    int *ps_cnt, *ps_src; float *ps_raw;   // map -> seed   [ns], [ns*SEED_MAP_MAX], [..][3]
    int *as_cnt, *as_src; float *as_raw;   // agent -> seed [ns], [ns*as_stride], [..][3]
    int *one_cnt, *occ_src;                // occupancy node -> seed: always one edge, source = scene
    // new-agent edges: destination = the appended row; indexed by global row
    int *hp_cnt, *hp_start, *hp_src; float *hp_raw;   // map -> new agent   [R], [R], [ns*NEW_MAP_MAX], [..][3]
    int *ha_cnt, *ha_start, *ha_src; float *ha_raw;   // agent -> new agent [R], [R], [ns*NEW_AGENT_MAX], [..][3]
    int *hp_cnt_s, *ha_cnt_s;                          // [ns] the same counts per scene (slot validity of the embeddings)
    // occupancy
    float *occ;                        // [ns][G] 0/1 occupancy of the ego-centric grid at column cur
    float *occ_emb;                    // [ns][128] seed_agent_occ_embed(occ)
    float *kv_occ;                     // [3][ns][256] K|V of the occupancy node for the three occ2sa layers
    // query row
    float *x_seed;                     // [ns*SEED_ROW_STRIDE][128], row 4b
    const float *seed_feat;            // [128] `_build_agent_feature(..., state_index=invalid)` - a constant
    // head outputs of the query row
    float *pos_logits, *ag_occ_logits, *pt_occ_logits;   // [ns*SEED_ROW_STRIDE][G], row 4b
    float *small_logits;               // state [rows][2] | type [rows][3] | shape [rows][3], rows = ns * seed_stride
    // per row
    int *ins_col;                      // [R] column at which the row was inserted, -1 for the scene's own agents
    float *shape_rows;                 // [R+1][3] shape fed to shape_emb (row R = 0.1)
    int *pred_type;                    // [R]
    float *pred_shape;                 // [R][3]
    // records of the insertions, one per appended row and indexed by its global row (the dense [11][S](+[G]) tensors the
    // reference returns are all zero except for these; the host rebuilds them, see include/infgen_b200.h "wire format")
    int *rec_meta;                     // [R][2]: decode iteration, slot (1..10) within the iteration
    float *rec_state_prob;             // [R]
    float *rec_pos_prob, *rec_ag_occ, *rec_pt_occ, *rec_occ_gt;   // [R][G]
    float *rec_softmax;                // [ns][2] max / sum-exp of the position logits of the last query (k_seed_records)
    int *err;
};

// y[n] = b[n] + sum_k x[k] W[k][n], n < N, one row; x: shared [4*K4]; W packed [K4][ldn][4]
__device__ __forceinline__ void gemv_row(const float *xs, const float *__restrict__ wp, int ldn, int K4,
                                         const float *__restrict__ bias, int N, float *out) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc = 0.f;
#pragma unroll 8
        for (int k4 = 0; k4 < K4; ++k4) {
            const float4 w = ldg4(wp + ((size_t)k4 * ldn + n) * 4);
            const float4 x = ld4(xs + 4 * k4);
            acc = fmaf(x.x, w.x, acc); acc = fmaf(x.y, w.y, acc); acc = fmaf(x.z, w.z, acc); acc = fmaf(x.w, w.w, acc);
        }
        out[n] = acc + (bias ? __ldg(bias + n) : 0.f);
    }
}
// The per-scene kernels of the insertion stage are single-CTA latency chains: every GEMV below keeps all its weight loads
// in flight at once (a rolled loop of dependent-looking L2 loads costs one round trip per trip).
// y[n] = b[n] + sum_{k<128} x[k] W[k][n], n < N <= 128: all NT threads, the two k-halves of an output are added through
// `red` ([2][128] floats, shared).  Contains two __syncthreads().  `out` may be shared or global.
__device__ __forceinline__ void gemv128(const float *xs, const float *__restrict__ wp, int ldn,
                                        const float *__restrict__ bias, int N, float *out, float *red) {
    const int n = threadIdx.x & 127, kh = threadIdx.x >> 7;
    float acc = 0.f;
    if (n < N) {
        float4 w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = ldg4(wp + ((size_t)(16 * kh + i) * ldn + n) * 4);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 x = ld4(xs + 4 * (16 * kh + i));
            acc = fmaf(x.x, w[i].x, acc); acc = fmaf(x.y, w[i].y, acc); acc = fmaf(x.z, w[i].z, acc); acc = fmaf(x.w, w[i].w, acc);
        }
    }
    red[kh * 128 + n] = acc;
    __syncthreads();
    if (kh == 0 && n < N) out[n] = (red[n] + red[128 + n]) + (bias ? __ldg(bias + n) : 0.f);
    __syncthreads();
}
// y[n] = b[n] + sum_{k<128} x[k] W[k][n], n < 256 = NT: one thread per output, 32 loads in flight
__device__ __forceinline__ void gemv256(const float *xs, const float *__restrict__ wp, const float *__restrict__ bias,
                                        float *out) {
    const int n = threadIdx.x;
    float acc = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = ldg4(wp + ((size_t)(16 * h + i) * 256 + n) * 4);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 x = ld4(xs + 4 * (16 * h + i));
            acc = fmaf(x.x, w[i].x, acc); acc = fmaf(x.y, w[i].y, acc); acc = fmaf(x.z, w[i].z, acc); acc = fmaf(x.w, w[i].w, acc);
        }
    }
    out[n] = acc + __ldg(bias + n);
}
// in-place LayerNorm (+ReLU) of one 128-vector in shared memory by warp 0
__device__ __forceinline__ void ln_row(float *s, const float *g, const float *b, bool relu) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float4 v = ln128(ld4(s + 4 * lane), g, b, lane);
        if (relu) v = relu4(v);
        st4(s + 4 * lane, v);
    }
}
// MLPLayer (layers.py:206-215) on one row held in shared memory: out[n_out] (shared or global)
__device__ __forceinline__ void mlp_head_row(const float *sx, const MlpHeadW &w, float *sh, float *out, float *red) {
    gemv128(sx, w.w0, 128, w.b0, 128, sh, red);
    ln_row(sh, w.ln_g, w.ln_b, true);
    __syncthreads();
    gemv128(sh, w.w3, w.n_pad, w.b3, w.n_out, out, red);
}

// agent -> seed edge of row rj (scene b) towards the query row, which sits on the ego pose (px, py, hd): raw features
__device__ __forceinline__ void seed_edge_write(const DecState &s, const InsState &q, int slot, int rj, int col, float dx,
                                                float dy, float hd, float hx, float hy) {
    q.as_src[slot] = rj;
    q.as_raw[(size_t)slot * 3 + 0] = norm2(dx, dy);
    q.as_raw[(size_t)slot * 3 + 1] = angle_between(hx, hy, dx, dy);
    q.as_raw[(size_t)slot * 3 + 2] = wrap_angle(__fsub_rn(s.head[(size_t)rj * s.T + col], hd));
}

// ---------------------------------------------------------------------------------------------------------------
// start of an iteration's insertion stage: reset the per-scene flags; map -> seed edges (the seed sits on the ego's
// pose at column cur for every pass of the iteration).  One CTA per scene.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_ins_begin(const DecState s, const InsState q) {
    const int b = blockIdx.x, col = *s.col, T = s.T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        q.active[b] = 1; q.n_new[b] = 0; q.pass[b] = 0; q.new_row[b] = -1; q.row_lo[b] = s.n_rows[b];
        q.one_cnt[b] = 1; q.occ_src[b] = b; q.ha_lo[b] = 0;
        if (b == 0) {
            q.flags[0] = 1; q.flags[1] = 0; *q.done_ctr = 0;
            *q.n_new_list = 0; *q.n_prev_list = 0;
            if (q.use_cond) cudaGraphSetConditional(q.h_pass, 1u);
        }
    }
    for (int i = threadIdx.x; i < s.n_rows[b]; i += NT) s.hv_src[b * s.cap + i] = -1;   // vectors were rebuilt (:2265)
    const int re = b * s.cap + s.ego_row[b];
    const float px = s.pos[((size_t)re * T + col) * 2], py = s.pos[((size_t)re * T + col) * 2 + 1];
    const float hd = s.head[(size_t)re * T + col];
    const float hx = cosf(hd), hy = sinf(hd);
    const int pt0 = s.pt_ptr[b], pt1 = s.pt_ptr[b + 1];
    // every warp scans a contiguous eighth of the scene's map tokens; counts first, then an ordered write
    __shared__ int s_wcnt[NWARP];
    const int per = ((pt1 - pt0 + NWARP - 1) / NWARP + 31) & ~31;
    const int w0 = pt0 + warp * per, w1 = min(pt1, w0 + per);
    int mine = 0;
    for (int p0 = w0; p0 < w1; p0 += 32) {
        const int p = p0 + lane;
        bool ok = false;
        if (p < w1) {
            const float dx = __fsub_rn(s.pt_pos[(size_t)p * 2], px), dy = __fsub_rn(s.pt_pos[(size_t)p * 2 + 1], py);
            ok = dist2(dx, dy) < q.r_seed2;
        }
        mine += __popc(__ballot_sync(0xffffffffu, ok));
    }
    if (lane == 0) s_wcnt[warp] = mine;
    __syncthreads();
    int cnt = 0, total = 0;
    for (int w = 0; w < NWARP; ++w) { if (w < warp) cnt += s_wcnt[w]; total += s_wcnt[w]; }
    for (int p0 = w0; p0 < w1; p0 += 32) {
        const int p = p0 + lane;
        float dx = 0.f, dy = 0.f;
        bool ok = false;
        if (p < w1) {
            dx = __fsub_rn(s.pt_pos[(size_t)p * 2], px);
            dy = __fsub_rn(s.pt_pos[(size_t)p * 2 + 1], py);
            ok = dist2(dx, dy) < q.r_seed2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int rank = cnt + __popc(mask & lanemask_lt());
        if (ok && rank < SEED_MAP_MAX) {
            const int slot = b * SEED_MAP_MAX + rank;
            q.ps_src[slot] = p;
            q.ps_raw[(size_t)slot * 3 + 0] = norm2(dx, dy);
            q.ps_raw[(size_t)slot * 3 + 1] = angle_between(hx, hy, dx, dy);
            q.ps_raw[(size_t)slot * 3 + 2] = wrap_angle(__fsub_rn(s.pt_ori[p], hd));
        }
        cnt += __popc(mask);
    }
    if (threadIdx.x == 0) q.ps_cnt[b] = min(total, SEED_MAP_MAX);
    // ---- agent -> seed edges (:1833-1841): rows within the radius of the ego pose (first SEED_AGENT_MAX by index, the
    //      query row itself being the last index), kept if they interact at column cur.  Built once per iteration; a
    //      row appended by a pass adds its edge in k_head_finalize (it is the last index) ----
    if (warp != 0) return;
    const int n = s.n_rows[b], r0 = b * s.cap;
    int acnt = 0, seen = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane, rj = r0 + j;
        float dx = 0.f, dy = 0.f;
        bool within = false;
        if (j < n) {
            dx = __fsub_rn(s.pos[((size_t)rj * T + col) * 2], px);
            dy = __fsub_rn(s.pos[((size_t)rj * T + col) * 2 + 1], py);
            within = dist2(dx, dy) < q.r_seed2;
        }
        const unsigned wm = __ballot_sync(0xffffffffu, within);
        const bool in_first = within && (seen + __popc(wm & lanemask_lt())) < SEED_AGENT_MAX;
        const bool ok = in_first && s.interact[(size_t)rj * T + col] != 0;
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (ok) seed_edge_write(s, q, b * q.as_stride + acnt + __popc(mask & lanemask_lt()), rj, col, dx, dy, hd, hx, hy);
        acnt += __popc(mask);
        seen += __popc(wm);
    }
    if (lane == 0) { q.as_cnt[b] = acnt; q.as_seen[b] = seen; }
}

// ---------------------------------------------------------------------------------------------------------------
// per pass: occupancy of the grid at column cur -> seed_agent_occ_embed -> K|V of the three occ2sa layers (:1850-1859);
// the query row's input feature.  One CTA per scene.
// ---------------------------------------------------------------------------------------------------------------
struct SeedPrepArgs {
    DecState s;
    InsState q;
    MlpHeadW occ_embed;        // seed_agent_occ_embed: G -> 128 -> 128
    AttnW occ2sa[3];
};
__global__ void __launch_bounds__(NT) k_seed_prepare(const SeedPrepArgs a) {
    __shared__ __align__(16) float sh[128];
    __shared__ __align__(16) float se[128];
    const DecState &s = a.s;
    const InsState &q = a.q;
    const int b = blockIdx.x, col = *s.col, T = s.T, G = s.G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = s.n_rows[b], r0 = b * s.cap;
    if (!q.active[b]) return;
    // ---- occupancy (:1851-1853) ----
    float *occ = q.occ + (size_t)b * G;
    for (int g = tid; g < G; g += NT) occ[g] = 0.f;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const int g = s.grid[(size_t)(r0 + i) * T + col];
        if (g >= 0) occ[g] = 1.f;
    }
    __syncthreads();
    // ---- first Linear of seed_agent_occ_embed on a 0/1 vector: bias + the columns of the occupied cells ----
    __shared__ int s_cells[2048];                       // occupied cells, ascending (G = 1961 cells at most)
    __shared__ int s_ncell;
    if (warp == 0) {
        int cnt = 0;
        for (int g0 = 0; g0 < G; g0 += 32) {
            const int g = g0 + lane;
            const bool ok = g < G && occ[g] != 0.f;
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (ok) s_cells[cnt + __popc(mask & lanemask_lt())] = g;
            cnt += __popc(mask);
        }
        if (lane == 0) s_ncell = cnt;
    }
    __syncthreads();
    __shared__ float s_red2[256];
    {   // two halves of the block take alternate cells, four loads in flight each
        const int n = tid & 127, half = tid >> 7, nc = s_ncell;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float *w0 = a.occ_embed.w0;
        int k = half;
        for (; k + 6 < nc; k += 8) {
            const int g0 = s_cells[k], g1 = s_cells[k + 2], g2 = s_cells[k + 4], g3 = s_cells[k + 6];
            a0 += __ldg(w0 + ((size_t)(g0 >> 2) * 128 + n) * 4 + (g0 & 3));
            a1 += __ldg(w0 + ((size_t)(g1 >> 2) * 128 + n) * 4 + (g1 & 3));
            a2 += __ldg(w0 + ((size_t)(g2 >> 2) * 128 + n) * 4 + (g2 & 3));
            a3 += __ldg(w0 + ((size_t)(g3 >> 2) * 128 + n) * 4 + (g3 & 3));
        }
        for (; k < nc; k += 2) {
            const int g0 = s_cells[k];
            a0 += __ldg(w0 + ((size_t)(g0 >> 2) * 128 + n) * 4 + (g0 & 3));
        }
        s_red2[tid] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    if (tid < 128) sh[tid] = (s_red2[tid] + s_red2[128 + tid]) + __ldg(a.occ_embed.b0 + tid);
    __syncthreads();
    ln_row(sh, a.occ_embed.ln_g, a.occ_embed.ln_b, true);
    __syncthreads();
    gemv128(sh, a.occ_embed.w3, a.occ_embed.n_pad, a.occ_embed.b3, 128, se, s_red2);
    if (tid < 128) q.occ_emb[(size_t)b * 128 + tid] = se[tid];
    // K|V of the occupancy node for the three occ2sa layers (layers.py:65-71, 107-108): three LayerNorms by three warps,
    // then three independent 128 -> 256 GEMVs
    __shared__ __align__(16) float sn3[3][128];
    if (warp < 3) st4(sn3[warp] + 4 * lane, ln128(ld4(se + 4 * lane), a.occ2sa[warp].ln_src_g, a.occ2sa[warp].ln_src_b, lane));
    __syncthreads();
    for (int i = 0; i < 3; ++i)
        gemv256(sn3[i], a.occ2sa[i].w_kv, a.occ2sa[i].b_kv, q.kv_occ + ((size_t)i * gridDim.x + b) * 256);
    // ---- query row feature ----
    if (tid < 128) q.x_seed[(size_t)b * q.seed_stride * 128 + tid] = q.seed_feat[tid];
}

// ---------------------------------------------------------------------------------------------------------------
// per pass: small heads of the query row, the decision (:1884-1911) and, on 'enter', the new row (:1913-1995).
// The three grid-sized heads (position, agent / map occupancy) were computed by k_mlp_layer.  One CTA per scene.
// ---------------------------------------------------------------------------------------------------------------
struct SeedDecideArgs {
    DecState s;
    InsState q;
    long long *tstamp;         // optional [64] clock64 stamps of CTA 0 (debug: phase breakdown, INFGEN_TSTAMP=1)
};
// order-preserving map float -> unsigned (max of the keys = max of the floats)
__device__ __forceinline__ unsigned float_key(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void seed_decide_scene(const SeedDecideArgs &a) {
    __shared__ float s_small[8];                        // state[2] type[3] shape[3]
    __shared__ float s_red[2 * NWARP];
    __shared__ float s_wv[NWARP * INSERT_LIMIT];
    __shared__ int s_wi[NWARP * INSERT_LIMIT];
    __shared__ float s_topv[INSERT_LIMIT], s_p[INSERT_LIMIT], s_occ[INSERT_LIMIT];
    __shared__ int s_topi[INSERT_LIMIT];
    const DecState &s = a.s;
    const InsState &q = a.q;
    const int b = blockIdx.x, col = *s.col, t = *s.iter, T = s.T, G = s.G, S = s.S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ts_n = 0;
    auto stamp = [&]() { if (a.tstamp && b == 0 && tid == 0 && ts_n < 32) a.tstamp[ts_n++] = clock64(); };
    stamp();
    // scalars of the decision, fetched while the heads run (the decision itself is one thread's chain)
    const int rows0 = s.n_rows[b], pass0 = q.pass[b], nnew0 = q.n_new[b];
    if (tid == 0) { q.new_row[b] = -1; q.row_lo[b] = rows0; }            // nothing appended by this pass so far
    if (!q.active[b]) return;
    if (tid < NWARP * INSERT_LIMIT) { s_wv[tid] = -INFINITY; s_wi[tid] = 0x7fffffff; }
    // the three small heads of the query row (state, type, shape; computed by k_mlp_layer beside the grid-sized ones)
    if (tid < 8) {
        const int rows = (int)gridDim.x * q.seed_stride, r = b * q.seed_stride;
        const float *p = tid < 2 ? q.small_logits + (size_t)r * 2 + tid
                                 : (tid < 5 ? q.small_logits + (size_t)rows * 2 + (size_t)r * 3 + (tid - 2)
                                            : q.small_logits + (size_t)rows * 5 + (size_t)r * 3 + (tid - 5));
        s_small[tid] = *p;
    }
    // the grid logits of the position head: warp w owns a contiguous slice of the cells
    const float *lg = q.pos_logits + (size_t)b * q.seed_stride * G;
    const int per = (G + NWARP - 1) / NWARP, g0 = warp * per, g1 = min(G, g0 + per);
    constexpr int VPL = 8;                              // cells per lane (G <= 8 * 32 * NWARP)
    float v[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const int g = g0 + lane + 32 * j;
        v[j] = g < g1 ? lg[g] : -INFINITY;
    }
    stamp();
    // ---- position: softmax over the grid, top-k, draw (:1896-1902) ----
    float mxv = -INFINITY;
#pragma unroll
    for (int j = 0; j < VPL; ++j) mxv = fmaxf(mxv, v[j]);
    mxv = warp_max(mxv);
    if (lane == 0) s_red[warp] = mxv;
    __syncthreads();
    float gmax = s_red[0];
#pragma unroll
    for (int w = 1; w < NWARP; ++w) gmax = fmaxf(gmax, s_red[w]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) sum += (g0 + lane + 32 * j < g1) ? expf(v[j] - gmax) : 0.f;
    sum = warp_sum(sum);
    if (lane == 0) s_red[NWARP + warp] = sum;
    stamp();
    // the beam largest logits of the slice, ties to the lower index: the largest key by one warp reduction, then the lowest
    // cell that holds it by another
    for (int k = 0; k < q.beam; ++k) {
        float bv = v[0]; int bi = g0 + lane;
#pragma unroll
        for (int j = 1; j < VPL; ++j)
            if (v[j] > bv) { bv = v[j]; bi = g0 + lane + 32 * j; }
        const unsigned key = float_key(bv), kmax = __reduce_max_sync(0xffffffffu, key);
        const int win = __reduce_min_sync(0xffffffffu, key == kmax ? bi : 0x7fffffff);
        const int wl = (win - g0) & 31;                 // the lane that owns cell `win`
        const float wv = __shfl_sync(0xffffffffu, bv, wl);
#pragma unroll
        for (int j = 0; j < VPL; ++j)
            if (win == g0 + lane + 32 * j) v[j] = -INFINITY;
        if (lane == 0) { s_wv[warp * INSERT_LIMIT + k] = wv; s_wi[warp * INSERT_LIMIT + k] = win; }
    }
    __syncthreads();
    stamp();
    float den = 0.f;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) den += s_red[NWARP + w];
    // merge the NWARP x beam candidates: a candidate's rank is the number of candidates that beat it (larger logit, or the
    // same logit at a lower cell) - the order the reference's topk returns them in
    // (slots beyond the beam hold -inf / INT_MAX and never beat a candidate)
    if (tid < NWARP * INSERT_LIMIT && (tid % INSERT_LIMIT) < q.beam) {
        const float cv = s_wv[tid];
        const int ci = s_wi[tid];
        int rank = 0;
#pragma unroll 16
        for (int o = 0; o < NWARP * INSERT_LIMIT; ++o) {
            const float ov = s_wv[o];
            const int oi = s_wi[o];
            rank += (ov > cv || (ov == cv && oi < ci)) ? 1 : 0;
        }
        if (rank < q.beam) { s_topv[rank] = cv; s_topi[rank] = ci; }
    }
    __syncthreads();
    // probabilities and occupancy of the candidates, one lane each
    const float *occ = q.occ + (size_t)b * G;
    if (tid < q.beam) {
        s_p[tid] = expf(s_topv[tid] - gmax) / den;
        s_occ[tid] = occ[s_topi[tid]];
    }
    __syncthreads();
    stamp();
    __shared__ int s_cell, s_append, s_row;
    if (tid == 0) {
        float total = 0.f;
        if (q.beam > 1)
            for (int k = 0; k < q.beam; ++k) total += s_p[k];
        // state (:1884-1889): argmax of the 2-way softmax, index 1 = 'enter'
        int enter = s_small[1] > s_small[0] ? 1 : 0;
        if (q.force_enter) enter = 1;
        int pick = 0, append = 0, pass = pass0;
        for (;;) {
            if (q.beam > 1) {
                const float thr = uniform01(q.seed ^ 0x5EEDu, (unsigned)s.scene_id[b], (unsigned)pass, (unsigned)t) * total;
                float c = 0.f;
                pick = q.beam - 1;
                for (int k = 0; k < q.beam; ++k) { c += s_p[k]; if (thr < c) { pick = k; break; } }
            }
            ++pass;
            if (s_occ[pick] != 0.f) {                   // overlap filter (:1906-1909): retry
                // with a deterministic choice (beam 1) every retry repeats this pass: the reference spins until the limit
                if (q.beam == 1 || pass >= INSERT_LIMIT) { q.active[b] = 0; break; }
                // The reference restores its features and runs the whole query again (`feat_a = raw_feat_a.clone();
                // continue`): nothing the query reads has changed, so every logit comes out the same and only the draw
                // (keyed by the pass index) differs - the retry is taken here, without another pass of the stage
                continue;
            }
            if (!enter || nnew0 + 1 > INSERT_LIMIT) {
                q.active[b] = 0;
            } else if (rows0 >= s.cap) {
                *q.err = 2;                              // row capacity exhausted
                q.active[b] = 0;
            } else {
                append = 1;
                if (pass >= INSERT_LIMIT) q.active[b] = 0;
            }
            break;
        }
        q.pass[b] = pass;
        s_cell = s_topi[pick];
        s_append = append;
        s_row = rows0;
    }
    __syncthreads();
    stamp();
    if (!s_append) return;
    // ---- 1.5 append the new row (:1913-1995) ----
    const int i_new = s_row, r = b * s.cap + i_new, cell = s_cell;
    const int re = b * s.cap + s.ego_row[b];
    const float ex = s.pos[((size_t)re * T + col) * 2], ey = s.pos[((size_t)re * T + col) * 2 + 1];
    const float eh = s.head[(size_t)re * T + col];
    for (int c = tid; c < T; c += NT) {
        const size_t o = (size_t)r * T + c;
        if (c != col) {                                  // (column col is written below)
            s.pos[o * 2] = 0.f; s.pos[o * 2 + 1] = 0.f; s.head[o] = 0.f;
            s.state[o] = ST_INVALID; s.token[o] = -1; s.grid[o] = -1;
            s.next_state[o] = 0;
        }
        s.interact[o] = c >= col ? 1 : 0;
        s.tsrc[o] = c >= col ? 1 : 0;                    // temporal_mask all true, sources from the BOS column on (:547-552)
        s.next_token[o] = -1;
    }
    const int n_new = nnew0 + 1;
    for (int k = tid; k < 5 * S; k += NT) {
        const size_t o = (size_t)r * (5 * S) + k;
        const bool ph = t > 0 && k >= (t - 1) * 5 && k < t * 5;   // placeholders of the previous 0.5 s: written below
        if (!ph) { s.pred_traj[o * 2] = 0.f; s.pred_traj[o * 2 + 1] = 0.f; s.pred_head[o] = 0.f; s.pred_state[o] = 0.f; }
    }
    if (tid == 0) {
        // decode_pos (attr_tokenizer.py:91-99): grid cell in the ego frame -> world
        const float th = __fsub_rn(eh, 1.5707963267948966f);
        const float c = cosf(th), sn = sinf(th);
        const float gx = s.grid_cells[(size_t)cell * 2], gy = s.grid_cells[(size_t)cell * 2 + 1];
        const float nx = __fadd_rn(__fadd_rn(__fmul_rn(gx, c), __fmul_rn(gy, -sn)), ex);
        const float ny = __fadd_rn(__fadd_rn(__fmul_rn(gx, sn), __fmul_rn(gy, c)), ey);
        const size_t o = (size_t)r * T + col;
        s.pos[o * 2] = nx; s.pos[o * 2 + 1] = ny;
        s.head[o] = eh;                                  // dummy value until the heading stage (:1948)
        s.grid[o] = cell;
        s.state[o] = ST_ENTER;
        s.token[o] = -2;                                 // BOS embedding at the insertion column (:1976)
        s.next_state[o] = ST_ENTER;                      // (:2114)
        int ty = 0;                                      // type (:1892-1893): argmax of the 3-way softmax
        if (s_small[3] > s_small[2]) ty = 1;
        if (s_small[4] > s_small[2 + ty]) ty = 2;
        const_cast<int *>(s.type)[r] = ty;
        q.pred_type[r] = ty;
        for (int k = 0; k < 3; ++k) { q.pred_shape[(size_t)r * 3 + k] = s_small[5 + k]; q.shape_rows[(size_t)r * 3 + k] = s_small[5 + k]; }
        q.ins_col[r] = col;
        if (t > 0)
            for (int k = 0; k < 5; ++k) {                // placeholders of the previous 0.5 s (:1967-1970)
                const size_t p = (size_t)r * (5 * S) + (size_t)(t - 1) * 5 + k;
                s.pred_traj[p * 2] = nx; s.pred_traj[p * 2 + 1] = ny; s.pred_head[p] = eh; s.pred_state[p] = (float)ST_ENTER;
            }
        const_cast<int *>(s.n_rows)[b] = i_new + 1;
        q.n_new[b] = n_new;
        q.new_row[b] = i_new;
        q.row_lo[b] = i_new;
        q.flags[1] = 1;
        // P(enter) of this query (:2105)
        const float m2 = fmaxf(s_small[0], s_small[1]);
        const float e0 = expf(s_small[0] - m2), e1 = expf(s_small[1] - m2);
        q.rec_state_prob[r] = e1 / (e0 + e1);
        q.rec_meta[(size_t)r * 2] = t; q.rec_meta[(size_t)r * 2 + 1] = n_new;
        // the grid-sized records of this insertion are written by k_seed_records, off the pass's critical path
        q.rec_softmax[2 * b] = gmax; q.rec_softmax[2 * b + 1] = den;
    }
    stamp();
}

// grid-sized records of the insertions of the last pass (:2099-2104): softmax of the position logits, the two occupancy
// heads and the occupancy the query saw.  Runs at the start of the heading stage on a side stream, before the occupancy
// node of the next pass is rebuilt.  One CTA per scene.
__global__ void __launch_bounds__(NT) k_seed_records(const DecState s, const InsState q) {
    const int b = blockIdx.x, G = s.G;
    const int i_new = q.new_row[b];
    if (i_new < 0) return;
    const int r = b * s.cap + i_new;
    const float gmax = q.rec_softmax[2 * b], den = q.rec_softmax[2 * b + 1];
    const float *lg = q.pos_logits + (size_t)b * q.seed_stride * G;
    const size_t ob = (size_t)r * G, qb = (size_t)b * q.seed_stride * G;
    for (int g = threadIdx.x; g < G; g += NT) {
        q.rec_pos_prob[ob + g] = expf(lg[g] - gmax) / den;
        q.rec_ag_occ[ob + g] = q.ag_occ_logits[qb + g];
        q.rec_pt_occ[ob + g] = q.pt_occ_logits[qb + g];
        q.rec_occ_gt[ob + g] = q.occ[(size_t)b * G + g];
    }
}

// One CTA per scene; the last CTA to finish summarises the pass for whoever drives the loop: flags[] for the host-driven
// path, the condition values of the WHILE / IF graph nodes for the replayed one.
__global__ void __launch_bounds__(NT) k_seed_decide(const SeedDecideArgs a) {
    seed_decide_scene(a);
    const InsState &q = a.q;
    __shared__ int s_last;
    if (gridDim.x > 1) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(q.done_ctr, 1) == (int)gridDim.x - 1;
        }
        __syncthreads();
    } else if (threadIdx.x == 0) {
        s_last = 1;                                     // one scene: its own thread 0 wrote everything the summary reads
    }
    if (threadIdx.x == 0 && s_last) {
        __threadfence();
        int any = 0, any_new = 0;
        const int n_prev = *q.n_new_list;
        for (int i = 0; i < n_prev; ++i) q.prev_list[i] = q.new_list[i];
        *q.n_prev_list = n_prev;
        for (int b = 0; b < (int)gridDim.x; ++b) {
            any |= ((volatile int *)q.active)[b] != 0;
            const int nr = ((volatile int *)q.new_row)[b];
            if (nr >= 0) q.new_list[any_new++] = b * a.s.cap + nr;
        }
        *q.n_new_list = any_new;
        any_new = any_new > 0;
        q.flags[0] = any; q.flags[1] = any_new;
        *q.done_ctr = 0;
        if (q.use_cond) {
            cudaGraphSetConditional(q.h_pass, (unsigned)any);
            cudaGraphSetConditional(q.h_new, (unsigned)any_new);
            q.stat[0] += 1; q.stat[1] += any_new;
        }
        if (a.tstamp) a.tstamp[40] = clock64();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// heading stage, edges of the appended rows (:2024-2035): agents within a2sa_radius (first NEW_AGENT_MAX by index)
// that interact at column cur, map tokens within pl2sa_radius (first NEW_MAP_MAX).  One CTA (two warps) per scene.
// ---------------------------------------------------------------------------------------------------------------
constexpr int NEW_EDGE_NT = 1024;      // warp 0: agents; warps 1..31: contiguous slices of the scene's map tokens
__global__ void __launch_bounds__(NEW_EDGE_NT) k_new_edges(const DecState s, const InsState q) {
    __shared__ int s_cnt[NEW_EDGE_NT / 32];
    const int b = blockIdx.x, col = *s.col, T = s.T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i_new = q.new_row[b];
    if (b == 0 && threadIdx.x == 0) *q.as_new_n = 0;
    if (i_new < 0) {
        if (threadIdx.x == 0) { q.hp_cnt_s[b] = 0; q.ha_cnt_s[b] = 0; }
        return;
    }
    const int r0 = b * s.cap, r = r0 + i_new, n = s.n_rows[b];
    const float px = s.pos[((size_t)r * T + col) * 2], py = s.pos[((size_t)r * T + col) * 2 + 1];
    const float hd = s.head[(size_t)r * T + col];
    const float hx = cosf(hd), hy = sinf(hd);
    // map tokens: the first NEW_MAP_MAX within the radius by ascending index.  A single warp walking the (up to 2048) tokens
    // 32 at a time is a chain of 64 dependent global round trips (41 us measured); seven warps count the hits of their
    // slices first, then write them behind the slices before them.
    constexpr int NMW = NEW_EDGE_NT / 32 - 1;
    const int pt0 = s.pt_ptr[b], pt1 = s.pt_ptr[b + 1];
    const int per = (((pt1 - pt0 + NMW - 1) / NMW) + 31) & ~31;
    const int p_lo = pt0 + (warp - 1) * per, p_hi = min(pt1, p_lo + per);
    if (warp > 0) {
        int cnt = 0;
        for (int p0 = p_lo; p0 < p_hi; p0 += 32) {
            const int p = p0 + lane;
            bool ok = false;
            if (p < p_hi) {
                const float dx = __fsub_rn(__ldg(s.pt_pos + (size_t)p * 2), px), dy = __fsub_rn(__ldg(s.pt_pos + (size_t)p * 2 + 1), py);
                ok = dist2(dx, dy) < q.r_new_m2;
            }
            cnt += __popc(__ballot_sync(0xffffffffu, ok));
        }
        if (lane == 0) s_cnt[warp] = cnt;
    } else {
        int cnt = 0, seen = 0;
        const int base = b * NEW_AGENT_MAX;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane, rj = r0 + j;
            float dx = 0.f, dy = 0.f;
            bool within = false;
            if (j < n) {
                dx = __fsub_rn(s.pos[((size_t)rj * T + col) * 2], px);
                dy = __fsub_rn(s.pos[((size_t)rj * T + col) * 2 + 1], py);
                within = dist2(dx, dy) < q.r_new_a2;
            }
            const unsigned wm = __ballot_sync(0xffffffffu, within);
            const bool in_first = within && (seen + __popc(wm & lanemask_lt())) < NEW_AGENT_MAX;
            const bool ok = in_first && j != i_new && s.interact[(size_t)rj * T + col] != 0;
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int slot = base + cnt + __popc(mask & lanemask_lt());
                q.ha_src[slot] = rj;
                q.ha_raw[(size_t)slot * 3 + 0] = norm2(dx, dy);
                q.ha_raw[(size_t)slot * 3 + 1] = angle_between(hx, hy, dx, dy);
                q.ha_raw[(size_t)slot * 3 + 2] = wrap_angle(__fsub_rn(s.head[(size_t)rj * T + col], hd));
            }
            cnt += __popc(mask);
            seen += __popc(wm);
        }
        if (lane == 0) { q.ha_cnt[r] = cnt; q.ha_start[r] = base; q.ha_cnt_s[b] = cnt; }
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            int tot = 0;
            for (int w = 1; w <= NMW; ++w) tot += s_cnt[w];
            tot = min(tot, NEW_MAP_MAX);
            q.hp_cnt[r] = tot; q.hp_start[r] = b * NEW_MAP_MAX; q.hp_cnt_s[b] = tot;
        }
        return;
    }
    int cnt = 0;                                      // hits of the slices before this warp's
    for (int w = 1; w < warp; ++w) cnt += s_cnt[w];
    const int base = b * NEW_MAP_MAX;
    for (int p0 = p_lo; p0 < p_hi && cnt < NEW_MAP_MAX; p0 += 32) {
        const int p = p0 + lane;
        float dx = 0.f, dy = 0.f;
        bool ok = false;
        if (p < p_hi) {
            dx = __fsub_rn(__ldg(s.pt_pos + (size_t)p * 2), px);
            dy = __fsub_rn(__ldg(s.pt_pos + (size_t)p * 2 + 1), py);
            ok = dist2(dx, dy) < q.r_new_m2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int rank = cnt + __popc(mask & lanemask_lt());
        if (ok && rank < NEW_MAP_MAX) {
            const int slot = base + rank;
            q.hp_src[slot] = p;
            q.hp_raw[(size_t)slot * 3 + 0] = norm2(dx, dy);
            q.hp_raw[(size_t)slot * 3 + 1] = angle_between(hx, hy, dx, dy);
            q.hp_raw[(size_t)slot * 3 + 2] = wrap_angle(__fsub_rn(__ldg(s.pt_ori + p), hd));
        }
        cnt += __popc(mask);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// heading stage, heads of the appended row and its final pose (:2060-2074).  One CTA per scene.
// ---------------------------------------------------------------------------------------------------------------
struct HeadFinalArgs {
    DecState s;
    InsState q;
    const float *x;            // [R][128] output of the heading stack
    MlpHeadW h_heading, h_offset;
};
__global__ void __launch_bounds__(NT) k_head_finalize(const HeadFinalArgs a) {
    __shared__ __align__(16) float sx[128];
    __shared__ __align__(16) float sh[128];
    __shared__ float s_out[128];
    const DecState &s = a.s;
    const InsState &q = a.q;
    const int b = blockIdx.x, col = *s.col, T = s.T;
    const int tid = threadIdx.x;
    const int i_new = q.new_row[b];
    if (i_new < 0) return;
    const int r = b * s.cap + i_new, re = b * s.cap + s.ego_row[b];
    if (tid < 128) sx[tid] = a.x[(size_t)r * 128 + tid];
    __syncthreads();
    __shared__ float s_red[256];
    mlp_head_row(sx, a.h_heading, sh, s_out, s_red);    // 120 heading bins
    __shared__ int s_bin;
    if (tid < 32) {                                     // argmax, first occurrence
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int k = tid; k < a.h_heading.n_out; k += 32)
            if (s_out[k] > bv) { bv = s_out[k]; bi = k; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (tid == 0) s_bin = bi;
    }
    __syncthreads();
    mlp_head_row(sx, a.h_offset, sh, s_out, s_red);     // 2 offsets
    if (tid == 0) {
        const float eh = s.head[(size_t)re * T + col];
        // decode_heading (attr_tokenizer.py:106-110) + wrap (:2063)
        float ang = __fsub_rn(__fmul_rn((float)s_bin, q.angle_interval), 180.f);
        ang = __fmul_rn(__fdiv_rn(ang, 360.f), 6.283185307179586f);
        const size_t o = (size_t)r * T + col;
        s.head[o] = wrap_angle(__fadd_rn(ang, eh));
        s.pos[o * 2] = __fadd_rn(s.pos[o * 2], __fmul_rn(tanhf(s_out[0]), 2.f));
        s.pos[o * 2 + 1] = __fadd_rn(s.pos[o * 2 + 1], __fmul_rn(tanhf(s_out[1]), 2.f));
        // sic (:2083): the heading vectors of EVERY agent inserted in this iteration become the newest one's
        for (int k = 0; k < q.n_new[b]; ++k) s.hv_src[b * s.cap + i_new - k] = r;
        // the new row's edge towards the query row of the following passes (it is the last index of the scene)
        const float ex = s.pos[((size_t)re * T + col) * 2], ey = s.pos[((size_t)re * T + col) * 2 + 1];
        const float dx = __fsub_rn(s.pos[o * 2], ex), dy = __fsub_rn(s.pos[o * 2 + 1], ey);
        if (dist2(dx, dy) < q.r_seed2) {
            const int rank = q.as_seen[b];
            q.as_seen[b] = rank + 1;
            if (rank < SEED_AGENT_MAX && s.interact[o] != 0) {
                const int slot = b * q.as_stride + q.as_cnt[b];
                seed_edge_write(s, q, slot, r, col, dx, dy, eh, cosf(eh), sinf(eh));
                q.as_cnt[b] += 1;
                q.as_new_list[atomicAdd(q.as_new_n, 1)] = slot;
            }
        }
    }
}

// cat[r] += type_a_emb[type[r]] for the rows [row_lo[b], n_rows[b]) (their shape embedding was just written)
// (one block per scene)
__global__ void k_add_type_emb_rows(const DecState s, const int *row_lo, float *cat, const float *type_emb) {
    const int b = blockIdx.x;
    for (int i = row_lo[b]; i < s.n_rows[b]; ++i) {
        const int r = b * s.cap + i;
        cat[(size_t)r * 128 + threadIdx.x] += type_emb[s.type[r] * 128 + threadIdx.x];
    }
}

// copy rows [row_lo[b], n_rows[b]) of every scene from one [R][128] buffer to another
__global__ void k_copy_new_rows(const DecState s, const int *row_lo, const float *src, float *dst) {
    const int r = blockIdx.x, b = r / s.cap, i = r - b * s.cap;
    if ((row_lo && i < row_lo[b]) || i >= s.n_rows[b]) return;
    dst[(size_t)r * 128 + threadIdx.x] = src[(size_t)r * 128 + threadIdx.x];
}
// the same for the rows appended by the last pass, into two destinations: one block per scene
__global__ void k_copy_new_rows2(const DecState s, const int *row_lo, const float *src, float *dst0, float *dst1) {
    const int b = blockIdx.x;
    for (int i = row_lo ? row_lo[b] : 0; i < s.n_rows[b]; ++i) {
        const size_t o = (size_t)(b * s.cap + i) * 128 + threadIdx.x;
        const float v = src[o];
        dst0[o] = v; dst1[o] = v;
    }
}

}  // namespace infgen
