// Cluster-cooperative AttentionLayer kernel (reference layers.py:61-113) for sm_100a.
//
// One thread-block cluster of CL = 8 CTAs owns a tile of M destination rows and runs whole AttentionLayers on it:
// edge attention (layers.py:78-92) -> gated update / to_out / LayerNorm / FFN (layers.py:74-75, 94-99) -> the
// LayerNorm + q/s/k/v projections of the NEXT layer (layers.py:65-71, 106-108), for up to two layers per launch.
//
//   * CTA `c` of the cluster is attention head `c` AND column slice `c` of every projection: it owns output columns
//     [16c, 16c+16) of the 128-wide Linears and [64c, 64c+64) of the FFN hidden layer, so it streams only 1/8 of a
//     layer's weights.  Those slices are stored contiguously per (layer, CTA) ("cluster-sliced" chunks, built once in
//     engine.cu) and arrive in shared memory through cp.async.bulk (TMA bulk copy) + mbarrier while the attention
//     phase runs; the next chunk is requested as soon as its buffer is free.
//   * each projection's output slice is written straight into the shared memory of all 8 CTAs (DSMEM) and published
//     with one cluster barrier, so every CTA holds the full rows again for the next LayerNorm / Linear.
//   * attention is a single pass over a row's edges with an online softmax (running max, rescaled sums); the
//     relative-embedding fold of ops.cuh (per-node relative queries, per-head aggregated rhat) is kept, so an edge
//     costs one 16-dot (k), one 128-dot (rhat) and their weighted sums for the head of this CTA.
//
// Numerics: fp32 FFMA everywhere; softmax denominator + 1e-16 as torch_geometric.utils.softmax.
#pragma once
#include "common.cuh"
#include "ops.cuh"
#include <cooperative_groups.h>

namespace infgen {
namespace cg = cooperative_groups;

constexpr int CL = 8;                       // CTAs per cluster == NHEAD

// ---- cluster-sliced weight chunks (float offsets inside one (layer, CTA) chunk) ------------------------------------
namespace cs_post {
constexpr int WVR = 0;                      // [32][16][4]   to_v_r columns 16c..
constexpr int WG = WVR + 2048;              // [64][16][4]   to_g
constexpr int WO = WG + 4096;               // [32][16][4]   to_out
constexpr int W1 = WO + 2048;               // [32][64][4]   ff_mlp.0 columns 64c..
constexpr int W2 = W1 + 8192;               // [128][16][4]  ff_mlp.3
constexpr int BVR = W2 + 8192;              // [16]
constexpr int BG = BVR + 16;
constexpr int BO = BG + 16;
constexpr int B1 = BO + 16;                 // [64]
constexpr int B2 = B1 + 64;                 // [16]
constexpr int LN_DST_G = B2 + 16;           // ten full LayerNorm vectors
constexpr int LN_DST_B = LN_DST_G + 128;
constexpr int LN_R_G = LN_DST_B + 128;
constexpr int LN_R_B = LN_R_G + 128;
constexpr int LN_POST_G = LN_R_B + 128;
constexpr int LN_POST_B = LN_POST_G + 128;
constexpr int LN_FFPRE_G = LN_POST_B + 128;
constexpr int LN_FFPRE_B = LN_FFPRE_G + 128;
constexpr int LN_FFPOST_G = LN_FFPRE_B + 128;
constexpr int LN_FFPOST_B = LN_FFPOST_G + 128;
constexpr int FLOATS = LN_FFPOST_B + 128;   // 25984 floats = 103,936 B
}  // namespace cs_post
namespace cs_pre {
constexpr int WQS = 0;                      // [32][32][4]   n < 16: to_q column 16c+n, else to_s column 16c+n-16
constexpr int WKV = WQS + 4096;             // [32][32][4]   to_k | to_v likewise
constexpr int WKR = WKV + 4096;             // [16][128]     to_k_r rows 16c..16c+16 (row = output channel)
constexpr int BQS = WKR + 2048;             // [32]
constexpr int BKV = BQS + 32;               // [32]
constexpr int LN_DST_G = BKV + 32;
constexpr int LN_DST_B = LN_DST_G + 128;
constexpr int LN_R_G = LN_DST_B + 128;
constexpr int FLOATS = LN_R_G + 128;        // 10688 floats = 42,752 B
}  // namespace cs_pre
static_assert(cs_post::FLOATS % 4 == 0 && cs_pre::FLOATS % 4 == 0, "bulk copies need 16-byte multiples");

struct PreArgs {
    const float *w;            // [CL][cs_pre::FLOATS] chunks of the layer whose inputs are projected; NULL = no pre
    int pre_kv;                // also project k|v of these rows (non-bipartite layers)
    float *kv_out;             // K|V rows of 256 floats
    int kv_ring;               // 1: row r -> slot r*RING + (col & (RING-1)); 0: slot r
    int col_add;
    int to_global;             // store q / s / qr to global memory (consumed by a later launch)
};
struct SubArgs {
    const float *w;            // [CL][cs_post::FLOATS] chunks of this layer
    int has_attn;              // 0: rows receive no edges (history prefill), agg = 0
    int has_pos;
    const float *kv;           // K|V rows of 256 floats
    const int *cnt;            // [R] edges of row r
    const int *start;          // [R] first edge slot (NULL: r * stride)
    int stride;
    const int *src;            // [slots] K/V row of the source
    const float *rhat;         // [slots][128]
    PreArgs pre;               // projections of the following layer
    float *trace_out;          // optional copy of the layer output [R][128]
    int row_shift;             // edge lists are indexed by (row >> row_shift) (query rows that sit alone in their tile)
    int wide;                  // 1: only the first row of each tile has edges and all warps of the CTA share them
    int elist;                 // which of the (up to 3) distinct edge lists of the launch this layer uses
    int grid_sync;             // wait for every CTA of the grid before the attention (K/V written by other clusters in
                               // this launch); only legal when the whole grid is co-resident
};
constexpr int MAX_SUB = 18;
struct LayerArgs {
    RowSpace rows;
    float *x;                  // [R][128] residual stream, updated in place
    float *q, *s, *qr;         // [R][128], [R][128], [R][8][128] hand-over between launches
    const int *col_ptr;        // device: current column (temporal ring slot)
    int ring;                  // ring depth
    PreArgs pre0;              // optional projections run before the first layer (else q/s/qr come from global)
    int n_sub;
    SubArgs sub[MAX_SUB];
    unsigned *grid_bar;        // zeroed counter for the grid barriers (grid_sync)
    // "ride-along" row of a seed-query launch (one scene per tile, M = rows.cap): tile position 1 of scene b carries row
    // ride_cap * b + ride_row[b] of x2 (ride_row[b] >= 0 and the scene's query row active) through the same layers WITHOUT
    // edges; the K|V projections of a `pre` with pre_kv are then stored for that row alone (slot = its row in x2).  The
    // row appended by the previous insertion pass becomes a source of this pass's agent -> seed attention that way,
    // without a chain of edge-less layers of its own in front of the query.
    const float *x2;
    const int *ride_row;
    int ride_cap;
    int no_store;              // 1: x is read only (launches that exist for the K|V rows of their `pre` projections)
    long long *tstamp;         // optional [256] clock64 stamps of CTA 0 (debug: phase breakdown)
};

// leading dimensions of the activation tiles: 4 (mod 32) floats so that the row lanes of slice_gemm hit distinct banks
constexpr int LD1 = 132;       // 128-wide tiles
constexpr int LD2 = 260;       // [agg | x_dst]
constexpr int LD5 = 516;       // FFN hidden
constexpr int RED_FLOATS = 16 * (NT + 16);   // k-split partials: up to 16 registers x (NT + pad) floats

template <int M>
struct LayerSmem {
    static constexpr int WPOST = 0;
    static constexpr int WPRE = WPOST + cs_post::FLOATS;
    static constexpr int X = WPRE + cs_pre::FLOATS;       // [M][LD1] residual
    static constexpr int CAT = X + M * LD1;               // [M][LD2] agg | LN_dst(x)
    static constexpr int U = CAT + M * LD2;               // [M][LD1]
    static constexpr int O = U + M * LD1;                 // [M][LD1]
    static constexpr int H = O + M * LD1;                 // [M][LD5]
    static constexpr int Y = H + M * LD5;                 // [M][LD1]
    static constexpr int RED = Y + M * LD1;               // k-split partials
    static constexpr int RAGG = RED + RED_FLOATS;         // [M][LD1] own head
    static constexpr int QR = RAGG + M * LD1;             // [M][128] own head
    static constexpr int Q = QR + M * 128;                // [M][16]
    static constexpr int S = Q + M * 16;                  // [M][16]
    static constexpr int AGG = S + M * 16;                // [M][16]
    static constexpr int SAL = AGG + M * 16;              // [M] (padded to 16)
    static constexpr int MERGE = SAL + 16;                // [NWARP][160]
    static constexpr int MBAR = MERGE + NWARP * 160;      // 2 x uint64 (weight chunks) + 5 x uint64 (exchanges)
    static constexpr int TOTAL = MBAR + 16;
    static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
    static_assert(BYTES <= 227 * 1024, "shared memory budget");
};

// DSMEM exchange without a cluster-wide barrier: a remote store that signals the DESTINATION CTA's mbarrier with its byte
// count (st.async ... mbarrier::complete_tx::bytes); the consumer posts the expected bytes and waits on its own barrier.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t raddr, float v, uint32_t rmbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(raddr), "f"(v), "r"(rmbar)
                 : "memory");
}

__device__ __forceinline__ void st_async_v4(uint32_t raddr, const float4 v, uint32_t rmbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rmbar)
                 : "memory");
}

// LayerNorm of a 128-vector (4 channels per lane) with the affine vectors in shared (or any generic) memory
__device__ __forceinline__ float4 ln128s(const float4 v, const float *g, const float *b, int lane) {
    float mean, rstd;
    ln_stats(v, mean, rstd);
    const float4 gg = ld4(g + 4 * lane), bb = ld4(b + 4 * lane);
    return make_float4((v.x - mean) * rstd * gg.x + bb.x, (v.y - mean) * rstd * gg.y + bb.y,
                       (v.z - mean) * rstd * gg.z + bb.z, (v.w - mean) * rstd * gg.w + bb.w);
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// ---------------------------------------------------------------------------------------------------------------------
// Column-slice GEMM:  Y[m][n] = sum_k X[m][k] W[k][n],  m < M, n < NL, k < 4*K4.
//   X  shared, row-major, leading dimension ldx (4 mod 32);  W shared, [K4][NL][4].
//   A thread owns a 4 x TN register tile (rows 4*rg + i, columns cg + NCG*j) for one of KS k-slices - shared-memory
//   bandwidth, not FFMA issue, bounds these tiny GEMMs, and the tile cuts the loads per FMA to (4 + TN) / (16 TN) -
//   and the KS partial tiles are reduced through `red` ([register][thread] layout, conflict-free both ways).
//   epi(m, n, value) runs once per output.  Contains one __syncthreads().
// ---------------------------------------------------------------------------------------------------------------------
template <int M, int NL, int TN, typename Epi>
__device__ __forceinline__ void slice_gemm(const float *xs, int ldx, const float *w, int K4, float *red, Epi epi) {
    constexpr int NCG = NL / TN, RG = M / 4, KS = NT / (NCG * RG);
    constexpr int RSTR = NT + NCG;                         // floats per register row of the partial buffer
    static_assert(NCG * RG * KS == NT && M % 4 == 0, "thread mapping");
    static_assert(4 * TN * RSTR <= RED_FLOATS, "reduction scratch too small");
    const int tid = threadIdx.x, cg = tid % NCG, rg = (tid / NCG) % RG, ks = tid / (NCG * RG);
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const float *xr = xs + (size_t)(4 * rg) * ldx;
#pragma unroll 2
    for (int k4 = ks; k4 < K4; k4 += KS) {
        float4 wv[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) wv[j] = ld4(w + (k4 * NL + cg + NCG * j) * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 x = ld4(xr + i * ldx + 4 * k4);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                acc[i][j] = fmaf(x.x, wv[j].x, acc[i][j]);
                acc[i][j] = fmaf(x.y, wv[j].y, acc[i][j]);
                acc[i][j] = fmaf(x.z, wv[j].z, acc[i][j]);
                acc[i][j] = fmaf(x.w, wv[j].w, acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) red[(i * TN + j) * RSTR + tid] = acc[i][j];
    __syncthreads();
    for (int o = tid; o < M * NL; o += NT) {
        const int m = o / NL, n = o % NL;
        const float *p = red + ((m & 3) * TN + n / NCG) * RSTR + (n % NCG) + NCG * (m >> 2);
        float v = 0.f;
#pragma unroll 8
        for (int k = 0; k < KS; ++k) v += p[k * NCG * RG];
        epi(m, n, v);
    }
}
// The same product for outputs that leave the CTA: fin(m, n, value) finishes one output, the four outputs of columns
// 4j .. 4j+3 of a row are then gathered in the lane of column 4j (three shuffles) and handed to send4(m, 4j, float4) - one
// 16-byte remote store per peer instead of four 4-byte ones (the receiving mbarrier counts one transaction per store).
template <int M, int NL, int TN, typename Fin, typename Send4>
__device__ __forceinline__ void slice_gemm_x(const float *xs, int ldx, const float *w, int K4, float *red, Fin fin, Send4 send4) {
    static_assert((M * NL) % 32 == 0 && NL % 4 == 0, "whole warps in the output loop, whole quads per row");
    const int lane = threadIdx.x & 31;
    slice_gemm<M, NL, TN>(xs, ldx, w, K4, red, [&](int m, int n, float v) {
        v = fin(m, n, v);
        const int b = lane & ~3;
        const float4 q = make_float4(__shfl_sync(0xffffffffu, v, b), __shfl_sync(0xffffffffu, v, b + 1),
                                     __shfl_sync(0xffffffffu, v, b + 2), __shfl_sync(0xffffffffu, v, b + 3));
        if ((lane & 3) == 0) send4(m, n, q);
    });
}

// ---------------------------------------------------------------------------------------------------------------------
// edge attention of head `c` for the M rows of the cluster (layers.py:78-92): online softmax, NWARP / M warps per row,
// chunks of 8 edges, two chunks in flight (the loads of chunk i+1 are issued before chunk i is reduced).
//   lane l loads float4 #l of every rhat row of the chunk (a 512-byte coalesced row per edge) and, for ONE edge of the
//   chunk (l >> 2), quarter (l & 3) of the 16-wide K and V head slices, so a chunk costs 40 registers.
// ---------------------------------------------------------------------------------------------------------------------
struct AttnPre {               // per warp: its share of the row's edges, source rows of the first 64 of them
    int e0, eb0, eb1;          // first slot of the row; [eb0, eb1) = this warp's edges (relative to e0)
    int src0, src1;
};
template <int M>
__device__ __forceinline__ AttnPre attn_prefetch(const SubArgs &A, const int *s_rid) {
    const int nparts = A.wide ? NWARP : NWARP / M;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = A.wide ? 0 : warp / nparts, part = A.wide ? warp : warp % nparts;
    const int r = s_rid[m];                              // global row, -1 = inactive
    AttnPre p;
    int n = 0;
    p.e0 = 0;
    if (A.has_attn && r >= 0) {
        const int ri = r >> A.row_shift;
        n = A.cnt[ri];
        p.e0 = A.start ? A.start[ri] : ri * A.stride;
    }
    const int share = nparts == 1 ? n : (((n + nparts - 1) / nparts + 7) & ~7);
    p.eb0 = min(n, part * share);
    p.eb1 = min(n, p.eb0 + share);
    p.src0 = (p.eb0 + lane < p.eb1) ? A.src[p.e0 + p.eb0 + lane] : 0;
    p.src1 = (p.eb0 + 32 + lane < p.eb1) ? A.src[p.e0 + p.eb0 + 32 + lane] : 0;
    return p;
}

struct AttnChunk {
    float4 rh[8];
    float4 k, v;
};

template <int M>
__device__ __forceinline__ void attn_phase(const SubArgs &A, const AttnPre &P, int c, const float *sq, const float *sqr,
                                           float *sagg, float *sragg, float *ssal, float *smerge) {
    const int nparts = A.wide ? NWARP : NWARP / M;      // warps sharing one row's edges
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = A.wide ? 0 : warp / nparts, part = A.wide ? warp : warp % nparts;
    const int eq = lane >> 2, qd = lane & 3;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 qr4 = ld4(sqr + m * 128 + 4 * lane);
    const float4 q4 = ld4(sq + m * 16 + 4 * qd);
    float mx = -INFINITY, den = 0.f;
    float4 ra = z4, av = z4;
    const float *kvb = A.kv + 16 * c + 4 * qd;
    const float *rhb = A.rhat + (size_t)P.e0 * 128 + 4 * lane;

    for (int sb = P.eb0; sb < P.eb1; sb += 64) {          // super-block of 64 edges: source rows held in two registers
        int src0 = P.src0, src1 = P.src1;
        if (sb != P.eb0) {
            src0 = (sb + lane < P.eb1) ? A.src[P.e0 + sb + lane] : 0;
            src1 = (sb + 32 + lane < P.eb1) ? A.src[P.e0 + sb + 32 + lane] : 0;
        }
        const int send = min(P.eb1, sb + 64);
        const int nch = (send - sb + 7) >> 3;
        auto load = [&](AttnChunk &ck, int ch) {
            const int eb = sb + 8 * ch;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                ck.rh[j] = (A.has_pos && eb + j < send) ? ld4(rhb + (size_t)(eb + j) * 128) : z4;
            const int idx = 8 * ch + eq;
            const int s0 = __shfl_sync(0xffffffffu, src0, idx & 31), s1 = __shfl_sync(0xffffffffu, src1, idx & 31);
            const size_t sj = (size_t)(idx < 32 ? s0 : s1);
            ck.k = z4; ck.v = z4;
            if (eb + eq < send) {
                ck.k = __ldcg(reinterpret_cast<const float4 *>(kvb + sj * 256));
                ck.v = __ldcg(reinterpret_cast<const float4 *>(kvb + sj * 256 + 128));
            }
        };
        auto compute = [&](const AttnChunk &ck, int ch) {
            const int eb = sb + 8 * ch;
            float pk = dot4(q4, ck.k);
            pk += __shfl_xor_sync(0xffffffffu, pk, 1);
            pk += __shfl_xor_sync(0xffffffffu, pk, 2);
            // 8 rhat dot products, 32 partial sums each: butterfly reduce-scatter (9 shuffles) that leaves the score of
            // edge l >> 2 in lane l - the lane group that also holds that edge's K/V quarter
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = dot4(qr4, ck.rh[j]);
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float recv = __shfl_xor_sync(0xffffffffu, b4 ? v[i] : v[i + 4], 16);
                v[i] = (b4 ? v[i + 4] : v[i]) + recv;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float recv = __shfl_xor_sync(0xffffffffu, b3 ? v[i] : v[i + 2], 8);
                v[i] = (b3 ? v[i + 2] : v[i]) + recv;
            }
            {
                const float recv = __shfl_xor_sync(0xffffffffu, b2 ? v[0] : v[1], 4);
                v[0] = (b2 ? v[1] : v[0]) + recv;
            }
            float pr = v[0];
            pr += __shfl_xor_sync(0xffffffffu, pr, 1);
            pr += __shfl_xor_sync(0xffffffffu, pr, 2);
            float p = (pr + pk) * 0.25f;                      // head_dim ** -0.5
            if (eb + eq >= send) p = -INFINITY;
            float pm = p;
            pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, 4));
            pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, 8));
            pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, 16));
            const float mn = fmaxf(mx, pm);                   // finite: edge eb exists
            const float sc = expf(mx - mn);                   // 0 on the first chunk
            const float wmine = expf(p - mn);                 // 0 for padded edges
            den = fmaf(den, sc, wmine);                       // per lane group; folded over the groups after the loop
            ra.x *= sc; ra.y *= sc; ra.z *= sc; ra.w *= sc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float w = __shfl_sync(0xffffffffu, wmine, 4 * j);
                ra.x = fmaf(w, ck.rh[j].x, ra.x); ra.y = fmaf(w, ck.rh[j].y, ra.y);
                ra.z = fmaf(w, ck.rh[j].z, ra.z); ra.w = fmaf(w, ck.rh[j].w, ra.w);
            }
            av.x = fmaf(av.x, sc, wmine * ck.v.x); av.y = fmaf(av.y, sc, wmine * ck.v.y);
            av.z = fmaf(av.z, sc, wmine * ck.v.z); av.w = fmaf(av.w, sc, wmine * ck.v.w);
            mx = mn;
        };
        AttnChunk ca, cb;
        if (nch > 0) load(ca, 0);
        for (int ch = 0; ch < nch; ch += 2) {
            if (ch + 1 < nch) load(cb, ch + 1);
            compute(ca, ch);
            if (ch + 2 < nch) load(ca, ch + 2);
            if (ch + 1 < nch) compute(cb, ch + 1);
        }
    }
    // V partial sums and the softmax denominator live per edge group: fold the 8 groups (lanes with equal l & 3)
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        den += __shfl_xor_sync(0xffffffffu, den, o);
        av.x += __shfl_xor_sync(0xffffffffu, av.x, o); av.y += __shfl_xor_sync(0xffffffffu, av.y, o);
        av.z += __shfl_xor_sync(0xffffffffu, av.z, o); av.w += __shfl_xor_sync(0xffffffffu, av.w, o);
    }
    if (nparts > 1) {                                    // fold the partial softmax states of the row's warps
        float *mg = smerge + warp * 160;
        if (part != 0) {
            if (lane == 0) { mg[0] = mx; mg[1] = den; }
            if (lane < 4) st4(mg + 4 + 4 * lane, av);
            st4(mg + 32 + 4 * lane, ra);
        }
        __syncthreads();
        if (part == 0) {
            for (int k = 1; k < nparts; ++k) {
                const float *og = smerge + (warp + k) * 160;
                const float mx1 = og[0], den1 = og[1];
                const float mn = fmaxf(mx, mx1);
                if (mn > -INFINITY) {
                    const float f0 = expf(mx - mn), f1 = expf(mx1 - mn);
                    const float4 ra1 = ld4(og + 32 + 4 * lane);
                    const float4 av1 = ld4(og + 4 + 4 * qd);
                    den = den * f0 + den1 * f1;
                    ra = make_float4(ra.x * f0 + ra1.x * f1, ra.y * f0 + ra1.y * f1, ra.z * f0 + ra1.z * f1,
                                     ra.w * f0 + ra1.w * f1);
                    av = make_float4(av.x * f0 + av1.x * f1, av.y * f0 + av1.y * f1, av.z * f0 + av1.z * f1,
                                     av.w * f0 + av1.w * f1);
                    mx = mn;
                }
            }
        }
    }
    if (A.wide && warp > 0 && warp < M) {                // the other rows of a query tile have no edges
        if (lane < 4) st4(sagg + warp * 16 + 4 * lane, z4);
        st4(sragg + warp * LD1 + 4 * lane, z4);
        if (lane == 0) ssal[warp] = 0.f;
    }
    if (part == 0) {
        const float inv = 1.0f / (den + 1e-16f);          // torch_geometric.utils.softmax denominator
        if (lane < 4) st4(sagg + m * 16 + 4 * lane, make_float4(av.x * inv, av.y * inv, av.z * inv, av.w * inv));
        st4(sragg + m * LD1 + 4 * lane, make_float4(ra.x * inv, ra.y * inv, ra.z * inv, ra.w * inv));
        if (lane == 0) ssal[m] = den * inv;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------------
template <int M>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1) k_layer(const LayerArgs a) {
    extern __shared__ __align__(16) float smem[];
    using L = LayerSmem<M>;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();                       // head / column slice of this CTA
    const int tile = (int)(blockIdx.x / CL);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ int s_rid[8];                                       // global row of every tile position, -1 = inactive
    __shared__ int s_alt[8];                                       // ride-along: row of x2 carried by the position, else -1
    if (tid < M) {
        const int r = a.rows.tile_row(tile, tid, M);
        int alt = -1;
        if (a.ride_row && tid == 1 && a.rows.active_row(r - 1)) {
            const int nr = a.ride_row[tile];
            if (nr >= 0) alt = tile * a.ride_cap + nr;
        }
        s_alt[tid] = alt;
        s_rid[tid] = (alt >= 0 || a.rows.active_row(r)) ? r : -1;
    }
    __syncthreads();
    unsigned act_mask = 0;                                         // bit m: tile position m holds an active row
#pragma unroll
    for (int m = 0; m < M; ++m) act_mask |= s_rid[m] >= 0 ? (1u << m) : 0u;
    if (!act_mask) return;                                         // uniform over the whole cluster
    auto active = [&](int m) { return (act_mask >> m) & 1u; };
    auto rid = [&](int m) { return s_rid[m]; };
    const int col_now = a.col_ptr ? *a.col_ptr : 0;
    int ts_n = 0;
    auto stamp = [&]() {
        if (a.tstamp && blockIdx.x == 0 && tid == 0 && ts_n < 256) a.tstamp[ts_n++] = clock64();
    };
    stamp();

    float *wpost = smem + L::WPOST, *wpre = smem + L::WPRE, *sx = smem + L::X, *scat = smem + L::CAT,
          *su = smem + L::U, *so = smem + L::O, *sh = smem + L::H, *sy = smem + L::Y, *sred = smem + L::RED,
          *sragg = smem + L::RAGG, *sqr = smem + L::QR, *sq = smem + L::Q, *ss = smem + L::S, *sagg = smem + L::AGG,
          *ssal = smem + L::SAL, *smerge = smem + L::MERGE;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::MBAR);  // [0] post buffer, [1] pre buffer
    uint64_t *xbar = mbar + 2;                                      // [5] exchange barriers: agg, u, o, h, y
    uint32_t post_par = 0, pre_par = 0, xpar = 0;                   // xpar: parity bit per exchange barrier
    const uint32_t sbase = smem_u32(smem);
    uint32_t rbase[CL];                                             // this CTA's shared window as seen ... of every peer
#pragma unroll
    for (int p = 0; p < CL; ++p) rbase[p] = mapa_u32(sbase, (uint32_t)p);
    // send four floats to the same shared-memory location of all CL CTAs, completing 16 bytes on their exchange barrier xb
    auto xsend4 = [&](const float *dst, int xb, const float4 v) {   // dst: 16-byte aligned
        const uint32_t off = smem_u32(dst) - sbase, boff = smem_u32(&xbar[xb]) - sbase;
#pragma unroll
        for (int p = 0; p < CL; ++p) st_async_v4(rbase[p] + off, v, rbase[p] + boff);
    };
    auto xexpect = [&](int xb, uint32_t bytes) {                   // once per exchange, any time before the wait
        if (tid == 0) mbar_expect_tx(&xbar[xb], bytes);
    };
    auto xwait = [&](int xb) {
        mbar_wait(&xbar[xb], (xpar >> xb) & 1u);
        xpar ^= 1u << xb;
    };
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // the sequence of PRE chunks this launch consumes: pre0, sub[0].pre, sub[1].pre (those that exist)
    // PRE chunks are consumed in the order pre0, sub[0].pre, sub[1].pre, ... (those that exist); `pre_cursor` is the
    // position (-1 = pre0, i = sub[i].pre) of the next one to request
    auto pre_at = [&](int pos) -> const float * { return pos < 0 ? a.pre0.w : a.sub[pos].pre.w; };
    auto pre_advance = [&](int pos) {
        while (pos < a.n_sub && pre_at(pos) == nullptr) ++pos;
        return pos;
    };
    int pre_cursor = pre_advance(-1);

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        for (int i = 0; i < 5; ++i) mbar_init(&xbar[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        if (a.n_sub > 0) chunk_request(wpost, a.sub[0].w + (size_t)c * cs_post::FLOATS, cs_post::FLOATS, &mbar[0]);
        if (pre_cursor < a.n_sub) chunk_request(wpre, pre_at(pre_cursor) + (size_t)c * cs_pre::FLOATS, cs_pre::FLOATS, &mbar[1]);
    }
    if (pre_cursor < a.n_sub) pre_cursor = pre_advance(pre_cursor + 1);

    // ---- residual rows; q / s / qr of the first layer when they come from an earlier launch ----------------------
    for (int m = warp; m < M; m += NWARP) {
        const int r = rid(m);
        const bool act = active(m);
        const float *xrow = s_alt[m] >= 0 ? a.x2 + (size_t)s_alt[m] * 128 : a.x + (size_t)r * 128;
        st4(sx + m * LD1 + 4 * lane, act ? ld4(xrow + 4 * lane) : z4);
        if (!a.pre0.w) {
            st4(sqr + m * 128 + 4 * lane, act ? ld4(a.qr + (size_t)r * 1024 + c * 128 + 4 * lane) : z4);
            if (lane < 4) st4(sq + m * 16 + 4 * lane, act ? ld4(a.q + (size_t)r * 128 + 16 * c + 4 * lane) : z4);
            else if (lane < 8) st4(ss + m * 16 + 4 * (lane - 4), act ? ld4(a.s + (size_t)r * 128 + 16 * c + 4 * (lane - 4)) : z4);
        }
    }
    // edge lists of every layer of this launch (they do not depend on anything computed here)
    AttnPre apre0, apre1, apre2;
    {
        int f0 = -1, f1 = -1, f2 = -1;
        for (int i = a.n_sub - 1; i >= 0; --i) {
            if (a.sub[i].elist == 0) f0 = i;
            else if (a.sub[i].elist == 1) f1 = i;
            else f2 = i;
        }
        apre0 = attn_prefetch<M>(a.sub[f0 < 0 ? 0 : f0], s_rid);
        apre1 = attn_prefetch<M>(a.sub[f1 < 0 ? 0 : f1], s_rid);
        apre2 = attn_prefetch<M>(a.sub[f2 < 0 ? 0 : f2], s_rid);
    }
    // grid barriers count the CTAs that own at least one active row (the others returned above)
    unsigned bar_target = 0, bar_step = 0;
    {
        bool any_sync = false;
        for (int i = 0; i < a.n_sub; ++i) any_sync |= a.sub[i].grid_sync != 0;
        if (any_sync) {
            __shared__ unsigned s_bar_step;
            if (tid == 0) {
                unsigned n = 0;
                for (int r0 = 0; r0 < a.rows.n_total; r0 += M) {
                    bool act = false;
                    for (int m = 0; m < M; ++m) act |= a.rows.active(r0 + m);
                    n += act ? CL : 0;
                }
                s_bar_step = n;
            }
            __syncthreads();
            bar_step = s_bar_step;
        }
    }
    // peers must be resident before anyone writes into their shared memory
    cluster.sync();
    stamp();

    // ---- LayerNorm + q/s/k/v projections + relative-query fold of one layer (layers.py:65-71, 106-108) ----------
    auto do_pre = [&](const PreArgs &P) {
        mbar_wait(&mbar[1], pre_par);
        pre_par ^= 1;
        for (int m = warp; m < M; m += NWARP)
            st4(su + m * LD1 + 4 * lane,
                ln128s(ld4(sx + m * LD1 + 4 * lane), wpre + cs_pre::LN_DST_G, wpre + cs_pre::LN_DST_B, lane));
        __syncthreads();
        slice_gemm<M, 32, 4>(su, LD1, wpre + cs_pre::WQS, 32, sred, [&](int m, int n, float v) {
            v += wpre[cs_pre::BQS + n];
            const int r = rid(m);
            const bool st = P.to_global && active(m);
            if (n < 16) {
                sq[m * 16 + n] = v;
                if (st) a.q[(size_t)r * 128 + 16 * c + n] = v;
            } else {
                ss[m * 16 + n - 16] = v;
                if (st) a.s[(size_t)r * 128 + 16 * c + n - 16] = v;
            }
        });
        __syncthreads();
        if (P.pre_kv) {
            const int col = col_now + P.col_add;
            slice_gemm<M, 32, 4>(su, LD1, wpre + cs_pre::WKV, 32, sred, [&](int m, int n, float v) {
                const int r = a.ride_row ? s_alt[m] : rid(m);     // ride-along launches: K|V of the carried row only
                if (active(m) && r >= 0) {
                    const size_t slot = P.kv_ring ? ((size_t)r * a.ring + (col & (a.ring - 1))) : (size_t)r;
                    const int o = n < 16 ? 16 * c + n : 128 + 16 * c + (n - 16);
                    P.kv_out[slot * 256 + o] = v + wpre[cs_pre::BKV + n];
                }
            });
            __syncthreads();
        }
        // qr[m][ch] = g_r[ch] * sum_d q[m][d] * Wkr[16c+d][ch]
        {
            const int ch = tid & 127;
            const float g = wpre[cs_pre::LN_R_G + ch];
            float wk[16];
#pragma unroll
            for (int d = 0; d < 16; ++d) wk[d] = wpre[cs_pre::WKR + d * 128 + ch];
            for (int m = tid >> 7; m < M; m += 2) {
                float acc = 0.f;
#pragma unroll
                for (int d4 = 0; d4 < 4; ++d4) {
                    const float4 qv = ld4(sq + m * 16 + 4 * d4);
                    acc = fmaf(qv.x, wk[4 * d4 + 0], acc);
                    acc = fmaf(qv.y, wk[4 * d4 + 1], acc);
                    acc = fmaf(qv.z, wk[4 * d4 + 2], acc);
                    acc = fmaf(qv.w, wk[4 * d4 + 3], acc);
                }
                acc *= g;
                sqr[m * 128 + ch] = acc;
                const int r = rid(m);
                if (P.to_global && active(m)) a.qr[(size_t)r * 1024 + c * 128 + ch] = acc;
            }
        }
        __syncthreads();
        if (tid == 0 && pre_cursor < a.n_sub)
            chunk_request(wpre, pre_at(pre_cursor) + (size_t)c * cs_pre::FLOATS, cs_pre::FLOATS, &mbar[1]);
        if (pre_cursor < a.n_sub) pre_cursor = pre_advance(pre_cursor + 1);
    };

    if (a.pre0.w) do_pre(a.pre0);
    stamp();

    for (int si = 0; si < a.n_sub; ++si) {
        const SubArgs &A = a.sub[si];
        // ---- edge attention of head c ------------------------------------------------------------------------
        if (A.grid_sync) {                                         // K/V rows of other clusters must have landed
            bar_target += bar_step;
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(a.grid_bar, 1u);
                unsigned seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.grid_bar) : "memory");
                } while (seen < bar_target);
            }
            __syncthreads();
        }
        attn_phase<M>(A, A.elist == 0 ? apre0 : (A.elist == 1 ? apre1 : apre2), c, sq, sqr, sagg, sragg, ssal, smerge);
        stamp();
        mbar_wait(&mbar[0], post_par);
        post_par ^= 1;
        __syncthreads();
        stamp();
        // xd = LN_dst(x) (every CTA, full rows);  ragg' = g_r * ragg + b_r * sal (own head)
        for (int m = warp; m < M; m += NWARP) {
            st4(scat + m * LD2 + 128 + 4 * lane,
                ln128s(ld4(sx + m * LD1 + 4 * lane), wpost + cs_post::LN_DST_G, wpost + cs_post::LN_DST_B, lane));
            if (A.has_pos) {
                const float4 v = ld4(sragg + m * LD1 + 4 * lane);
                const float4 g = ld4(wpost + cs_post::LN_R_G + 4 * lane), b = ld4(wpost + cs_post::LN_R_B + 4 * lane);
                const float sa = ssal[m];
                st4(sragg + m * LD1 + 4 * lane, make_float4(fmaf(g.x, v.x, b.x * sa), fmaf(g.y, v.y, b.y * sa),
                                                            fmaf(g.z, v.z, b.z * sa), fmaf(g.w, v.w, b.w * sa)));
            }
        }
        __syncthreads();
        // ---- agg2 = agg + Wvr ragg' + bvr * sal  -> all CTAs ---------------------------------------------------
        if (!A.has_attn) {
            // edge-less pass (insertion stage: K|V of source-only rows): agg2 = 0 for every head - nothing to exchange
            for (int o = tid; o < M * 32; o += NT) st4(scat + (o >> 5) * LD2 + 4 * (o & 31), z4);
            __syncthreads();
        } else {
        xexpect(0, M * 128 * 4);
        if (A.has_pos) {
            slice_gemm_x<M, 16, (M == 4 ? 2 : 4)>(sragg, LD1, wpost + cs_post::WVR, 32, sred,
                [&](int m, int n, float v) { return v + sagg[m * 16 + n] + wpost[cs_post::BVR + n] * ssal[m]; },
                [&](int m, int n, const float4 v) { xsend4(scat + m * LD2 + 16 * c + n, 0, v); });
        } else {
            for (int o = tid; o < M * 4; o += NT) xsend4(scat + (o >> 2) * LD2 + 16 * c + 4 * (o & 3), 0, ld4(sagg + 4 * o));
        }
        xwait(0);
        }
        stamp();
        // ---- gate: g = sigmoid(Wg [agg | xd] + bg);  u = agg + g * (s - agg) ----------------------------------
        xexpect(1, M * 128 * 4);
        slice_gemm_x<M, 16, (M == 4 ? 2 : 4)>(scat, LD2, wpost + cs_post::WG, 64, sred,
            [&](int m, int n, float v) {
                const float g = sigmoidf(v + wpost[cs_post::BG + n]);
                const float ag = scat[m * LD2 + 16 * c + n];
                return ag + g * (ss[m * 16 + n] - ag);
            },
            [&](int m, int n, const float4 u) { xsend4(su + m * LD1 + 16 * c + n, 1, u); });
        xwait(1);
        stamp();
        // ---- to_out ------------------------------------------------------------------------------------------
        xexpect(2, M * 128 * 4);
        slice_gemm_x<M, 16, (M == 4 ? 2 : 4)>(su, LD1, wpost + cs_post::WO, 32, sred,
            [&](int m, int n, float v) { return v + wpost[cs_post::BO + n]; },
            [&](int m, int n, const float4 v) { xsend4(so + m * LD1 + 16 * c + n, 2, v); });
        xwait(2);
        stamp();
        // x1 = x + LN_post(o);  so = LN_ffpre(x1)
        for (int m = warp; m < M; m += NWARP) {
            float4 o = ld4(so + m * LD1 + 4 * lane);
            o = ln128s(o, wpost + cs_post::LN_POST_G, wpost + cs_post::LN_POST_B, lane);
            const float4 x1 = add4(ld4(sx + m * LD1 + 4 * lane), o);
            st4(sx + m * LD1 + 4 * lane, x1);
            st4(so + m * LD1 + 4 * lane, ln128s(x1, wpost + cs_post::LN_FFPRE_G, wpost + cs_post::LN_FFPRE_B, lane));
        }
        __syncthreads();
        // ---- FFN ---------------------------------------------------------------------------------------------
        xexpect(3, M * 512 * 4);
        slice_gemm_x<M, 64, 4>(so, LD1, wpost + cs_post::W1, 32, sred,
            [&](int m, int n, float v) { return fmaxf(v + wpost[cs_post::B1 + n], 0.f); },
            [&](int m, int n, const float4 v) { xsend4(sh + m * LD5 + 64 * c + n, 3, v); });
        xwait(3);
        stamp();
        xexpect(4, M * 128 * 4);
        slice_gemm_x<M, 16, (M == 4 ? 2 : 4)>(sh, LD5, wpost + cs_post::W2, 128, sred,
            [&](int m, int n, float v) { return v + wpost[cs_post::B2 + n]; },
            [&](int m, int n, const float4 v) { xsend4(sy + m * LD1 + 16 * c + n, 4, v); });
        xwait(4);
        stamp();
        // x2 = x1 + LN_ffpost(y)
        const bool last = si + 1 == a.n_sub;
        for (int m = warp; m < M; m += NWARP) {
            const int r = rid(m);
            float4 f = ld4(sy + m * LD1 + 4 * lane);
            f = ln128s(f, wpost + cs_post::LN_FFPOST_G, wpost + cs_post::LN_FFPOST_B, lane);
            const float4 x2 = add4(ld4(sx + m * LD1 + 4 * lane), f);
            st4(sx + m * LD1 + 4 * lane, x2);
            if ((m & (CL - 1)) == c && active(m) && s_alt[m] < 0) { // row m is stored by CTA m % 8
                if (last && !a.no_store) st4(a.x + (size_t)r * 128 + 4 * lane, x2);
                if (A.trace_out) st4(A.trace_out + (size_t)r * 128 + 4 * lane, x2);
            }
        }
        __syncthreads();
        if (tid == 0 && !last)
            chunk_request(wpost, a.sub[si + 1].w + (size_t)c * cs_post::FLOATS, cs_post::FLOATS, &mbar[0]);
        stamp();
        if (A.pre.w) do_pre(A.pre);
        stamp();
    }
    // no CTA may exit while a peer can still write into its shared memory
    cluster.sync();
}

}  // namespace infgen
