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
};
struct LayerArgs {
    RowSpace rows;
    float *x;                  // [R][128] residual stream, updated in place
    float *q, *s, *qr;         // [R][128], [R][128], [R][8][128] hand-over between launches
    const int *col_ptr;        // device: current column (temporal ring slot)
    int ring;                  // ring depth
    PreArgs pre0;              // optional projections run before the first layer (else q/s/qr come from global)
    int n_sub;
    SubArgs sub[2];
    long long *tstamp;         // optional [32] clock64 stamps of CTA 0 (debug: phase breakdown)
};

template <int M>
struct LayerSmem {
    static constexpr int WPOST = 0;
    static constexpr int WPRE = WPOST + cs_post::FLOATS;
    static constexpr int X = WPRE + cs_pre::FLOATS;       // [M][128] residual
    static constexpr int CAT = X + M * 128;               // [M][256] agg | LN_dst(x)
    static constexpr int U = CAT + M * 256;               // [M][128]
    static constexpr int O = U + M * 128;                 // [M][128]
    static constexpr int H = O + M * 128;                 // [M][512]
    static constexpr int Y = H + M * 512;                 // [M][128]
    static constexpr int RED = Y + M * 128;               // k-split partials
    static constexpr int RAGG = RED + M * 256 + 256;      // [M][128] own head
    static constexpr int QR = RAGG + M * 128;             // [M][128] own head
    static constexpr int Q = QR + M * 128;                // [M][16]
    static constexpr int S = Q + M * 16;                  // [M][16]
    static constexpr int AGG = S + M * 16;                // [M][16]
    static constexpr int SAL = AGG + M * 16;              // [M] (padded to 16)
    static constexpr int MERGE = SAL + 16;                // [NWARP][160]
    static constexpr int MBAR = MERGE + NWARP * 160;      // 2 x uint64
    static constexpr int TOTAL = MBAR + 4;
    static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
};

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
//   X  shared, row-major, leading dimension ldx;  W shared, [K4][NL][4];  the K range is split over KS thread groups
//   (NL * KS == NT) and reduced through `red`;  epi(m, n, value) runs once per output.  Contains one __syncthreads().
// ---------------------------------------------------------------------------------------------------------------------
template <int M, int NL, int KS, typename Epi>
__device__ __forceinline__ void slice_gemm(const float *xs, int ldx, const float *w, int K4, float *red, Epi epi) {
    static_assert(NL * KS == NT, "thread mapping");
    constexpr int PAD = (NL == 16) ? 16 : 0;             // de-conflict the two k-groups of a warp
    constexpr int KSTRIDE = M * NL + PAD;
    static_assert(KS * KSTRIDE <= M * 256 + 256, "reduction scratch too small");
    const int tid = threadIdx.x, col = tid % NL, ks = tid / NL;
    float acc[M];
#pragma unroll
    for (int m = 0; m < M; ++m) acc[m] = 0.f;
#pragma unroll 4
    for (int k4 = ks; k4 < K4; k4 += KS) {
        const float4 wv = ld4(w + (k4 * NL + col) * 4);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const float4 x = ld4(xs + m * ldx + 4 * k4);
            acc[m] = fmaf(x.x, wv.x, acc[m]);
            acc[m] = fmaf(x.y, wv.y, acc[m]);
            acc[m] = fmaf(x.z, wv.z, acc[m]);
            acc[m] = fmaf(x.w, wv.w, acc[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) red[ks * KSTRIDE + m * NL + col] = acc[m];
    __syncthreads();
    for (int o = tid; o < M * NL; o += NT) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < KS; ++k) v += red[k * KSTRIDE + o];
        epi(o / NL, o % NL, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// edge attention of head `c` for the M rows of the cluster (layers.py:78-92), online softmax, NWARP / M warps per row
// ---------------------------------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void attn_phase(const SubArgs &A, const RowSpace &rows, int row0, int c, const float *sq,
                                           const float *sqr, float *sagg, float *sragg, float *ssal, float *smerge) {
    constexpr int WPR = NWARP / M;
    constexpr int CH = 8;                                 // edges in flight per warp
    static_assert(WPR == 1 || WPR == 2, "1 or 2 warps per row");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = warp / WPR, part = warp % WPR;
    const int r = row0 + m;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int n = 0, e0 = 0;
    if (A.has_attn && rows.active(r)) {
        n = A.cnt[r];
        e0 = A.start ? A.start[r] : r * A.stride;
    }
    // this warp's contiguous share of the row's edges
    const int share = WPR == 1 ? n : (((n + WPR - 1) / WPR + 3) & ~3);
    const int eb0 = part * share, eb1 = min(n, eb0 + share);
    const float4 qr4 = ld4(sqr + m * 128 + 4 * lane);
    const float4 q4 = lane < 4 ? ld4(sq + m * 16 + 4 * lane) : z4;
    float mx = -INFINITY, den = 0.f;
    float4 ra = z4, av = z4;
    const float *kvb = A.kv + 16 * c + 4 * (lane & 3);
    for (int blk = eb0; blk < eb1; blk += 32) {           // source rows of 32 edges with one coalesced load
        const int my_src = (blk + lane < eb1) ? A.src[e0 + blk + lane] : 0;
        const int blk_end = min(eb1, blk + 32);
        for (int eb = blk; eb < blk_end; eb += CH) {
            float p[CH];
            float4 rh[CH], v4[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int e = eb + j;
                const int sj = __shfl_sync(0xffffffffu, my_src, (e - blk) & 31);
                p[j] = 0.f; rh[j] = z4; v4[j] = z4;
                if (e < blk_end) {
                    if (A.has_pos) rh[j] = ld4(A.rhat + (size_t)(e0 + e) * 128 + 4 * lane);
                    if (lane < 4) {
                        p[j] = dot4(q4, ld4(kvb + (size_t)sj * 256));
                        v4[j] = ld4(kvb + (size_t)sj * 256 + 128);
                    }
                }
            }
            float pm = -INFINITY;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                p[j] = warp_sum(p[j] + dot4(qr4, rh[j])) * 0.25f;      // head_dim ** -0.5
                if (eb + j >= blk_end) p[j] = -INFINITY;
                pm = fmaxf(pm, p[j]);
            }
            const float mn = fmaxf(mx, pm);                   // finite: edge eb exists
            const float sc = expf(mx - mn);                   // 0 on the first chunk
            den *= sc;
            ra.x *= sc; ra.y *= sc; ra.z *= sc; ra.w *= sc;
            av.x *= sc; av.y *= sc; av.z *= sc; av.w *= sc;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const float w = expf(p[j] - mn);              // 0 for padded edges
                den += w;
                ra.x = fmaf(w, rh[j].x, ra.x); ra.y = fmaf(w, rh[j].y, ra.y);
                ra.z = fmaf(w, rh[j].z, ra.z); ra.w = fmaf(w, rh[j].w, ra.w);
                av.x = fmaf(w, v4[j].x, av.x); av.y = fmaf(w, v4[j].y, av.y);
                av.z = fmaf(w, v4[j].z, av.z); av.w = fmaf(w, v4[j].w, av.w);
            }
            mx = mn;
        }
    }
    if (WPR == 2) {
        float *mg = smerge + warp * 160;
        if (part == 1) {
            if (lane == 0) { mg[0] = mx; mg[1] = den; }
            if (lane < 4) st4(mg + 4 + 4 * lane, av);
            st4(mg + 32 + 4 * lane, ra);
        }
        __syncthreads();
        if (part == 0) {
            const float *og = smerge + (warp + 1) * 160;
            const float mx1 = og[0], den1 = og[1];
            const float mn = fmaxf(mx, mx1);
            if (mn > -INFINITY) {
                const float f0 = expf(mx - mn), f1 = expf(mx1 - mn);
                const float4 ra1 = ld4(og + 32 + 4 * lane);
                const float4 av1 = lane < 4 ? ld4(og + 4 + 4 * lane) : z4;
                den = den * f0 + den1 * f1;
                ra = make_float4(ra.x * f0 + ra1.x * f1, ra.y * f0 + ra1.y * f1, ra.z * f0 + ra1.z * f1,
                                 ra.w * f0 + ra1.w * f1);
                av = make_float4(av.x * f0 + av1.x * f1, av.y * f0 + av1.y * f1, av.z * f0 + av1.z * f1,
                                 av.w * f0 + av1.w * f1);
            }
        }
    }
    if (part == 0) {
        const float inv = 1.0f / (den + 1e-16f);          // torch_geometric.utils.softmax denominator
        if (lane < 4) st4(sagg + m * 16 + 4 * lane, make_float4(av.x * inv, av.y * inv, av.z * inv, av.w * inv));
        st4(sragg + m * 128 + 4 * lane, make_float4(ra.x * inv, ra.y * inv, ra.z * inv, ra.w * inv));
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
    const int row0 = (int)(blockIdx.x / CL) * M;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    bool any = false;
#pragma unroll
    for (int m = 0; m < M; ++m) any |= a.rows.active(row0 + m);
    if (!any) return;                                              // uniform over the whole cluster
    int ts_n = 0;
    auto stamp = [&]() {
        if (a.tstamp && blockIdx.x == 0 && tid == 0 && ts_n < 32) a.tstamp[ts_n++] = clock64();
    };
    stamp();

    float *wpost = smem + L::WPOST, *wpre = smem + L::WPRE, *sx = smem + L::X, *scat = smem + L::CAT,
          *su = smem + L::U, *so = smem + L::O, *sh = smem + L::H, *sy = smem + L::Y, *sred = smem + L::RED,
          *sragg = smem + L::RAGG, *sqr = smem + L::QR, *sq = smem + L::Q, *ss = smem + L::S, *sagg = smem + L::AGG,
          *ssal = smem + L::SAL, *smerge = smem + L::MERGE;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L::MBAR);  // [0] post buffer, [1] pre buffer
    uint32_t post_par = 0, pre_par = 0;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // the sequence of PRE chunks this launch consumes: pre0, sub[0].pre, sub[1].pre (those that exist)
    const float *pre_seq[3];
    int n_pre = 0;
    if (a.pre0.w) pre_seq[n_pre++] = a.pre0.w;
    for (int i = 0; i < a.n_sub; ++i)
        if (a.sub[i].pre.w) pre_seq[n_pre++] = a.sub[i].pre.w;
    int pre_next = 0;                                              // next PRE chunk to request

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        if (a.n_sub > 0) chunk_request(wpost, a.sub[0].w + (size_t)c * cs_post::FLOATS, cs_post::FLOATS, &mbar[0]);
        if (n_pre > 0) chunk_request(wpre, pre_seq[0] + (size_t)c * cs_pre::FLOATS, cs_pre::FLOATS, &mbar[1]);
    }
    pre_next = n_pre > 0 ? 1 : 0;

    // ---- residual rows; q / s / qr of the first layer when they come from an earlier launch ----------------------
    for (int m = warp; m < M; m += NWARP) {
        const int r = row0 + m;
        const bool act = a.rows.active(r);
        st4(sx + m * 128 + 4 * lane, act ? ld4(a.x + (size_t)r * 128 + 4 * lane) : z4);
        if (!a.pre0.w) {
            st4(sqr + m * 128 + 4 * lane, act ? ld4(a.qr + (size_t)r * 1024 + c * 128 + 4 * lane) : z4);
            if (lane < 4) st4(sq + m * 16 + 4 * lane, act ? ld4(a.q + (size_t)r * 128 + 16 * c + 4 * lane) : z4);
            else if (lane < 8) st4(ss + m * 16 + 4 * (lane - 4), act ? ld4(a.s + (size_t)r * 128 + 16 * c + 4 * (lane - 4)) : z4);
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
            st4(su + m * 128 + 4 * lane,
                ln128s(ld4(sx + m * 128 + 4 * lane), wpre + cs_pre::LN_DST_G, wpre + cs_pre::LN_DST_B, lane));
        __syncthreads();
        slice_gemm<M, 32, 8>(su, 128, wpre + cs_pre::WQS, 32, sred, [&](int m, int n, float v) {
            v += wpre[cs_pre::BQS + n];
            const int r = row0 + m;
            const bool st = P.to_global && a.rows.active(r);
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
            const int col = a.col_ptr ? (*a.col_ptr + P.col_add) : 0;
            slice_gemm<M, 32, 8>(su, 128, wpre + cs_pre::WKV, 32, sred, [&](int m, int n, float v) {
                const int r = row0 + m;
                if (a.rows.active(r)) {
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
                const int r = row0 + m;
                if (P.to_global && a.rows.active(r)) a.qr[(size_t)r * 1024 + c * 128 + ch] = acc;
            }
        }
        __syncthreads();
        if (tid == 0 && pre_next < n_pre)
            chunk_request(wpre, pre_seq[pre_next] + (size_t)c * cs_pre::FLOATS, cs_pre::FLOATS, &mbar[1]);
        if (pre_next < n_pre) ++pre_next;
    };

    if (a.pre0.w) do_pre(a.pre0);
    stamp();

    for (int si = 0; si < a.n_sub; ++si) {
        const SubArgs &A = a.sub[si];
        // ---- edge attention of head c ------------------------------------------------------------------------
        attn_phase<M>(A, a.rows, row0, c, sq, sqr, sagg, sragg, ssal, smerge);
        stamp();
        mbar_wait(&mbar[0], post_par);
        post_par ^= 1;
        __syncthreads();
        stamp();
        // xd = LN_dst(x) (every CTA, full rows);  ragg' = g_r * ragg + b_r * sal (own head)
        for (int m = warp; m < M; m += NWARP) {
            st4(scat + m * 256 + 128 + 4 * lane,
                ln128s(ld4(sx + m * 128 + 4 * lane), wpost + cs_post::LN_DST_G, wpost + cs_post::LN_DST_B, lane));
            if (A.has_pos) {
                const float4 v = ld4(sragg + m * 128 + 4 * lane);
                const float4 g = ld4(wpost + cs_post::LN_R_G + 4 * lane), b = ld4(wpost + cs_post::LN_R_B + 4 * lane);
                const float sa = ssal[m];
                st4(sragg + m * 128 + 4 * lane, make_float4(fmaf(g.x, v.x, b.x * sa), fmaf(g.y, v.y, b.y * sa),
                                                            fmaf(g.z, v.z, b.z * sa), fmaf(g.w, v.w, b.w * sa)));
            }
        }
        __syncthreads();
        // ---- agg2 = agg + Wvr ragg' + bvr * sal  -> all CTAs ---------------------------------------------------
        if (A.has_pos) {
            slice_gemm<M, 16, 16>(sragg, 128, wpost + cs_post::WVR, 32, sred, [&](int m, int n, float v) {
                v += sagg[m * 16 + n] + wpost[cs_post::BVR + n] * ssal[m];
#pragma unroll
                for (int p = 0; p < CL; ++p) cluster.map_shared_rank(scat, p)[m * 256 + 16 * c + n] = v;
            });
        } else {
            for (int o = tid; o < M * 16; o += NT) {
                const float v = sagg[o];
#pragma unroll
                for (int p = 0; p < CL; ++p) cluster.map_shared_rank(scat, p)[(o >> 4) * 256 + 16 * c + (o & 15)] = v;
            }
        }
        cluster.sync();
        stamp();
        // ---- gate: g = sigmoid(Wg [agg | xd] + bg);  u = agg + g * (s - agg) ----------------------------------
        slice_gemm<M, 16, 16>(scat, 256, wpost + cs_post::WG, 64, sred, [&](int m, int n, float v) {
            const float g = sigmoidf(v + wpost[cs_post::BG + n]);
            const float ag = scat[m * 256 + 16 * c + n];
            const float u = ag + g * (ss[m * 16 + n] - ag);
#pragma unroll
            for (int p = 0; p < CL; ++p) cluster.map_shared_rank(su, p)[m * 128 + 16 * c + n] = u;
        });
        cluster.sync();
        stamp();
        // ---- to_out ------------------------------------------------------------------------------------------
        slice_gemm<M, 16, 16>(su, 128, wpost + cs_post::WO, 32, sred, [&](int m, int n, float v) {
            v += wpost[cs_post::BO + n];
#pragma unroll
            for (int p = 0; p < CL; ++p) cluster.map_shared_rank(so, p)[m * 128 + 16 * c + n] = v;
        });
        cluster.sync();
        stamp();
        // x1 = x + LN_post(o);  so = LN_ffpre(x1)
        for (int m = warp; m < M; m += NWARP) {
            float4 o = ld4(so + m * 128 + 4 * lane);
            o = ln128s(o, wpost + cs_post::LN_POST_G, wpost + cs_post::LN_POST_B, lane);
            const float4 x1 = add4(ld4(sx + m * 128 + 4 * lane), o);
            st4(sx + m * 128 + 4 * lane, x1);
            st4(so + m * 128 + 4 * lane, ln128s(x1, wpost + cs_post::LN_FFPRE_G, wpost + cs_post::LN_FFPRE_B, lane));
        }
        __syncthreads();
        // ---- FFN ---------------------------------------------------------------------------------------------
        slice_gemm<M, 64, 4>(so, 128, wpost + cs_post::W1, 32, sred, [&](int m, int n, float v) {
            v = fmaxf(v + wpost[cs_post::B1 + n], 0.f);
#pragma unroll
            for (int p = 0; p < CL; ++p) cluster.map_shared_rank(sh, p)[m * 512 + 64 * c + n] = v;
        });
        cluster.sync();
        stamp();
        slice_gemm<M, 16, 16>(sh, 512, wpost + cs_post::W2, 128, sred, [&](int m, int n, float v) {
            v += wpost[cs_post::B2 + n];
#pragma unroll
            for (int p = 0; p < CL; ++p) cluster.map_shared_rank(sy, p)[m * 128 + 16 * c + n] = v;
        });
        cluster.sync();
        stamp();
        // x2 = x1 + LN_ffpost(y)
        const bool last = si + 1 == a.n_sub;
        for (int m = warp; m < M; m += NWARP) {
            const int r = row0 + m;
            float4 f = ld4(sy + m * 128 + 4 * lane);
            f = ln128s(f, wpost + cs_post::LN_FFPOST_G, wpost + cs_post::LN_FFPOST_B, lane);
            const float4 x2 = add4(ld4(sx + m * 128 + 4 * lane), f);
            st4(sx + m * 128 + 4 * lane, x2);
            if ((m & (CL - 1)) == c && a.rows.active(r)) {         // row m is stored by CTA m % 8
                if (last) st4(a.x + (size_t)r * 128 + 4 * lane, x2);
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
