// Throughput path of the AttentionLayer (reference layers.py:61-113) for batches whose row tiles no longer fit one wave of
// clusters (layer.cuh is the latency path: 8-CTA clusters, 8 rows each, at most 15 of them co-resident on a B200, so a
// batch of 32 scenes needs 18 waves of latency-bound phases per layer).  Here a layer is two grid-wide kernels:
//
//   k_attn    edge attention (layers.py:78-92) of ALL rows: one warp per destination row serves all 8 heads, thousands of
//             warps in flight - the gather of K/V rows and relative embeddings (1.5 KB per edge, each byte read once) is
//             bound by L2/HBM bandwidth, not by a dependent chain.  Writes the per-head softmax-weighted sums agg / ragg /
//             sal to global memory.  Same algebra as the attention phase of k_layer (relative-projection fold, online
//             softmax, + 1e-16 denominator).
//   k_node    everything else of the layer for a tile of 16 rows per CTA, whole rows local to the CTA (no cluster, no
//             DSMEM exchange): to_v_r fold -> gate -> to_out -> LN -> FFN -> LN (layers.py:74-75, 94-99), then the
//             LayerNorm + q/s/k/v projections and the relative-query fold of the NEXT layer (layers.py:65-71, 106-108).
//             The weights of both halves are streamed through the shared-memory ring of stream.cuh (cp.async.bulk +
//             mbarriers) from "node-packed" copies: every Linear cut into 128-column blocks stored contiguously in
//             consumption order, so the whole layer is two linear streams.
//
// The kernel boundary between k_node (writes K/V of the rows) and the next k_attn (reads K/V of any row of the scene)
// is the grid-wide dependency of the agent<->agent layers.
#pragma once
#include "common.cuh"
#include "stream.cuh"
#include "ops.cuh"
#include "layer.cuh"

namespace infgen {

// node-packed weights of one AttentionLayer (float offsets); every block is [k4][128][4]
namespace np {
constexpr int VR = 0;                       // [32]   to_v_r
constexpr int G = VR + 16384;               // [64]   to_g
constexpr int OUT = G + 32768;              // [32]   to_out
constexpr int FF1 = OUT + 16384;            // 4 x [32]  ff_mlp.0 columns 128j..
constexpr int FF2 = FF1 + 65536;            // [128]  ff_mlp.3
constexpr int QS = FF2 + 65536;             // 2 x [32]  to_q | to_s
constexpr int KV = QS + 32768;              // 2 x [32]  to_k | to_v
constexpr int FLOATS = KV + 32768;          // 262,144 floats = 1 MB
}  // namespace np

// dst[j][k4][n][4] = src[k4][128 j + n][4]: a packed [K4][N][4] matrix cut into 128-column blocks
__global__ void k_node_pack(const float *__restrict__ src, float *__restrict__ dst, int K4, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K4 * N) return;
    const int k4 = i / N, n = i % N, j = n >> 7, nl = n & 127;
    const float4 v = *reinterpret_cast<const float4 *>(src + (size_t)i * 4);
    *reinterpret_cast<float4 *>(dst + (((size_t)j * K4 + k4) * 128 + nl) * 4) = v;
}

// ---------------------------------------------------------------------------------------------------------------------
struct AttnArgs {
    RowSpace rows;
    SubArgs sub;               // kv, cnt, start, stride, src, rhat, has_attn, has_pos
    const float *q, *qr;       // [R][128], [R][8][128]
    float *agg;                // [R][128] OUT: the attention output of the row INCLUDING the folded relative value term,
                               //   agg2 = sum_e a_e v_j + Wvr (g_r * sum_e a_e rhat_e + b_r * sal) + bvr * sal   (ops.cuh algebra)
    const float *vrf;          // the layer's folded to_v_r table (vrf::FLOATS floats, k_vr_fold_pack); NULL when !has_pos
};
// folded to_v_r of one layer, in the access order of the k_attn epilogue:
//   W[((h * 4 + i) * 4 + j4) * 32 + lane][c] = g_r[4 lane + i] * Wvr[4 lane + i][16 h + 4 j4 + c]     (16,384 floats)
//   C[n] = sum_k b_r[k] Wvr[k][n] + bvr[n]                                                               (128 floats)
namespace vrf {
constexpr int W = 0, C = 16384, FLOATS = 16384 + 128;
}
// w_vr: packed [32 k4][128 n][4]
__global__ void __launch_bounds__(128) k_vr_fold_pack(const float *__restrict__ w_vr, const float *__restrict__ g_r,
                                                     const float *__restrict__ b_r, const float *__restrict__ b_vr,
                                                     float *__restrict__ out) {
    const int n = threadIdx.x;                                // output column
    auto wvr = [&](int k, int col) { return w_vr[((size_t)(k >> 2) * 128 + col) * 4 + (k & 3)]; };
    float c = b_vr[n];
    for (int k = 0; k < 128; ++k) c = fmaf(b_r[k], wvr(k, n), c);
    out[vrf::C + n] = c;
    for (int idx = threadIdx.x; idx < 16384; idx += blockDim.x) {
        const int cc = idx & 3, lane = (idx >> 2) & 31, j4 = (idx >> 7) & 3, i = (idx >> 9) & 3, h = idx >> 11;
        const int k = 4 * lane + i, col = 16 * h + 4 * j4 + cc;
        out[vrf::W + idx] = g_r[k] * wvr(k, col);
    }
}

constexpr int AW = 4;          // warps (= rows) per CTA

// One warp per destination row, ALL 8 heads: lane l holds float4 #l of every 128-vector, i.e. dims 4(l&3).. of head l>>2.
// An edge costs one coalesced 512-byte row each of rhat, K and V (1.5 KB - the head-per-CTA variant re-read rhat for every
// head: 5 KB per edge).  The eight relative scores (one per head) are reduced with the butterfly reduce-scatter of
// attn_phase, which leaves head h's score in lane group h, next to its K/V slice; each lane group runs its own online
// softmax.  Loads of the next two edges are in flight while two are reduced.
//
// What bounds it (measured, round 2): NOT the memory system - a plain LDG.128 gather of the same rows with two edges in
// flight per warp runs at 10-11 TB/s out of L2 and 7.2 TB/s out of HBM (tools/probe/gather_bw.cu,
// profiles/r2_gather_bw_probe.log; cp.async.bulk staging of the 1 KB rows reaches only 3-4 TB/s, and splitting a row's
// edges over four warps made the kernel slower) - but instruction issue: ~185 instructions per edge and warp, 80 of them
// the rescale-and-accumulate of the eight per-head relative sums.  Hence the LAZY rescale below: the softmax reference
// value of a head only moves when a score exceeds it by more than ATTN_LAZY (weights stay below e^ATTN_LAZY, far
// from fp32 overflow for <= 2048 edges), so the common edge needs no rescale multiply and half the broadcasts.
// Staging: every warp owns a ring of ATTN_DEPTH edge slots in shared memory ([rhat 512 B | K 512 B | V 512 B] each), filled
// with 16-byte cp.async copies (LDGSTS: no registers held while the rows are in flight), ATTN_DEPTH edges ahead of the edge
// being reduced.  The launch is a single wave of ~14 warps per SM, every warp walking its row's edges as one chain, so
// the depth of this prefetch - not bandwidth - sets the kernel time (with two edges in flight in registers the warp
// stalled on L2 latency for every pair: ncu long-scoreboard 2.4 of 6.7 stalled warps per issue).
constexpr int ATTN_DEPTH = 4;                          // (8 fits only one CTA per SM next to the 64 KB to_v_r table)
constexpr int ATTN_SLOT = 384;                         // floats per edge slot
constexpr int ATTN_WARPS = 8;                          // warps per (persistent) CTA
constexpr int ATTN_SMEM_FLOATS = vrf::FLOATS + ATTN_WARPS * ATTN_DEPTH * ATTN_SLOT;
constexpr size_t ATTN_SMEM = (size_t)ATTN_SMEM_FLOATS * sizeof(float);        // 113 KB: two CTAs per SM
__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent CTAs (2 per SM): the folded to_v_r table of the layer is brought into shared memory once per CTA, then every
// warp walks rows warp_global, warp_global + n_warps, ...
__global__ void __launch_bounds__(ATTN_WARPS * 32, 2) k_attn(const AttnArgs a) {
    extern __shared__ __align__(16) float smem_attn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SubArgs &A = a.sub;
    float *sW = smem_attn, *sC = smem_attn + vrf::C;
    float *ring = smem_attn + vrf::FLOATS + (size_t)warp * ATTN_DEPTH * ATTN_SLOT + 4 * lane;
    if (a.vrf) {
        for (int i = threadIdx.x; i < vrf::FLOATS / 4; i += ATTN_WARPS * 32) st4(sW + 4 * i, ldg4(a.vrf + 4 * i));
    }
    __syncthreads();
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    constexpr float ATTN_LAZY = 8.0f;
    // the g-th ACTIVE row of the row space (rows of scene b: [b * cap, b * cap + n_rows[b])), -1 past the end: the warps
    // deal the active rows round-robin, so a sparsely filled capacity row space does not unbalance them
    auto active_row = [&](int g) -> int {
        if (a.rows.cap == 0) return g < a.rows.n_total ? g : -1;
        const int ns = a.rows.n_total / a.rows.cap;
        int acc = 0;
        for (int b0 = 0; b0 < ns; b0 += 32) {
            const int nb = b0 + lane < ns ? a.rows.n_rows[b0 + lane] : 0;
            int inc = nb;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, g < acc + inc);
            if (hit) {
                const int l = __ffs(hit) - 1;
                const int excl = __shfl_sync(0xffffffffu, inc - nb, l);
                return (b0 + l) * a.rows.cap + (g - acc - excl);
            }
            acc += __shfl_sync(0xffffffffu, inc, 31);
        }
        return -1;
    };
    for (int g = blockIdx.x * ATTN_WARPS + warp;; g += gridDim.x * ATTN_WARPS) {
        const int r = active_row(g);
        if (r < 0) break;
        const int n = A.has_attn ? A.cnt[r] : 0;
        const int e0 = A.start ? A.start[r] : r * A.stride;
        const float *rhb = A.rhat + (size_t)e0 * 128 + 4 * lane;
        const float *kvb = A.kv + 4 * lane;
        // source rows of the edges, 64 at a time in two registers (edge e of the row: lane e & 31 of word (e >> 5) & 1)
        int src_w0 = lane < n ? A.src[e0 + lane] : 0;
        int src_w1 = 32 + lane < n ? A.src[e0 + 32 + lane] : 0;
        auto issue = [&](int e) {                            // edge e -> slot e % ATTN_DEPTH (all lanes; e < n)
            const int s0 = __shfl_sync(0xffffffffu, src_w0, e & 31), s1 = __shfl_sync(0xffffffffu, src_w1, e & 31);
            const int sj = (e & 32) ? s1 : s0;
            float *d = ring + (e % ATTN_DEPTH) * ATTN_SLOT;
            if (A.has_pos) cp_async16(d, rhb + (size_t)e * 128);
            const float *p = kvb + (size_t)sj * 256;
            cp_async16(d + 128, p);
            cp_async16(d + 256, p + 128);
        };
        for (int e = 0; e < ATTN_DEPTH; ++e) {               // prologue: one commit group per edge slot
            if (e < n) issue(e);
            cp_async_commit();
        }
        const float4 q4 = ld4(a.q + (size_t)r * 128 + 4 * lane);
        float4 qr4[8], ra[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            qr4[h] = A.has_pos ? ld4(a.qr + ((size_t)r * 8 + h) * 128 + 4 * lane) : z4;
            ra[h] = z4;
        }
        float mx = -INFINITY, den = 0.f;                      // mx: softmax reference of this lane's head (one of its scores,
        float4 av = z4;                                       //     at most ATTN_LAZY below the running maximum)
        // score of one edge for the head of this lane's group (all four lanes of the group hold it)
        auto score = [&](const float4 rh, const float4 k) {
            float pk = dot4(q4, k);
            pk += __shfl_xor_sync(0xffffffffu, pk, 1);
            pk += __shfl_xor_sync(0xffffffffu, pk, 2);
            float v[8];
#pragma unroll
            for (int h = 0; h < 8; ++h) v[h] = dot4(qr4[h], rh);
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
            return (pr + pk) * 0.25f;                        // head_dim ** -0.5
        };
        for (int e = 0; e < n; e += 2) {
            // edges e, e + 1 have landed when at most ATTN_DEPTH - 2 younger groups are pending
            cp_async_wait<ATTN_DEPTH - 2>();
            const bool two = e + 1 < n;
            const float *d0 = ring + (e % ATTN_DEPTH) * ATTN_SLOT, *d1 = ring + ((e + 1) % ATTN_DEPTH) * ATTN_SLOT;
            // (every lane reads back exactly the 16-byte pieces it copied itself: no cross-lane hazard on the ring)
            const float4 rh0 = A.has_pos ? ld4(d0) : z4, k0 = ld4(d0 + 128), v0 = ld4(d0 + 256);
            float4 rh1 = z4, k1 = z4, v1 = z4;
            if (two) { rh1 = A.has_pos ? ld4(d1) : z4; k1 = ld4(d1 + 128); v1 = ld4(d1 + 256); }
            // the next two edges of the ring (two commit groups; empty ones past the end keep the group count in step)
            if (((e + ATTN_DEPTH) & 63) == 0 && e + ATTN_DEPTH < n) {     // crossing into the next 64 edges: refill the source words
                src_w0 = e + ATTN_DEPTH + lane < n ? A.src[e0 + e + ATTN_DEPTH + lane] : 0;
                src_w1 = e + ATTN_DEPTH + 32 + lane < n ? A.src[e0 + e + ATTN_DEPTH + 32 + lane] : 0;
            }
            if (e + ATTN_DEPTH < n) issue(e + ATTN_DEPTH);
            cp_async_commit();
            if (e + ATTN_DEPTH + 1 < n) issue(e + ATTN_DEPTH + 1);
            cp_async_commit();
            // two edges per step: their score chains are independent
            const float p0 = score(rh0, k0);
            const float p1 = two ? score(rh1, k1) : -INFINITY;
            const float pm = fmaxf(p0, p1);
            if (__any_sync(0xffffffffu, pm > mx + ATTN_LAZY)) {
                // some head's score left the window above its reference (always on a row's first edges): move the
                // reference of the heads concerned and rescale their sums
                const float mn = pm > mx + ATTN_LAZY ? pm : mx;
                const float sc = expf(mx - mn);              // 0 on the first edge, 1 for heads that keep their reference
                den *= sc;
                av.x *= sc; av.y *= sc; av.z *= sc; av.w *= sc;
                mx = mn;
                if (A.has_pos) {
#pragma unroll
                    for (int h = 0; h < 8; ++h) {
                        const float sh = __shfl_sync(0xffffffffu, sc, 4 * h);
                        ra[h].x *= sh; ra[h].y *= sh; ra[h].z *= sh; ra[h].w *= sh;
                    }
                }
            }
            const float w0 = expf(p0 - mx), w1 = expf(p1 - mx);   // w1 = 0 for an absent edge
            den += w0 + w1;
            av.x = fmaf(w0, v0.x, fmaf(w1, v1.x, av.x)); av.y = fmaf(w0, v0.y, fmaf(w1, v1.y, av.y));
            av.z = fmaf(w0, v0.z, fmaf(w1, v1.z, av.z)); av.w = fmaf(w0, v0.w, fmaf(w1, v1.w, av.w));
            if (A.has_pos) {
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float g0 = __shfl_sync(0xffffffffu, w0, 4 * h), g1 = __shfl_sync(0xffffffffu, w1, 4 * h);
                    ra[h].x = fmaf(g0, rh0.x, fmaf(g1, rh1.x, ra[h].x)); ra[h].y = fmaf(g0, rh0.y, fmaf(g1, rh1.y, ra[h].y));
                    ra[h].z = fmaf(g0, rh0.z, fmaf(g1, rh1.z, ra[h].z)); ra[h].w = fmaf(g0, rh0.w, fmaf(g1, rh1.w, ra[h].w));
                }
            }
        }
        cp_async_wait<0>();
        const float inv = 1.0f / (den + 1e-16f);             // torch_geometric.utils.softmax denominator
        float4 out = make_float4(av.x * inv, av.y * inv, av.z * inv, av.w * inv);
        if (A.has_pos && a.vrf) {
            // ---- folded relative value term: agg2[16h + j] += sum_k ragg_h[k] W'[k][16h + j] + C[16h + j] * sal_h ----------
            // lane l holds ragg_h[4l .. 4l+3]: 16 partial products per head, reduce-scattered over the warp (column
            // j = lane >> 1 ends up in lanes 2j, 2j + 1), staged through the warp's first ring slot (idle now)
            float *sv = ring - 4 * lane;                     // [128] scratch
            __syncwarp();
#pragma unroll 1
            for (int h = 0; h < 8; ++h) {
                const float ih = __shfl_sync(0xffffffffu, inv, 4 * h);
                const float rr[4] = {ra[h].x * ih, ra[h].y * ih, ra[h].z * ih, ra[h].w * ih};
                float p[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) p[j] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 w = ld4(sW + ((size_t)((h * 4 + i) * 4 + j4) * 32 + lane) * 4);
                        p[4 * j4 + 0] = fmaf(rr[i], w.x, p[4 * j4 + 0]); p[4 * j4 + 1] = fmaf(rr[i], w.y, p[4 * j4 + 1]);
                        p[4 * j4 + 2] = fmaf(rr[i], w.z, p[4 * j4 + 2]); p[4 * j4 + 3] = fmaf(rr[i], w.w, p[4 * j4 + 3]);
                    }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float recv = __shfl_xor_sync(0xffffffffu, b4 ? p[i] : p[i + 8], 16);
                    p[i] = (b4 ? p[i + 8] : p[i]) + recv;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float recv = __shfl_xor_sync(0xffffffffu, b3 ? p[i] : p[i + 4], 8);
                    p[i] = (b3 ? p[i + 4] : p[i]) + recv;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float recv = __shfl_xor_sync(0xffffffffu, b2 ? p[i] : p[i + 2], 4);
                    p[i] = (b2 ? p[i + 2] : p[i]) + recv;
                }
                {
                    const float recv = __shfl_xor_sync(0xffffffffu, b1 ? p[0] : p[1], 2);
                    p[0] = (b1 ? p[1] : p[0]) + recv;
                }
                p[0] += __shfl_xor_sync(0xffffffffu, p[0], 1);
                // column index of the value this lane ends up with: bit 3 = b4, bit 2 = b3, bit 1 = b2, bit 0 = b1
                if ((lane & 1) == 0) sv[16 * h + (lane >> 1)] = p[0];
            }
            __syncwarp();
            const float4 vr = ld4(sv + 4 * lane), c4 = ld4(sC + 4 * lane);
            const float sal = den * inv;                     // of this lane's head (lane >> 2)
            out.x += vr.x + c4.x * sal; out.y += vr.y + c4.y * sal; out.z += vr.z + c4.z * sal; out.w += vr.w + c4.w * sal;
            __syncwarp();                                    // the scratch is a ring slot again from the next row on
        }
        st4(a.agg + (size_t)r * 128 + 4 * lane, out);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
struct NodeArgs {
    RowSpace rows;
    float *x;                  // [R][128] residual stream, updated in place
    const float *agg;          // [R][128] attention output of k_attn (relative value term folded in)
    float *q, *s, *qr;         // hand-over: s of the post layer is read, q / s / qr of the pre layer are written
    const float *w_post;       // node-packed weights of the layer being finished (NULL: only the pre half runs)
    AttnW lw;                  // its bias / LayerNorm vectors
    const float *w_pre;        // node-packed weights of the following layer (NULL: none)
    AttnW pw;
    int pre_kv;                // also project k|v of these rows (non-bipartite layers)
    int edgeless;              // rows without incoming edges (insertion stage: K|V of source-only rows): agg = 0, and the
                               // pre half only produces s (and k|v): no q, no relative queries
    float *kv_out;
    int kv_ring, col_add, ring;
    const int *col_ptr;
    float *trace_out;          // optional copy of the layer output [R][128]
    const float *tc_post, *tc_pre;   // tensor-core weight images of the two layers (node_tc.cuh), k_node_tc only
};
#ifdef INFGEN_NODE_TRACE
__device__ long long g_node_trace[32];   // debug: clock64 stamps of consumer thread 0 of CTA 0 at the phase boundaries
#define NODE_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_node_trace[i] = clock64(); } while (0)
#else
#define NODE_STAMP(i) do {} while (0)
#endif
constexpr int NM = 16;         // rows per CTA
constexpr int NODE_STAGES = 4;  // weight ring depth (see stream.cuh)

struct NodeSmem {
    static constexpr int RING = 0;
    static constexpr int X = RING + ws_smem_floats<NODE_STAGES>();     // [NM][LD1] residual
    static constexpr int S = X + NM * LD1;              // [NM][LD1] skip projection of the post layer
    static constexpr int CAT = S + NM * LD1;            // [NM][LD2] agg2 | LN_dst(x)
    static constexpr int U = CAT + NM * LD2;            // [NM][LD1]
    static constexpr int O = U + NM * LD1;              // [NM][LD1]
    static constexpr int Y = O + NM * LD1;              // [NM][LD1]
    static constexpr int SAL = Y + NM * LD1;            // [NM][8]
    static constexpr int RED = SAL + NM * 8;            // 2 x [NM][RED_LD] k-split partial sums (alternating)
    static constexpr int RAGG = RED + 2 * NM * RED_LD;      // [8 heads][NM][LD1] normalised relative sums; later the FFN hidden
    static constexpr int TOTAL = RAGG + 8 * NM * LD1;   //                      tile [NM][LD5] and the q tile [NM][LD1]
    static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
    static_assert(NM * LD5 <= 8 * NM * LD1, "FFN hidden tile must fit the ragg region");
    static_assert(BYTES <= 227 * 1024, "shared memory budget");
};

// MMA: the Linears run on mma.sync TF32 with the 3xTF32 split (stream_gemm_mma) instead of the FFMA register tile
template <bool MMA>
__global__ void __launch_bounds__(NT_S) k_node(const NodeArgs a) {
    extern __shared__ __align__(16) float smem_n[];
    using L = NodeSmem;
    float *smem = smem_n;
    WsSmemT<NODE_STAGES> wsm(smem + L::RING);
    float *sx = smem + L::X, *ss = smem + L::S, *scat = smem + L::CAT, *su = smem + L::U, *so = smem + L::O, *sy = smem + L::Y,
          *ssal = smem + L::SAL, *sred = smem + L::RED, *sr = smem + L::RAGG, *sh = smem + L::RAGG, *sq = smem + L::RAGG;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * NM;
    unsigned act_mask = 0;
#pragma unroll
    for (int m = 0; m < NM; ++m) act_mask |= a.rows.active(row0 + m) ? (1u << m) : 0u;
    if (!act_mask) return;
    auto active = [&](int m) { return (act_mask >> m) & 1u; };
    const bool post = a.w_post != nullptr, pre = a.w_pre != nullptr;
    const int has_pos = a.lw.has_pos;
    NODE_STAMP(0);
    ws_init(wsm);
    if (warp == NWARP) {
        if (lane < NODE_STAGES) {
            WSeg segs[2];
            int n = 0;
            if (post) segs[n++] = WSeg{a.w_post + np::G, 352, 512};      // (to_v_r is folded into k_attn)
            if (pre) segs[n++] = WSeg{a.w_pre + np::QS, a.pre_kv ? 128 : 64, 512};
            ws_produce(wsm, segs, n);
        }
        return;
    }
    WsConsT<NODE_STAGES> ws(wsm);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int red_sel = 0;
    auto gemm = [&](const float *xs, int ldx, int K4, auto epi) {
        if constexpr (MMA) stream_gemm_mma<NM>(ws, xs, ldx, K4, epi);
        else {
            stream_gemm_ks16(ws, xs, ldx, K4, sred + red_sel * NM * RED_LD, epi);
            red_sel ^= 1;
        }
    };

    // ---- residual rows, skip projection, attention sums -----------------------------------------------------------------
    for (int m = warp; m < NM; m += NWARP) {
        const int r = row0 + m;
        const bool act = active(m);
        const float4 x = act ? ld4(a.x + (size_t)r * 128 + 4 * lane) : z4;
        st4(sx + m * LD1 + 4 * lane, x);
        if (post) {
            st4(ss + m * LD1 + 4 * lane, act ? ld4(a.s + (size_t)r * 128 + 4 * lane) : z4);
            st4(scat + m * LD2 + 4 * lane, (act && !a.edgeless) ? ld4(a.agg + (size_t)r * 128 + 4 * lane) : z4);
            st4(scat + m * LD2 + 128 + 4 * lane, ln128(x, a.lw.ln_dst_g, a.lw.ln_dst_b, lane));
        }
    }
    csync();
    NODE_STAMP(1);
    if (post) {
        // (agg arrives from k_attn with the relative value term already folded in: agg2 of ops.cuh)
        NODE_STAMP(2);
        // ---- gate: g = sigmoid(Wg [agg2 | xd] + bg);  u = agg2 + g * (s - agg2) ---------------------------------------------
        gemm(scat, LD2, 64, [&](int m, int n, float v) {
            const float g = sigmoidf(v + __ldg(a.lw.b_g + n));
            const float ag = scat[m * LD2 + n];
            su[m * LD1 + n] = ag + g * (ss[m * LD1 + n] - ag);
        });
        csync();
        NODE_STAMP(3);
        // ---- to_out -------------------------------------------------------------------------------------------------------
        gemm(su, LD1, 32, [&](int m, int n, float v) { so[m * LD1 + n] = v + __ldg(a.lw.b_out + n); });
        csync();
        NODE_STAMP(4);
        // x1 = x + LN_post(o);  so = LN_ffpre(x1)
        for (int m = warp; m < NM; m += NWARP) {
            float4 o = ld4(so + m * LD1 + 4 * lane);
            o = ln128(o, a.lw.ln_post_g, a.lw.ln_post_b, lane);
            const float4 x1 = add4(ld4(sx + m * LD1 + 4 * lane), o);
            st4(sx + m * LD1 + 4 * lane, x1);
            st4(so + m * LD1 + 4 * lane, ln128(x1, a.lw.ln_ffpre_g, a.lw.ln_ffpre_b, lane));
        }
        csync();
        NODE_STAMP(5);
        // ---- FFN ----------------------------------------------------------------------------------------------------------
        for (int j = 0; j < 4; ++j)
            gemm(so, LD1, 32, [&](int m, int n, float v) {
                sh[m * LD5 + 128 * j + n] = fmaxf(v + __ldg(a.lw.b_ff1 + 128 * j + n), 0.f);
            });
        csync();
        NODE_STAMP(6);
        gemm(sh, LD5, 128, [&](int m, int n, float v) { sy[m * LD1 + n] = v + __ldg(a.lw.b_ff2 + n); });
        csync();
        NODE_STAMP(7);
        // x2 = x1 + LN_ffpost(y)
        for (int m = warp; m < NM; m += NWARP) {
            const int r = row0 + m;
            float4 f = ld4(sy + m * LD1 + 4 * lane);
            f = ln128(f, a.lw.ln_ffpost_g, a.lw.ln_ffpost_b, lane);
            const float4 x2 = add4(ld4(sx + m * LD1 + 4 * lane), f);
            st4(sx + m * LD1 + 4 * lane, x2);
            if (active(m)) {
                st4(a.x + (size_t)r * 128 + 4 * lane, x2);
                if (a.trace_out) st4(a.trace_out + (size_t)r * 128 + 4 * lane, x2);
            }
        }
        csync();
    }
    NODE_STAMP(8);
    if (!pre) return;
    // ---- LayerNorm + q/s/k/v projections + relative-query fold of the next layer (layers.py:65-71, 106-108) ---------------
    for (int m = warp; m < NM; m += NWARP)
        st4(su + m * LD1 + 4 * lane, ln128(ld4(sx + m * LD1 + 4 * lane), a.pw.ln_dst_g, a.pw.ln_dst_b, lane));
    csync();
    gemm(su, LD1, 32, [&](int m, int n, float v) {
        v += __ldg(a.pw.b_qs + n);
        sq[m * LD1 + n] = v;
        if (active(m) && !a.edgeless) a.q[(size_t)(row0 + m) * 128 + n] = v;
    });
    gemm(su, LD1, 32, [&](int m, int n, float v) {
        if (active(m)) a.s[(size_t)(row0 + m) * 128 + n] = v + __ldg(a.pw.b_qs + 128 + n);
    });
    NODE_STAMP(9);
    // slice of Wkr for the relative-query fold below (thread = head `warp`, channels 4 lane ..): requested now, so that
    // the loads are in flight during the k/v projections
    float4 wk[16];
    const bool fold = a.pw.has_pos && !a.edgeless;
    if (fold) {
#pragma unroll
        for (int d = 0; d < 16; ++d) wk[d] = ldg4(a.pw.w_kr + (size_t)(16 * warp + d) * 128 + 4 * lane);
    }
    if (a.pre_kv) {
        const int col = (a.col_ptr ? *a.col_ptr : 0) + a.col_add;
        for (int j = 0; j < 2; ++j)
            gemm(su, LD1, 32, [&](int m, int n, float v) {
                if (active(m)) {
                    const int r = row0 + m;
                    const size_t slot = a.kv_ring ? ((size_t)r * a.ring + (col & (a.ring - 1))) : (size_t)r;
                    a.kv_out[slot * 256 + 128 * j + n] = v + __ldg(a.pw.b_kv + 128 * j + n);
                }
            });
    }
    csync();
    NODE_STAMP(10);
    // qr[m][h][ch] = g_r[ch] * sum_d q[m][16h+d] * Wkr[16h+d][ch]
    // thread = (head h = warp, channels 4 lane .. 4 lane + 3): the 16 x 4 slice of Wkr stays in registers and a row costs
    // 4 broadcast LDS.128 of q for 64 FMA (one thread per channel needed 4 LDS.128 per 16 FMA and was LSU-bound:
    // 13-15 k cycles for 262 k MAC, now ~3 k)
    if (fold) {
        const int h = warp;
        const float4 g4 = ldg4(a.pw.ln_r_g + 4 * lane);
        for (int m = 0; m < NM; ++m) {
            if (!active(m)) continue;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int d4 = 0; d4 < 4; ++d4) {
                const float4 qv = ld4(sq + m * LD1 + 16 * h + 4 * d4);
                const float qd[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 w = wk[4 * d4 + i];
                    acc.x = fmaf(qd[i], w.x, acc.x); acc.y = fmaf(qd[i], w.y, acc.y);
                    acc.z = fmaf(qd[i], w.z, acc.z); acc.w = fmaf(qd[i], w.w, acc.w);
                }
            }
            st4(a.qr + ((size_t)(row0 + m) * 8 + h) * 128 + 4 * lane,
                make_float4(acc.x * g4.x, acc.y * g4.y, acc.z * g4.z, acc.w * g4.w));
        }
    }
    NODE_STAMP(11);
}

}  // namespace infgen
