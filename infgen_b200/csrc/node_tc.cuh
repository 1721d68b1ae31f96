// The dense half of an AttentionLayer (reference layers.py:65-75, 94-113) for tiles of 128 rows on the 5th-generation
// tensor cores: k_node (node.cuh) with every Linear a chain of tcgen05.mma kind::tf32 instructions (M = 128, N = 128,
// K = 8), 3xTF32 error-compensated split, accumulators and the A operands in tensor memory - the machinery of
// k_fourier_tc (fourier_tc.cuh) applied to
//
//   post (finish layer lw):  gate  G  = [agg2 | LN_dst(x)] Wg            (K = 256)    u = agg2 + sigmoid(G + bg) (s - agg2)
//                            out   o  = u Wout + bout                     (K = 128)    x1 = x + LN_post(o)
//                            FFN   h_j = relu(LN_ffpre(x1) W1[:, 128j..] + b1), y += h_j W2[128j.., :]   j = 0..3
//                                  x2 = x1 + LN_ffpost(y + b2)
//   pre (project layer pw):  s | k | v | q = LN_dst'(x2) {Ws, Wk, Wv, Wq} + b    (one A operand, four GEMMs)
//                            qr_h = g_r * (q_h Wkr_h)   h = 0..7                  (K = 16 per head: two k-steps of chunk h / 2)
//
// One CTA per 128 ACTIVE rows of the row space (the rows of a capacity row space are compacted on the fly, so a batch of
// 32 scenes with ~90 of 224 rows in use is 23 tiles, not 56).  A launch is therefore a few dozen CTAs, each of which
// keeps its tensor core busy for the whole kernel: 768 MMAs per tile and layer (262,144 MAC per row x 3 passes),
// 49 k cycles at the 64-cycle issue floor of an M = N = 128 instruction.
//
// Roles (576 threads), as in k_fourier_tc: warp 0 issues the MMAs (uniform control flow, elect.sync), warps 1-16 are row
// threads (thread = one row x 32 of its 128 columns; TMEM lane group = warp % 4), warp 17 streams the weight chunks
// (32 KB cp.async.bulk each: [hi 16 KB | lo 16 KB] of a 32-k x 128-n block in the no-swizzle K-major core-matrix layout)
// through a 4-stage ring.  The GEMMs of a tile form one dependent chain (every LayerNorm needs whole rows), so the job
// list below is walked in the same order by all three roles; accumulators alternate between two TMEM regions so that the
// epilogue of one job overlaps the MMAs of the next wherever the data flow allows it (s / k / v / q, the eight heads of
// the relative-query fold, FFN up_{j+1} behind down_j).
//
// TMEM columns: [0,128) accumulator 0, [128,256) accumulator 1, [256,512) four A stages of [hi 32 | lo 32] columns.
// Global rows travel through per-warp shared-memory staging tiles (a thread owns 32 columns of ONE row, so direct vector
// loads / stores of a warp would touch 32 cache lines per instruction); the residual row x1 stays in that tile during the
// FFN: 576 threads leave 112 registers per thread, not enough for a second persistent 32-float row slice.
#pragma once
#include "common.cuh"
#include "ops.cuh"
#include "fourier_tc.cuh"
#include "node.cuh"

namespace infgen {
namespace ntc {
constexpr int TM = 128;
constexpr int NB = 4;                              // weight ring stages
constexpr uint32_t ACC0 = 0, ACC1 = 128, TC_A = 256;
constexpr int CHUNK = ftc::CHUNK;                  // floats per chunk: [hi 4096 | lo 4096]
constexpr int RT = 512, THREADS = RT + 64;
// 128 x 128 weight blocks of one layer image, 4 chunks each, in this order
enum { B_G0 = 0, B_G1, B_OUT, B_UP0, B_DN0, B_UP1, B_DN1, B_UP2, B_DN2, B_UP3, B_DN3, B_S, B_K, B_V, B_Q, B_KR, N_BLK };
constexpr size_t IMG_FLOATS = (size_t)N_BLK * 4 * CHUNK;       // 2 MB per layer
constexpr int SM_B = 0;
constexpr int SM_EX = SM_B + NB * CHUNK;           // [2 buffers][4 quarters][128] LayerNorm partials
constexpr int TILE_LD = 36, TILE_FLOATS = 32 * TILE_LD;     // per-warp staging tile [32 rows][32 + 4 columns]
constexpr int SM_TILE = SM_EX + 1024;              // [16 row warps][TILE_FLOATS]
constexpr int SM_ROW = SM_TILE + 16 * TILE_FLOATS; // [128] int: global row of every tile row (-1: none)
constexpr int SM_BAR = SM_ROW + 128;               // full_a[4] empty_a[4] full_b[NB] empty_b[NB] acc_done[2] acc_free[2]
constexpr int N_BAR = 8 + 2 * NB + 4;
constexpr int SM_TMEM = SM_BAR + 2 * N_BAR;
constexpr int SM_FLOATS = SM_TMEM + 4;
constexpr size_t SMEM = (size_t)SM_FLOATS * sizeof(float);
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr int MAX_JOBS = 20;
#ifndef INFGEN_NTC_BACKOFF
#define INFGEN_NTC_BACKOFF 0
#endif
// row warps poll their mbarriers without __nanosleep: one lane per warp polls, and a sleeping poller was measured to add
// ~1-2 k cycles to every hand-over of the dependent chain (the sleep is far coarser than the 32 ns asked for)
constexpr bool NTC_ROW_BACKOFF = INFGEN_NTC_BACKOFF != 0;
// job code: blk[0:5) img[5] acc[6] accum[7] fresh[8] release[9] done[10] fold[11]
__host__ __device__ constexpr uint32_t job(int blk, int img, int acc, int accum, int fresh, int done, int fold = 0) {
    return (uint32_t)blk | ((uint32_t)img << 5) | ((uint32_t)acc << 6) | ((uint32_t)accum << 7) | ((uint32_t)fresh << 8) |
           ((uint32_t)done << 10) | ((uint32_t)fold << 11);
}
// the GEMMs of one launch in issue order (shared by the weight producer, the MMA warp and - implicitly - the row threads)
__device__ __forceinline__ int job_list(bool post, bool pre, bool pre_kv, bool edgeless, bool fold, uint32_t *J) {
    int n = 0;
    if (post) {
        if (!edgeless) J[n++] = job(B_G0, 0, 0, 0, 1, 0);
        J[n++] = job(B_G1, 0, 0, edgeless ? 0 : 1, 1, 1);
        J[n++] = job(B_OUT, 0, 1, 0, 1, 1);
        for (int j = 0; j < 4; ++j) {
            J[n++] = job(B_UP0 + 2 * j, 0, 0, 0, 1, 1);
            J[n++] = job(B_DN0 + 2 * j, 0, 1, j > 0, 1, j == 3);
        }
    }
    if (pre) {
        int t = 0;
        J[n++] = job(B_S, 1, t, 0, 1, 1); t ^= 1;
        if (pre_kv) {
            J[n++] = job(B_K, 1, t, 0, 0, 1); t ^= 1;
            J[n++] = job(B_V, 1, t, 0, 0, 1); t ^= 1;
        }
        if (!edgeless) { J[n++] = job(B_Q, 1, t, 0, 0, 1); t ^= 1; }
        if (fold) J[n++] = job(B_KR, 1, t, 0, 1, 1, 1);
    }
    for (int i = 0; i < n; ++i)                    // a job releases its A stages unless the next one re-uses them
        if (i == n - 1 || ((J[i + 1] >> 8) & 1u)) J[i] |= 1u << 9;
    return n;
}
}  // namespace ntc

__device__ int g_ntc_hang[8];
// bounded mbarrier wait (see ftc_wait): code / 100 selects the class slot of g_ntc_hang
template <bool BACKOFF = false>
__device__ __forceinline__ void ntc_wait(uint64_t *b, uint32_t parity, int code) {
    const uint32_t addr = smem_u32(b);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (BACKOFF) __nanosleep(32);
        if (clock64() - t0 > 20000000ll) {
            atomicAdd(&g_ntc_hang[0], 1);
            atomicCAS(&g_ntc_hang[min(code / 100, 7)], 0, code * 1000 + (int)threadIdx.x);
            return;
        }
    }
}

// The 12 MMAs of one 32-k chunk (four k-steps x {A_lo B_hi, A_hi B_lo, A_hi B_hi}) as ONE instruction sequence behind a single
// elect.sync: at = TMEM address of the A stage ([hi 32 | lo 32] columns), b0 = descriptor of the chunk's B_hi image (B_lo is
// 16 KB = 1024 descriptor units behind, a k-step is 4 KB = 256 units).  keep_first = 0 overwrites the accumulator.
// (One elect.sync + setp per MMA made the issuing warp, not the tensor core, the pace setter: ~98 cycles per MMA.)
__device__ __forceinline__ void umma_chunk_tf32x3(uint32_t acc, uint32_t at, uint64_t b0, uint32_t keep_first) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, pe, pt;\n\t"
        ".reg .b32 ah, al;\n\t"
        ".reg .b64 bh, bl;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 pt, %3, %3;\n\t"
        "add.u32 al, %1, 32;\n\t"
        "add.u64 bl, %2, 1024;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], %2, %3, p;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bl, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, pt;\n\t"
        "add.u32 ah, %1, 8;\n\t"
        "add.u32 al, %1, 40;\n\t"
        "add.u64 bh, %2, 256;\n\t"
        "add.u64 bl, %2, 1280;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %3, pt;\n\t"
        "add.u32 ah, %1, 16;\n\t"
        "add.u32 al, %1, 48;\n\t"
        "add.u64 bh, %2, 512;\n\t"
        "add.u64 bl, %2, 1536;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %3, pt;\n\t"
        "add.u32 ah, %1, 24;\n\t"
        "add.u32 al, %1, 56;\n\t"
        "add.u64 bh, %2, 768;\n\t"
        "add.u64 bl, %2, 1792;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %3, pt;\n\t"
        "}\n" ::"r"(acc),
        "r"(at), "l"(b0), "r"(ftc::IDESC), "r"(keep_first)
        : "memory");
}

// [128 k][128 n] row-major (element (k, n) at k * 128 + n) -> packed [32 k4][128 n][4]
__global__ void k_pack_kn(const float *__restrict__ src, float *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 16384) return;
    const int k = i >> 7, n = i & 127;
    dst[((size_t)(k >> 2) * 128 + n) * 4 + (k & 3)] = src[i];
}

#ifdef INFGEN_NTC_TRACE
__device__ long long g_ntc_trace[2][64];          // clock64 stamps of [0: row thread 32, 1: MMA lane 0] of CTA 0
#define NTC_STAMP(s) do { if (blockIdx.x == 0 && a.w_post && a.w_pre && a.pre_kv && trace_n < 64) g_ntc_trace[s][trace_n++] = clock64(); } while (0)
#else
#define NTC_STAMP(s) do {} while (0)
#endif

__global__ void __launch_bounds__(ntc::THREADS, 1) k_node_tc(const NodeArgs a) {
    using namespace ntc;
    extern __shared__ __align__(128) float smem_ntc[];
    float *smem = smem_ntc;
    float *sB = smem + SM_B, *sex = smem + SM_EX;
    int *s_row = reinterpret_cast<int *>(smem + SM_ROW);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *full_a = bars, *empty_a = full_a + 4, *full_b = empty_a + 4, *empty_b = full_b + NB, *acc_done = empty_b + NB,
             *acc_free = acc_done + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + SM_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef INFGEN_NTC_TRACE
    int trace_n = 0;
#endif

    // ---- the 128 active rows of this tile ---------------------------------------------------------------------------------
    int grow = -1;
    if (tid < TM) {
        int g = blockIdx.x * TM + tid;
        if (a.rows.list) {
            grow = g < *a.rows.n_list ? a.rows.list[g] : -1;
        } else if (a.rows.cap == 0) {
            grow = g < a.rows.n_total ? g : -1;
        } else {
            const int ns = a.rows.n_total / a.rows.cap;
            for (int b = 0; b < ns; ++b) {
                const int lo = a.rows.row_lo ? max(a.rows.row_lo[b], 0) : 0;
                const int c = max(a.rows.n_rows[b] - lo, 0);
                if (g < c) { grow = b * a.rows.cap + lo + g; break; }
                g -= c;
            }
        }
        s_row[tid] = grow;
    }
    if (!__syncthreads_or(grow >= 0)) return;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&full_a[i], 4); mbar_init(&empty_a[i], 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_done[i], 1); mbar_init(&acc_free[i], RT / 32); }
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    const bool post = a.w_post != nullptr, pre = a.w_pre != nullptr;
    const bool edgeless = a.edgeless != 0, pre_kv = a.pre_kv != 0;
    const bool fold = pre && a.pw.has_pos && !edgeless;
    uint32_t J[MAX_JOBS];
    const int n_jobs = job_list(post, pre, pre_kv, edgeless, fold, J);

    if (warp == RT / 32 + 1) {
        // ---- weight producer ------------------------------------------------------------------------------------------------
        int ci = 0;
        for (int jb = 0; jb < n_jobs; ++jb) {
            const uint32_t code = J[jb];
            const float *img = ((code >> 5) & 1u) ? a.tc_pre : a.tc_post;
            const float *src = img + (size_t)(code & 31u) * 4 * CHUNK;
            for (int c = 0; c < 4; ++c, ++ci) {
                const int st = ci % NB, use = ci / NB;
                if (lane == 0) {
                    if (use > 0) ntc_wait(&empty_b[st], (uint32_t)(use - 1) & 1u, 100);
                    mbar_expect_tx(&full_b[st], CHUNK * 4u);
                    bulk_g2s(sB + st * CHUNK, src + (size_t)c * CHUNK, CHUNK * 4u, &full_b[st]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 0) {
        // ---- MMA issuer -----------------------------------------------------------------------------------------------------
        int ci = 0, n_fresh = 0, n_done[2] = {0, 0};
        if (lane == 0) NTC_STAMP(1);
        for (int jb = 0; jb < n_jobs; ++jb) {
            const uint32_t code = J[jb];
            const int ab = (code >> 6) & 1u;
            const bool accum = (code >> 7) & 1u, fresh = (code >> 8) & 1u, release = (code >> 9) & 1u, done = (code >> 10) & 1u;
            const bool is_fold = (code >> 11) & 1u;
            if (!is_fold && !accum && n_done[ab] > 0) ntc_wait(&acc_free[ab], (uint32_t)(n_done[ab] - 1) & 1u, 600);
            for (int c = 0; c < 4; ++c, ++ci) {
                const int sb = ci % NB;
                if (fresh) ntc_wait(&full_a[c], (uint32_t)n_fresh & 1u, 200);
                ntc_wait(&full_b[sb], (uint32_t)(ci / NB) & 1u, 300);
                tc_fence_after();
                const uint32_t at = tmem + TC_A + 64u * (uint32_t)c;
                const uint64_t b0 = umma_desc(smem_u32(sB + sb * CHUNK));
                if (!is_fold) {
                    umma_chunk_tf32x3(tmem + (ab ? ACC1 : ACC0), at, b0, (accum ? 1u : 0u) | (c > 0));
                } else {
                    // relative-query fold: heads 2c, 2c + 1 are the two 16-k halves of this chunk, each with its own accumulator
                    for (int hh = 0; hh < 2; ++hh) {
                        const int hb = (ab + 2 * c + hh) & 1;
                        if (n_done[hb] > 0) ntc_wait(&acc_free[hb], (uint32_t)(n_done[hb] - 1) & 1u, 600);
                        tc_fence_after();
                        const uint32_t acc = tmem + (hb ? ACC1 : ACC0);
#pragma unroll
                        for (int k2 = 0; k2 < 2; ++k2) {
                            const int ks = 2 * hh + k2;
                            const uint32_t ah = at + 8u * ks, al = at + 32u + 8u * ks;
                            const uint64_t bh = b0 + (uint64_t)(ks * (4096 >> 4)), bl = bh + (16384 >> 4);
                            umma_tf32_ts(acc, al, bh, k2 > 0);
                            umma_tf32_ts(acc, ah, bl, 1u);
                            umma_tf32_ts(acc, ah, bh, 1u);
                        }
                        umma_commit_elect(&acc_done[hb]);
                        ++n_done[hb];
                    }
                }
                if (release) umma_commit_elect(&empty_a[c]);
                umma_commit_elect(&empty_b[sb]);
                if (c == 3 && done && !is_fold) umma_commit_elect(&acc_done[ab]);
            }
            if (done && !is_fold) ++n_done[ab];
            if (fresh) ++n_fresh;
            if (lane == 0) NTC_STAMP(1);
        }
    } else {
        // ---- row threads ----------------------------------------------------------------------------------------------------
        const int qd = (warp - 1) >> 2, r = 32 * (warp & 3) + lane;
        const int cb = 32 * qd;                                        // first of this thread's 32 columns
        const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t ta = lane_base + TC_A + 64u * (uint32_t)qd;     // this thread's A stage
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // Global rows <-> registers.  In the TMEM layout a thread owns 32 consecutive columns of ONE row, so a direct
        // LDG.128 / STG.128 of a warp touches 32 different cache lines (32 L1 wavefronts per instruction: measured 4-5 k
        // cycles per [128 x 128] transfer, the whole kernel 2.5x slower than the FFMA one).  Every transfer therefore goes
        // through a per-warp [32 rows][32 + 4 columns] staging tile: eight lanes move one 128-byte row segment.
        float *tile = smem + SM_TILE + (warp - 1) * TILE_FLOATS;
        int grow8[8];                                                  // global rows of the coalesced side: row 4 it + lane / 8
#pragma unroll
        for (int it = 0; it < 8; ++it) grow8[it] = s_row[32 * (warp & 3) + 4 * it + (lane >> 3)];
        const int c4 = lane & 7;
        auto g2tile = [&](const float *p, size_t mul, size_t add) {     // tile <- rows of p (row g at p + g * mul + add)
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int g = grow8[it];
                const float4 v = g >= 0 ? ld4(p + (size_t)g * mul + add + cb + 4 * c4) : z4;
                st4(tile + (4 * it + (lane >> 3)) * TILE_LD + 4 * c4, v);
            }
            __syncwarp();
        };
        auto tile2g = [&](float *p, size_t mul, size_t add) {
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int g = grow8[it];
                if (g >= 0) st4(p + (size_t)g * mul + add + cb + 4 * c4, ld4(tile + (4 * it + (lane >> 3)) * TILE_LD + 4 * c4));
            }
        };
        auto tile2reg = [&](float *val) {                              // this thread's row of the tile
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 v = ld4(tile + lane * TILE_LD + 4 * i4);
                val[4 * i4] = v.x; val[4 * i4 + 1] = v.y; val[4 * i4 + 2] = v.z; val[4 * i4 + 3] = v.w;
            }
        };
        auto reg2tile = [&](const float *val) {
            __syncwarp();
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4)
                st4(tile + lane * TILE_LD + 4 * i4, make_float4(val[4 * i4], val[4 * i4 + 1], val[4 * i4 + 2], val[4 * i4 + 3]));
        };
        auto load_rows = [&](const float *p, size_t mul, size_t add, float *val) { g2tile(p, mul, add); tile2reg(val); };
        auto store_rows = [&](float *p, size_t mul, size_t add, const float *val) { reg2tile(val); tile2g(p, mul, add); };
        auto rt_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory"); };
        auto warp_wait = [&](uint64_t *b, uint32_t parity, int code) {
            if (lane == 0) ntc_wait<ntc::NTC_ROW_BACKOFF>(b, parity, code);
            __syncwarp();
            tc_fence_after();
        };
        int n_put = 0, n_seen[2] = {0, 0};
        // this thread's 32 k of the row -> [hi | lo] columns of A stage qd
        auto put = [&](const float *val) {
            if (n_put > 0) warp_wait(&empty_a[qd], (uint32_t)(n_put - 1) & 1u, 400);
            ++n_put;
            float t[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __uint_as_float((__float_as_uint(val[i]) + 0x1000u) & 0xFFFFE000u);
            tmem_st32(ta, t);
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = val[i] - t[i];
            tmem_st32(ta + 32u, t);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[qd]);
            if (tid == 32) NTC_STAMP(0);
        };
        // result of the next finished job of accumulator ab -> val; the accumulator is free again afterwards
        auto take = [&](int ab, float *val) {
            warp_wait(&acc_done[ab], (uint32_t)n_seen[ab] & 1u, 500);
            ++n_seen[ab];
            if (tid == 32) NTC_STAMP(0);
            tmem_ld32(lane_base + (ab ? ACC1 : ACC0) + (uint32_t)cb, val);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[ab]);
        };
        int exb = 0;
        auto row_sum = [&](float part) {
            float *e = sex + exb * 512;
            exb ^= 1;
            e[qd * 128 + r] = part;
            rt_sync();
            return (e[r] + e[128 + r]) + (e[256 + r] + e[384 + r]);
        };
        auto row_stats = [&](const float *val, float &mean, float &rstd) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) s += val[i];
            mean = row_sum(s) * (1.0f / HID);
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float c = val[i] - mean; q = fmaf(c, c, q); }
            rstd = 1.0f / sqrtf(row_sum(q) * (1.0f / HID) + LN_EPS);
        };
        // val = (val - mean) * rstd * g + b
        auto affine = [&](float *val, float mean, float rstd, const float *g, const float *b) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 g4 = ldg4(g + cb + 4 * i4), b4 = ldg4(b + cb + 4 * i4);
                val[4 * i4 + 0] = (val[4 * i4 + 0] - mean) * rstd * g4.x + b4.x;
                val[4 * i4 + 1] = (val[4 * i4 + 1] - mean) * rstd * g4.y + b4.y;
                val[4 * i4 + 2] = (val[4 * i4 + 2] - mean) * rstd * g4.z + b4.z;
                val[4 * i4 + 3] = (val[4 * i4 + 3] - mean) * rstd * g4.w + b4.w;
            }
        };
        auto add_bias = [&](float *val, const float *b) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 b4 = ldg4(b + cb + 4 * i4);
                val[4 * i4 + 0] += b4.x; val[4 * i4 + 1] += b4.y; val[4 * i4 + 2] += b4.z; val[4 * i4 + 3] += b4.w;
            }
        };
        // val += this thread's row of the tile
        auto add_tile = [&](float *val) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 v = ld4(tile + lane * TILE_LD + 4 * i4);
                val[4 * i4 + 0] += v.x; val[4 * i4 + 1] += v.y; val[4 * i4 + 2] += v.z; val[4 * i4 + 3] += v.w;
            }
        };
        float val[32];
        float mean, rstd;
        if (tid == 32) NTC_STAMP(0);
        if (post) {
            // ---- gate ---------------------------------------------------------------------------------------------------------
            if (!edgeless) {
                load_rows(a.agg, 128, 0, val);
                put(val);
            }
            load_rows(a.x, 128, 0, val);
            row_stats(val, mean, rstd);
            affine(val, mean, rstd, a.lw.ln_dst_g, a.lw.ln_dst_b);
            put(val);
            {
                float sv[32];
                load_rows(a.s, 128, 0, sv);                            // (in flight while the gate GEMM runs)
                if (!edgeless) g2tile(a.agg, 128, 0);
                take(0, val);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 bg = ldg4(a.lw.b_g + cb + 4 * i4);
                    const float4 ag = edgeless ? z4 : ld4(tile + lane * TILE_LD + 4 * i4);
                    val[4 * i4 + 0] = ag.x + sigmoidf(val[4 * i4 + 0] + bg.x) * (sv[4 * i4 + 0] - ag.x);
                    val[4 * i4 + 1] = ag.y + sigmoidf(val[4 * i4 + 1] + bg.y) * (sv[4 * i4 + 1] - ag.y);
                    val[4 * i4 + 2] = ag.z + sigmoidf(val[4 * i4 + 2] + bg.z) * (sv[4 * i4 + 2] - ag.z);
                    val[4 * i4 + 3] = ag.w + sigmoidf(val[4 * i4 + 3] + bg.w) * (sv[4 * i4 + 3] - ag.w);
                }
            }
            put(val);
            // ---- to_out, x1 = x + LN_post(o) ---------------------------------------------------------------------------------
            g2tile(a.x, 128, 0);                                       // (in flight while to_out runs)
            take(1, val);
            add_bias(val, a.lw.b_out);
            row_stats(val, mean, rstd);
            affine(val, mean, rstd, a.lw.ln_post_g, a.lw.ln_post_b);
            add_tile(val);
            reg2tile(val);                                             // x1 stays in the staging tile during the FFN
            __syncwarp();
            float mean1, rstd1;
            row_stats(val, mean1, rstd1);
            // ---- FFN ------------------------------------------------------------------------------------------------------------
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                if (j > 0) tile2reg(val);
                affine(val, mean1, rstd1, a.lw.ln_ffpre_g, a.lw.ln_ffpre_b);
                put(val);
                take(0, val);
                const float *b1 = a.lw.b_ff1 + 128 * j;
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 b4 = ldg4(b1 + cb + 4 * i4);
                    val[4 * i4 + 0] = fmaxf(val[4 * i4 + 0] + b4.x, 0.f); val[4 * i4 + 1] = fmaxf(val[4 * i4 + 1] + b4.y, 0.f);
                    val[4 * i4 + 2] = fmaxf(val[4 * i4 + 2] + b4.z, 0.f); val[4 * i4 + 3] = fmaxf(val[4 * i4 + 3] + b4.w, 0.f);
                }
                put(val);
            }
            take(1, val);
            add_bias(val, a.lw.b_ff2);
            row_stats(val, mean, rstd);
            affine(val, mean, rstd, a.lw.ln_ffpost_g, a.lw.ln_ffpost_b);
            add_tile(val);                                             // x2 = x1 + LN_ffpost(y)
            reg2tile(val);
            tile2g(a.x, 128, 0);
            if (a.trace_out) tile2g(a.trace_out, 128, 0);
        } else {
            load_rows(a.x, 128, 0, val);
        }
        if (pre) {
            // ---- LayerNorm + s | k | v | q of the next layer --------------------------------------------------------------------
            row_stats(val, mean, rstd);
            affine(val, mean, rstd, a.pw.ln_dst_g, a.pw.ln_dst_b);
            put(val);
            int t = 0;
            take(t, val); t ^= 1;
            add_bias(val, a.pw.b_qs + 128);
            store_rows(a.s, 128, 0, val);
            if (pre_kv) {
                const int col = (a.col_ptr ? *a.col_ptr : 0) + a.col_add;
                const size_t mul = a.kv_ring ? (size_t)a.ring * 256 : 256;
                const size_t add = a.kv_ring ? (size_t)(col & (a.ring - 1)) * 256 : 0;
                take(t, val); t ^= 1;
                add_bias(val, a.pw.b_kv);
                store_rows(a.kv_out, mul, add, val);
                take(t, val); t ^= 1;
                add_bias(val, a.pw.b_kv + 128);
                store_rows(a.kv_out, mul, add + 128, val);
            }
            if (!edgeless) {
                take(t, val); t ^= 1;
                add_bias(val, a.pw.b_qs);
                store_rows(a.q, 128, 0, val);
                if (fold) {
                    put(val);
                    float g[32];
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const float4 g4 = ldg4(a.pw.ln_r_g + cb + 4 * i4);
                        g[4 * i4] = g4.x; g[4 * i4 + 1] = g4.y; g[4 * i4 + 2] = g4.z; g[4 * i4 + 3] = g4.w;
                    }
#pragma unroll 1
                    for (int h = 0; h < 8; ++h) {
                        take(t, val); t ^= 1;
#pragma unroll
                        for (int i = 0; i < 32; ++i) val[i] *= g[i];
                        store_rows(a.qr, 1024, 128 * h, val);
                    }
                }
            }
        }
        if (tid == 32) NTC_STAMP(0);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace infgen
