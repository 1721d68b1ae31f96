// Weight streaming for the row-tile kernels (Fourier / MLP embeddings, heads): a producer warp brings the K-major packed
// weights [K4][N][4] of every Linear of the kernel, in order, through a shared-memory ring with cp.async.bulk (TMA bulk
// copies, completion on mbarriers); the eight consumer warps run register-tiled FFMA GEMMs straight out of the ring.
// The producer runs ahead across GEMM boundaries, so the first rows of the next Linear are already resident while the
// LayerNorm between two Linears executes.
//
//   CTA = NT_S = 288 threads: warps 0..7 consume, lanes 0..WS_STAGES-1 of warp 8 produce (one ring stage each).
//   ring stage = WS_ROWS k4-rows x 128 columns x 4 floats = 16 KB; WS_STAGES stages.
//   consumers synchronise among themselves with named barrier 1 (csync); __syncthreads() is only legal before the roles
//   split.
#pragma once
#include "common.cuh"

namespace infgen {

constexpr int WS_STAGES = 3;
constexpr int WS_ROWS = 8;                               // k4 rows (= 32 k) per stage
constexpr int WS_STAGE_FLOATS = WS_ROWS * 128 * 4;       // 4096 floats = 16 KB
constexpr int WS_RING_FLOATS = WS_STAGES * WS_STAGE_FLOATS;
constexpr int WS_SMEM_FLOATS = WS_RING_FLOATS + 4 * WS_STAGES;   // + full/empty mbarriers (uint64 each)
constexpr int NT_S = NT + 32;
constexpr int WS_MAX_SEGS = 12;

#ifdef INFGEN_WS_TRACE
// debug: clock64 of every chunk issue (producer) / acquire completion (consumer warp 0) of CTA 0 of the last kernel
__device__ long long g_ws_trace[2][256];
__device__ int g_ws_trace_n[2];
#ifndef INFGEN_WS_TRACE_SMEM
#define INFGEN_WS_TRACE_SMEM 0u          // trace only the kernel launched with this much dynamic shared memory (0: all)
#endif
__device__ __forceinline__ bool ws_trace_on() {
    unsigned v;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(v));
    return INFGEN_WS_TRACE_SMEM == 0u || v == INFGEN_WS_TRACE_SMEM;
}
#define WS_TRACE(which)                                                                                       \
    do {                                                                                                      \
        if (blockIdx.x == 0 && blockIdx.y == 0 && ws_trace_on()) {                                                             \
            const int _i = g_ws_trace_n[which]++;                                                             \
            if (_i < 256) g_ws_trace[which][_i] = clock64();                                                  \
        }                                                                                                     \
    } while (0)
#else
#define WS_TRACE(which) do {} while (0)
#endif

// one Linear (or a 128-column slice of one): `k4` packed rows, `ld` floats between consecutive k4 rows (512 when the
// matrix is exactly 128 columns wide, i.e. contiguous)
struct WSeg {
    const float *p;
    int k4, ld;
};

__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// ring depth S.  Measured (tools/probe/ws_trace.py, k_embed_column; bench with 3 vs 6 stages): the consumers, not the ring,
// bound these kernels - the producer runs a full ring ahead and a stage is consumed in ~1.28 k cycles at M = 8 (40 LDS.128
// per warp and stage x 4 wavefronts x 8 warps), ~1.5 k at M = 16 - so 3 stages are enough and 6 change nothing.
template <int S>
__host__ __device__ constexpr int ws_smem_floats() { return S * WS_STAGE_FLOATS + 4 * S; }
template <int S = WS_STAGES>
struct WsSmemT {
    float *ring;
    uint64_t *full, *empty;
    __device__ __forceinline__ explicit WsSmemT(float *base)
        : ring(base), full(reinterpret_cast<uint64_t *>(base + S * WS_STAGE_FLOATS)),
          empty(reinterpret_cast<uint64_t *>(base + S * WS_STAGE_FLOATS) + S) {}
};
using WsSmem = WsSmemT<WS_STAGES>;

// all threads, before the roles split (contains __syncthreads)
template <int S>
__device__ __forceinline__ void ws_init(const WsSmemT<S> &ws) {
#ifdef INFGEN_WS_TRACE
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && ws_trace_on()) { g_ws_trace_n[0] = 0; g_ws_trace_n[1] = 0; }
#endif
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&ws.full[i], 1);
            mbar_init(&ws.empty[i], NWARP);
        }
        fence_mbar_init();
    }
    __syncthreads();
}

// producers: lanes 0..WS_STAGES-1 of the producer warp; lane s owns ring stage s and issues every chunk that lands in
// it (the issue latency of one bulk copy, ~450 cycles from one thread, is spread over the lanes)
template <int S>
__device__ __forceinline__ void ws_produce(const WsSmemT<S> &ws, const WSeg *segs, int n_seg) {
    const int my_stage = threadIdx.x & 31;
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < n_seg; ++s) {
        const WSeg sg = segs[s];
        for (int r0 = 0; r0 < sg.k4; r0 += WS_ROWS) {
            if (stage == my_stage) {
                const int rows = min(WS_ROWS, sg.k4 - r0);
                mbar_wait(&ws.empty[stage], phase ^ 1u);
                float *dst = ws.ring + stage * WS_STAGE_FLOATS;
                mbar_expect_tx(&ws.full[stage], (uint32_t)rows * 2048u);
                WS_TRACE(0);
                if (sg.ld == 512) {
                    bulk_g2s(dst, sg.p + (size_t)r0 * 512, (uint32_t)rows * 2048u, &ws.full[stage]);
                } else {
                    for (int i = 0; i < rows; ++i)
                        bulk_g2s(dst + i * 512, sg.p + (size_t)(r0 + i) * sg.ld, 2048u, &ws.full[stage]);
                }
            }
            if (++stage == S) { stage = 0; phase ^= 1u; }
        }
    }
}

// consumer-side cursor (same walk as the producer)
template <int S = WS_STAGES>
struct WsConsT {
    float *ring;
    uint64_t *full, *empty;
    int stage;
    uint32_t phase;
    __device__ __forceinline__ explicit WsConsT(const WsSmemT<S> &ws) : ring(ws.ring), full(ws.full), empty(ws.empty), stage(0), phase(0) {}
    __device__ __forceinline__ const float *acquire() {
        mbar_wait(&full[stage], phase);
        if (threadIdx.x == 0) WS_TRACE(1);
        return ring + stage * WS_STAGE_FLOATS;
    }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stage]);
        if (++stage == S) { stage = 0; phase ^= 1u; }
    }
};
using WsCons = WsConsT<WS_STAGES>;

// ---------------------------------------------------------------------------------------------------------------------
// Y[m][n] = sum_k X[m][k] W[k][n] for a tile of M rows (8, 16 or 32) and 128 columns; W streamed (next segment of the
// ring).  Warp w owns columns [16w, 16w+16) - every weight element is read from shared memory by exactly one warp -
// and inside the warp a thread owns NTT = M/8 columns {16w + cl + CL*j} and 4 rows {rl + RL*i}, CL = 16/NTT column
// lanes, RL = 32/CL row lanes.  X: shared, row-major, leading dimension ldx; ldx % 32 == 4 keeps the row-lane loads
// bank-conflict free.  epi(m, n, value) once per output.  No barrier inside: the caller csync()s before.
// ---------------------------------------------------------------------------------------------------------------------
// single (uniform over the CTA): only row 0 of the tile is live (the row an insertion pass appended).  The product is then
// a GEMV: lanes 0..15 of warp w own one column each and walk the k4 rows of every stage in the same order as the tile code
// (bitwise the same row 0), 16 instead of 40 shared-memory loads per warp and stage; epi runs for row 0 only.
template <int M, typename Cons, typename Epi>
__device__ __forceinline__ void stream_gemm(Cons &ws, const float *xs, int ldx, int K4, Epi epi, bool single = false) {
    static_assert(M == 8 || M == 16 || M == 32, "tile rows");
    constexpr int NTT = M / 8, CLN = 16 / NTT, RLN = 32 / CLN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (single) {
        const bool on = lane < 16;
        const int c = 16 * warp + (lane & 15);
        float a1 = 0.f;
        for (int kb = 0; kb < K4; kb += WS_ROWS) {
            const int rows = min(WS_ROWS, K4 - kb);
            const float *w = ws.acquire();
            if (on) {
                if (rows == WS_ROWS) {
                    float4 wv[WS_ROWS], xv[WS_ROWS];
#pragma unroll
                    for (int kk = 0; kk < WS_ROWS; ++kk) { wv[kk] = ld4(w + (kk * 128 + c) * 4); xv[kk] = ld4(xs + 4 * (kb + kk)); }
#pragma unroll
                    for (int kk = 0; kk < WS_ROWS; ++kk) {
                        a1 = fmaf(xv[kk].x, wv[kk].x, a1); a1 = fmaf(xv[kk].y, wv[kk].y, a1);
                        a1 = fmaf(xv[kk].z, wv[kk].z, a1); a1 = fmaf(xv[kk].w, wv[kk].w, a1);
                    }
                } else {
                    for (int kk = 0; kk < rows; ++kk) {
                        const float4 wv = ld4(w + (kk * 128 + c) * 4), xv = ld4(xs + 4 * (kb + kk));
                        a1 = fmaf(xv.x, wv.x, a1); a1 = fmaf(xv.y, wv.y, a1); a1 = fmaf(xv.z, wv.z, a1); a1 = fmaf(xv.w, wv.w, a1);
                    }
                }
            }
            ws.release();
        }
        if (on) epi(0, c, a1);
        return;
    }
    const int cl = lane % CLN, rl = lane / CLN;
    float acc[4][NTT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NTT; ++j) acc[i][j] = 0.f;
    const float *xr = xs + (size_t)rl * ldx;
    const int c0 = 16 * warp + cl;
    auto body = [&](const float *w, int kk, int k4) {
        float4 wv[NTT];
#pragma unroll
        for (int j = 0; j < NTT; ++j) wv[j] = ld4(w + (kk * 128 + c0 + CLN * j) * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 x = ld4(xr + (size_t)(RLN * i) * ldx + 4 * k4);
#pragma unroll
            for (int j = 0; j < NTT; ++j) {
                acc[i][j] = fmaf(x.x, wv[j].x, acc[i][j]);
                acc[i][j] = fmaf(x.y, wv[j].y, acc[i][j]);
                acc[i][j] = fmaf(x.z, wv[j].z, acc[i][j]);
                acc[i][j] = fmaf(x.w, wv[j].w, acc[i][j]);
            }
        }
    };
    for (int kb = 0; kb < K4; kb += WS_ROWS) {
        const int rows = min(WS_ROWS, K4 - kb);
        const float *w = ws.acquire();
        if (rows == WS_ROWS) {
#pragma unroll
            for (int kk = 0; kk < WS_ROWS; ++kk) body(w, kk, kb + kk);
        } else {
            for (int kk = 0; kk < rows; ++kk) body(w, kk, kb + kk);
        }
        ws.release();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NTT; ++j) epi(rl + RLN * i, c0 + CLN * j, acc[i][j]);
}

// ---------------------------------------------------------------------------------------------------------------------
// k-split variant for 16-row tiles.  The row-tile GEMMs are bound by shared-memory operand loads (an LDS.128 costs four
// wavefronts whatever its broadcast degree): the 4 x 2 register tile of stream_gemm<16> needs 6 LDS.128 per 32 FMA.  Here
// warps 0-3 take the even k4 rows of every stage and warps 4-7 the odd ones, each warp owning 32 columns, so a thread
// holds a 4 x 4 tile (8 LDS.128 per 64 FMA: 1,024 instead of 1,536 LSU cycles per 16 KB stage) and the two k-halves are
// added through `red` ([16][RED_LD] floats): each half finalises two of its four row groups.
// Contains one csync(); the caller csync()s before (X visible) as for stream_gemm and alternates `red` between two buffers.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RED_LD = 132;
template <typename Cons, typename Epi>
__device__ __forceinline__ void stream_gemm_ks16(Cons &ws, const float *xs, int ldx, int K4, float *red, Epi epi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kh = warp >> 2, c0 = 32 * (warp & 3) + (lane & 7), rl = lane >> 3;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float *xr = xs + (size_t)rl * ldx;
    for (int kb = 0; kb < K4; kb += WS_ROWS) {
        const float *w = ws.acquire();
        // operands of the next k4 row are loaded while the current one is multiplied (the kernel runs 2 warps per
        // scheduler: without the overlap neither the LSU nor the FMA pipe is half busy).  Fully unrolled: a rolled body
        // measured 74 instead of 65 us per k_node launch.
        float4 wv[2][4], xv[2][4];
        auto load = [&](int b, int kk2) {
            const int kk = 2 * kk2 + kh;
#pragma unroll
            for (int j = 0; j < 4; ++j) wv[b][j] = ld4(w + (kk * 128 + c0 + 8 * j) * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[b][i] = ld4(xr + (size_t)(4 * i) * ldx + 4 * (kb + kk));
        };
        load(0, 0);
#pragma unroll
        for (int kk2 = 0; kk2 < WS_ROWS / 2; ++kk2) {
            const int b = kk2 & 1;
            if (kk2 + 1 < WS_ROWS / 2) load(b ^ 1, kk2 + 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 x = xv[b][i];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(x.x, wv[b][j].x, acc[i][j]);
                    acc[i][j] = fmaf(x.y, wv[b][j].y, acc[i][j]);
                    acc[i][j] = fmaf(x.z, wv[b][j].z, acc[i][j]);
                    acc[i][j] = fmaf(x.w, wv[b][j].w, acc[i][j]);
                }
            }
        }
        ws.release();
    }
    // the half kh hands its partial sums of row groups {2, 3} (kh = 0) / {0, 1} (kh = 1) to the other half (the branches
    // keep every acc index a compile-time constant: registers, not local memory).  Callers alternate between two `red`
    // buffers from call to call, so these writes cannot overtake the epilogue reads of the previous call (a thread can be
    // at most one call ahead: the csync() below needs every thread).
    if (kh == 0) {
#pragma unroll
        for (int i = 2; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(rl + 4 * i) * RED_LD + c0 + 8 * j] = acc[i][j];
    } else {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(rl + 4 * i) * RED_LD + c0 + 8 * j] = acc[i][j];
    }
    csync();
    if (kh == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = rl + 4 * i, n = c0 + 8 * j;
                epi(m, n, acc[i][j] + red[m * RED_LD + n]);
            }
    } else {
#pragma unroll
        for (int i = 2; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = rl + 4 * i, n = c0 + 8 * j;
                epi(m, n, red[m * RED_LD + n] + acc[i][j]);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core variant of stream_gemm for M = 16 / 32 row tiles: mma.sync m16n8k8 TF32 with the 3xTF32 error-compensated
// split (x = hi + lo, both TF32; D += A_lo B_hi + A_hi B_lo + A_hi B_hi, fp32 accumulate), which keeps fp32-level
// accuracy (the dropped lo*lo term is 2^-22 relative) - the decode loop is closed: a plain TF32 product would flip
// near-tie token arg-maxes.  Operand fragments come straight from the activation tile and the weight ring with
// conflict-free 32-bit loads (12 per k-step of 8 for M = 32, against 8 x 128-bit per 4 k in the FFMA tile), which is
// what bounds these GEMMs.  Same contract as stream_gemm; warp w owns columns [16w, 16w+16).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int M, typename Cons, typename Epi>
__device__ __forceinline__ void stream_gemm_mma(Cons &ws, const float *xs, int ldx, int K4, Epi epi) {
    static_assert(M == 16 || M == 32, "tile rows");
    constexpr int MT = M / 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    float acc[MT][2][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
    const float *xa = xs + (size_t)g * ldx + t;
    const int nb = (16 * warp + g) * 4 + t;                // float offset of (n = 16w + g, element t) inside a k4 row
    auto kstep = [&](const float *w, int kk, int k4, bool full) {
        // B fragments: b0 = W[4*k4 + t][n], b1 = W[4*(k4+1) + t][n]
        uint32_t bh[2][2], bl[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float b0 = w[kk * 512 + nb + 32 * j];
            const float b1 = full ? w[(kk + 1) * 512 + nb + 32 * j] : 0.f;
            split_tf32(b0, bh[j][0], bl[j][0]);
            split_tf32(b1, bh[j][1], bl[j][1]);
        }
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const float *xr = xa + (size_t)(16 * i) * ldx + 4 * k4;
            const float a0 = xr[0], a1 = xr[8 * ldx];
            const float a2 = full ? xr[4] : 0.f, a3 = full ? xr[8 * ldx + 4] : 0.f;
            uint32_t ah[4], al[4];
            split_tf32(a0, ah[0], al[0]); split_tf32(a1, ah[1], al[1]);
            split_tf32(a2, ah[2], al[2]); split_tf32(a3, ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                mma_tf32(acc[i][j], al, bh[j]);
                mma_tf32(acc[i][j], ah, bl[j]);
                mma_tf32(acc[i][j], ah, bh[j]);
            }
        }
    };
    for (int kb = 0; kb < K4; kb += WS_ROWS) {
        const int rows = min(WS_ROWS, K4 - kb);
        const float *w = ws.acquire();
        if (rows == WS_ROWS) {
#pragma unroll
            for (int kk = 0; kk < WS_ROWS; kk += 2) kstep(w, kk, kb + kk, true);
        } else {
            for (int kk = 0; kk < rows; kk += 2) kstep(w, kk, kb + kk, kk + 1 < rows);
        }
        ws.release();
    }
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = 16 * i + g, n = 16 * warp + 8 * j + 2 * t;
            epi(m, n, acc[i][j][0]);
            epi(m, n + 1, acc[i][j][1]);
            epi(m + 8, n, acc[i][j][2]);
            epi(m + 8, n + 1, acc[i][j][3]);
        }
}

// dispatcher.  The FFMA tile is the default: on B200 the legacy mma.sync TF32 path measured ~170 MAC/clk/SM (k_embed_column,
// 1150 cycles per 16 KB weight stage at M = 16, tools/probe/ws_trace.py), i.e. after the 3x split it is no faster than
// FFMA (128 MAC/clk/SM) - only tcgen05 would be, and that needs 64/128-row tiles a single scene does not have.
// Build with -DINFGEN_MMA to use the tensor-core variant for the 16/32-row tiles (parity-tested, same tolerances).
template <int M, typename Cons, typename Epi>
__device__ __forceinline__ void tile_gemm(Cons &ws, const float *xs, int ldx, int K4, Epi epi, bool single = false) {
#ifdef INFGEN_MMA
    if constexpr (M == 16 || M == 32) {
        if (!single) {
            stream_gemm_mma<M>(ws, xs, ldx, K4, epi);
            return;
        }
    }
#endif
    stream_gemm<M>(ws, xs, ldx, K4, epi, single);
}

// In-place LayerNorm (+ optional ReLU) of M rows, one warp per row; callers csync() before and after
template <int M, bool RELU>
__device__ __forceinline__ void rows_layernorm_c(float *s, int ld, const float *__restrict__ g,
                                                 const float *__restrict__ b) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int m = warp; m < M; m += NWARP) {
        float4 v = ld4(s + m * ld + 4 * lane);
        v = ln128(v, g, b, lane);
        if (RELU) v = relu4(v);
        st4(s + m * ld + 4 * lane, v);
    }
}

}  // namespace infgen
