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

// one Linear (or a 128-column slice of one): `k4` packed rows, `ld` floats between consecutive k4 rows (512 when the
// matrix is exactly 128 columns wide, i.e. contiguous)
struct WSeg {
    const float *p;
    int k4, ld;
};

__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

struct WsSmem {
    float *ring;
    uint64_t *full, *empty;
    __device__ __forceinline__ explicit WsSmem(float *base)
        : ring(base), full(reinterpret_cast<uint64_t *>(base + WS_RING_FLOATS)),
          empty(reinterpret_cast<uint64_t *>(base + WS_RING_FLOATS) + WS_STAGES) {}
};

// all threads, before the roles split (contains __syncthreads)
__device__ __forceinline__ void ws_init(const WsSmem &ws) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < WS_STAGES; ++i) {
            mbar_init(&ws.full[i], 1);
            mbar_init(&ws.empty[i], NWARP);
        }
        fence_mbar_init();
    }
    __syncthreads();
}

// producers: lanes 0..WS_STAGES-1 of the producer warp; lane s owns ring stage s and issues every chunk that lands in
// it (the issue latency of one bulk copy, ~450 cycles from one thread, is spread over the lanes)
__device__ __forceinline__ void ws_produce(const WsSmem &ws, const WSeg *segs, int n_seg) {
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
                if (sg.ld == 512) {
                    bulk_g2s(dst, sg.p + (size_t)r0 * 512, (uint32_t)rows * 2048u, &ws.full[stage]);
                } else {
                    for (int i = 0; i < rows; ++i)
                        bulk_g2s(dst + i * 512, sg.p + (size_t)(r0 + i) * sg.ld, 2048u, &ws.full[stage]);
                }
            }
            if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
        }
    }
}

// consumer-side cursor (same walk as the producer)
struct WsCons {
    float *ring;
    uint64_t *full, *empty;
    int stage;
    uint32_t phase;
    __device__ __forceinline__ explicit WsCons(const WsSmem &ws) : ring(ws.ring), full(ws.full), empty(ws.empty), stage(0), phase(0) {}
    __device__ __forceinline__ const float *acquire() {
        mbar_wait(&full[stage], phase);
        return ring + stage * WS_STAGE_FLOATS;
    }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stage]);
        if (++stage == WS_STAGES) { stage = 0; phase ^= 1u; }
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Y[m][n] = sum_k X[m][k] W[k][n] for a tile of M rows (8, 16 or 32) and 128 columns; W streamed (next segment of the
// ring).  Warp w owns columns [16w, 16w+16) - every weight element is read from shared memory by exactly one warp -
// and inside the warp a thread owns NTT = M/8 columns {16w + cl + CL*j} and 4 rows {rl + RL*i}, CL = 16/NTT column
// lanes, RL = 32/CL row lanes.  X: shared, row-major, leading dimension ldx; ldx % 32 == 4 keeps the row-lane loads
// bank-conflict free.  epi(m, n, value) once per output.  No barrier inside: the caller csync()s before.
// ---------------------------------------------------------------------------------------------------------------------
template <int M, typename Epi>
__device__ __forceinline__ void stream_gemm(WsCons &ws, const float *xs, int ldx, int K4, Epi epi) {
    static_assert(M == 8 || M == 16 || M == 32, "tile rows");
    constexpr int NTT = M / 8, CLN = 16 / NTT, RLN = 32 / CLN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
