// Device-side building blocks shared by all kernels of the decode path (sm_100a).
//
// Everything here is fp32: the reference runs fp32 end to end and greedy token argmax has
// to survive a closed loop, so the GEMMs are FFMA register-tile GEMMs with weights streamed K-major from L2
// (layout [K/4][N][4], one coalesced 16-byte load per thread per 4 k) and activations broadcast from shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace infgen {

constexpr int HID = 128;        // hidden_dim
constexpr int NHEAD = 8;        // num_heads
constexpr int HDIM = 16;        // head_dim
constexpr int NT = 256;         // threads per CTA of every node/edge GEMM kernel
constexpr int NWARP = NT / 32;
constexpr float LN_EPS = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// LayerNorm statistics of a 128-vector held four consecutive channels per lane (torch.nn.LayerNorm, biased var).
__device__ __forceinline__ void ln_stats(const float4 v, float &mean, float &rstd) {
    float s = warp_sum((v.x + v.y) + (v.z + v.w));
    mean = s * (1.0f / HID);
    float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    float q = warp_sum((a * a + b * b) + (c * c + d * d));
    rstd = 1.0f / sqrtf(q * (1.0f / HID) + LN_EPS);
}
__device__ __forceinline__ float4 ln_apply(const float4 v, float mean, float rstd, const float *__restrict__ g,
                                           const float *__restrict__ b, int lane) {
    float4 gg = ldg4(g + 4 * lane), bb = ldg4(b + 4 * lane);
    float4 o;
    o.x = (v.x - mean) * rstd * gg.x + bb.x;
    o.y = (v.y - mean) * rstd * gg.y + bb.y;
    o.z = (v.z - mean) * rstd * gg.z + bb.z;
    o.w = (v.w - mean) * rstd * gg.w + bb.w;
    return o;
}
__device__ __forceinline__ float4 ln128(const float4 v, const float *__restrict__ g, const float *__restrict__ b,
                                        int lane) {
    float mean, rstd;
    ln_stats(v, mean, rstd);
    return ln_apply(v, mean, rstd, g, b, lane);
}
__device__ __forceinline__ float4 relu4(float4 v) {
    return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// In-place LayerNorm (+ optional ReLU) of M rows of 128 floats in shared memory; warp w owns rows w, w+8, ...
template <int M, bool RELU>
__device__ __forceinline__ void rows_layernorm(float *s, int ld, const float *__restrict__ g,
                                               const float *__restrict__ b) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int m = warp; m < M; m += NWARP) {
        float4 v = ld4(s + m * ld + 4 * lane);
        v = ln128(v, g, b, lane);
        if (RELU) v = relu4(v);
        st4(s + m * ld + 4 * lane, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Block GEMM:  Y[m][n] = sum_k X[m][k] * W[k][n],  m < M, k < 4*K4, n < N.
//   X   shared memory, row-major, leading dimension ldx (floats, multiple of 4, 16-byte aligned rows)
//   wp  global memory, packed [K4][ldw][4] (w.x..w.w = W[4*k4 .. 4*k4+3][n]), zero padded in k; ldw >= N is the
//       full output width of the packed matrix so a kernel can work on a column slice (wp pre-offset by 4*n0)
//   N   128 (two k-halves per column, reduced through `red`), 256 (one column per thread) or 512 (two columns)
//   epi(m, n, value) is called exactly once per output element by the owning thread.
//   `red` : shared scratch of M*128 floats, only touched when N == 128.  Contains __syncthreads() when N == 128.
// All NT threads must call it.  The caller provides the barrier that makes X visible beforehand.
// ---------------------------------------------------------------------------------------------------------------
template <int M, int N, typename Epi>
__device__ __forceinline__ void block_gemm(const float *xs, int ldx, const float *__restrict__ wp, int ldw, int K4,
                                           float *red, Epi epi) {
    static_assert(N == 128 || N == 256 || N == 512, "unsupported N");
    constexpr int CPT = (N == 512) ? 2 : 1;
    constexpr int KS = (N == 128) ? 2 : 1;
    const int tid = threadIdx.x;
    const int col = (N == 128) ? (tid & 127) : tid;
    const int kh = (N == 128) ? (tid >> 7) : 0;
    const int kper = (K4 + KS - 1) / KS;
    const int k_begin = kh * kper;
    const int k_end = min(K4, k_begin + kper);
    float acc[M][CPT];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[m][c] = 0.f;
    const float4 *w4 = reinterpret_cast<const float4 *>(wp);
#pragma unroll 4
    for (int k4 = k_begin; k4 < k_end; ++k4) {
        float4 w[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) w[c] = __ldg(w4 + (size_t)k4 * ldw + col + c * 256);
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const float4 x = ld4(xs + m * ldx + 4 * k4);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                acc[m][c] = fmaf(x.x, w[c].x, acc[m][c]);
                acc[m][c] = fmaf(x.y, w[c].y, acc[m][c]);
                acc[m][c] = fmaf(x.z, w[c].z, acc[m][c]);
                acc[m][c] = fmaf(x.w, w[c].w, acc[m][c]);
            }
        }
    }
    if (N == 128) {
        if (kh == 1) {
#pragma unroll
            for (int m = 0; m < M; ++m) red[m * 128 + col] = acc[m][0];
        }
        __syncthreads();
        if (kh == 0) {
#pragma unroll
            for (int m = 0; m < M; ++m) epi(m, col, acc[m][0] + red[m * 128 + col]);
        }
    } else {
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int c = 0; c < CPT; ++c) epi(m, col + c * 256, acc[m][c]);
    }
}

// ---- mbarrier / bulk-copy primitives ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// one thread: request `floats` floats from global into shared memory, completion on `b`
__device__ __forceinline__ void chunk_request(float *dst, const float *src, int floats, uint64_t *b) {
    fence_proxy_async();
    uint32_t bytes = (uint32_t)floats * 4u;
    mbar_expect_tx(b, bytes);
    const char *s = reinterpret_cast<const char *>(src);
    char *d = reinterpret_cast<char *>(dst);
    while (bytes) {
        const uint32_t n = bytes < 32768u ? bytes : 32768u;
        bulk_g2s(d, s, n, b);
        d += n; s += n; bytes -= n;
    }
}

// accurate logistic (the reference uses torch.sigmoid in fp32)
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// infgen/utils/func.py:58-62 wrap_angle: -pi + (a + pi) % (2 pi) with python-style modulo, fp32 constants
__device__ __forceinline__ float wrap_angle(float a) {
    const float pi = 3.14159265358979323846f;
    const float two_pi = 6.28318530717958647692f;
    float r = fmodf(__fadd_rn(a, pi), two_pi);
    if (r != 0.f && r < 0.f) r = __fadd_rn(r, two_pi);
    return __fadd_rn(-pi, r);
}
// infgen/utils/func.py:30-34 angle_between_2d_vectors(ctr, nbr)
__device__ __forceinline__ float angle_between(float cx, float cy, float nx, float ny) {
    float cross = __fsub_rn(__fmul_rn(cx, ny), __fmul_rn(cy, nx));
    // the reference takes the dot product as `(ctr * nbr).sum(-1)`, a reduction that starts from +0: for a zero
    // neighbour vector (the ego -> seed edge of every insertion pass: the query row sits on the ego, agent_decoder.py:
    // 1796-1797) and a heading in the third quadrant both products are -0 and the sum is +0, not -0 - and
    // atan2(+0, -0) would be pi instead of 0
    float dot = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(cx, nx)), __fmul_rn(cy, ny));
    return atan2f(cross, dot);
}
__device__ __forceinline__ float norm2(float x, float y) {
    return sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}
__device__ __forceinline__ float dist2(float x, float y) { return __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)); }

}  // namespace infgen
