// FourierEmbedding (reference layers.py:142-160) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// 3xTF32 error-compensated split, accumulators in TMEM.
//
// One CTA embeds a tile of 128 edge slots.  Every Linear of the module is a [128 slots x 128] x [128 x 128] GEMM:
//     per input dim d:  G1_d = [cos(2 pi f x_d) | sin(2 pi f x_d)] W0_d[0:128]     (+ x_d W0_d[128] + b0_d in the epilogue)
//                       G2_d = relu(LN(G1_d)) W3_d, summed over d IN TMEM (the accumulator is simply not cleared)
//     output:           G3   = relu(LN(sum_d G2_d + b3_d)) Wout + bout             (optionally standardised -> rhat)
//
// fp32 fidelity: the reference is fp32 and the greedy argmax of a closed loop has to survive, so every operand is split
// x = hi + lo (both TF32, cvt.rna) and D += A_lo B_hi + A_hi B_lo + A_hi B_hi (fp32 accumulate); measured error of this
// scheme on B200: 1.4e-6 against fp64 for K = 64, |D| <= 2.3 (tools/probe/umma_probe.cu), the same as an fp32 FMA chain.
//
// Operand layout (measured in the probe): NO-swizzle K-major core matrices, element (row r, k) of a [128 x 32] chunk at
// float offset (k / 4) * 512 + r * 4 + (k % 4), i.e. LBO (K direction) = 2048 B, SBO (8-row groups) = 128 B.  For the
// weights this is exactly the packed [K/4][128][4] layout of the blob, so the B images are the blob matrices split into
// hi / lo and cut into 32-k chunks of [hi 16 KB | lo 16 KB], stored in consumption order: one 32 KB cp.async.bulk each.
//
// Roles (320 threads):
//   warps 0-7  row threads: thread (r = 32 (w & 3) + lane, half h = w >> 2) owns slot r and 64 of its 128 columns (TMEM
//              lanes are only visible to warps with the same w & 3).  They generate the Fourier features and run the
//              epilogues (TMEM -> registers -> bias / LayerNorm / ReLU -> hi/lo A chunks in the A ring).
//   warp 8     one lane issues the MMAs (12 per chunk) and commits to the ring / accumulator mbarriers.
//   warp 9     streams the weight chunks through the B ring (one lane per stage).
// G1_{d+1} is issued before the epilogue of G1_d (two G1 accumulators), so the tensor core works while the row threads
// normalise; the three accumulators take 384 TMEM columns (512 allocated).
#pragma once
#include "common.cuh"
#include "ops.cuh"

namespace infgen {
namespace ftc {
constexpr int TM = 128;                          // slots per tile
constexpr int NA = 3, NB = 3;                    // ring depths (32 KB stages)
constexpr int CHUNK = 8192;                      // floats per chunk: [hi 4096 | lo 4096]
constexpr int RT = 256;                          // row threads
constexpr int THREADS = RT + 64;
constexpr int SM_A = 0;
constexpr int SM_B = SM_A + NA * CHUNK;
constexpr int SM_EX = SM_B + NB * CHUNK;         // [2 buffers][2 halves][128] LayerNorm partials
constexpr int SM_RAW = SM_EX + 512;              // [128][4]
constexpr int SM_VALID = SM_RAW + 512;           // [128] int
constexpr int SM_BAR = SM_VALID + 128;           // full_a[NA] empty_a[NA] full_b[NB] empty_b[NB] g1[4] g2 g3 (uint64 each)
constexpr int N_BAR = 2 * NA + 2 * NB + 6;
constexpr int SM_TMEM = SM_BAR + 2 * N_BAR;
constexpr int SM_FLOATS = SM_TMEM + 4;
constexpr size_t SMEM = (size_t)SM_FLOATS * sizeof(float);
static_assert(SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);   // f32 acc, tf32 x tf32, N=128, M=128
constexpr uint32_t TMEM_COLS = 512;
// floats of the weight image of a D-dim embedding: (2 D + 1) GEMMs x 4 chunks, then the D "x rows" W0_d[128][:]
__host__ __device__ constexpr size_t wimg_floats(int D) { return (size_t)(2 * D + 1) * 4 * CHUNK + (size_t)D * 128; }
// GEMM order (shared by the image builder, the MMA warp and the row threads): G1_0, then per d: [G1_{d+1}], G2_d; G3.
// type 0 = G1, 1 = G2, 2 = G3
__host__ __device__ inline int job_list(int D, int *type, int *dim) {
    int n = 0;
    type[n] = 0; dim[n++] = 0;
    for (int d = 0; d < D; ++d) {
        if (d + 1 < D) { type[n] = 0; dim[n++] = d + 1; }
        type[n] = 1; dim[n++] = d;
    }
    type[n] = 2; dim[n++] = 0;
    return n;
}
}  // namespace ftc

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {      // no swizzle, K-major, LBO 2048 B, SBO 128 B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(ftc::IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// weight image builder: 4 chunks of one packed [32 k4][128][4] matrix -> [hi | lo] chunks
__global__ void k_wimg_split(const float *__restrict__ src, float *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * 4096) return;
    const float x = src[i], hi = tf32_rna(x), lo = tf32_rna(x - hi);
    const int c = i >> 12, o = i & 4095;
    dst[c * ftc::CHUNK + o] = hi;
    dst[c * ftc::CHUNK + 4096 + o] = lo;
}
// row 128 (the raw-input feature) of a packed [33 k4][128][4] first Linear
__global__ void k_wimg_xrow(const float *__restrict__ w0, float *__restrict__ dst) {
    dst[threadIdx.x] = w0[(32 * 128 + threadIdx.x) * 4];
}

__global__ void __launch_bounds__(ftc::THREADS, 1) k_fourier_tc(const FourierBatch fb) {
    using namespace ftc;
    extern __shared__ __align__(128) float smem_tc[];
    float *smem = smem_tc;
    float *sA = smem + SM_A, *sB = smem + SM_B, *sex = smem + SM_EX, *sraw = smem + SM_RAW;
    int *s_valid = reinterpret_cast<int *>(smem + SM_VALID);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *full_a = bars, *empty_a = full_a + NA, *full_b = empty_a + NA, *empty_b = full_b + NB, *g1_done = empty_b + NB,
             *g2_done = g1_done + 4, *g3_done = g2_done + 1;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + SM_TMEM);

    int j = 0;
    while (j + 1 < fb.n_jobs && (int)blockIdx.x >= fb.tile0[j + 1]) ++j;
    const FourierArgs &a = fb.job[j];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = ((int)blockIdx.x - fb.tile0[j]) * TM;
    const int D = a.dim;
    int v = 0;
    if (tid < TM) {
        const int s = s0 + tid;
        if (s < a.n_slots) v = a.cnt ? ((s % a.stride) < a.cnt[s / a.stride]) : 1;
        s_valid[tid] = v;
        for (int d = 0; d < 4; ++d) sraw[tid * 4 + d] = (v && d < D) ? a.raw[(size_t)s * D + d] : 0.f;
    }
    if (!__syncthreads_or(v)) return;

    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(&full_a[i], RT / 2); mbar_init(&empty_a[i], 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
        for (int i = 0; i < 6; ++i) mbar_init(&g1_done[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const int n_jobs = 2 * D + 1, n_chunks = 4 * n_jobs;

    if (warp == 9) {
        // ---- weight producer: lane s owns ring stage s ---------------------------------------------------------------
        if (lane < NB) {
            for (int i = lane; i < n_chunks; i += NB) {
                const int use = i / NB;
                if (use > 0) mbar_wait(&empty_b[lane], (uint32_t)(use - 1) & 1u);
                mbar_expect_tx(&full_b[lane], CHUNK * 4u);
                bulk_g2s(sB + lane * CHUNK, a.w.wimg + (size_t)i * CHUNK, CHUNK * 4u, &full_b[lane]);
            }
        }
    } else if (warp == 8) {
        // ---- MMA issuer -------------------------------------------------------------------------------------------------
        if (lane == 0) {
            int jt[9], jd[9];
            job_list(D, jt, jd);
            int ci = 0;
            for (int jb = 0; jb < n_jobs; ++jb) {
                const int type = jt[jb], d = jd[jb];
                const uint32_t acc = tmem + (type == 1 ? 256u : (uint32_t)(((type == 0 ? d : D) & 1) * 128));
                for (int c = 0; c < 4; ++c, ++ci) {
                    const int sa = ci % NA, sb = ci % NB;
                    mbar_wait(&full_a[sa], (uint32_t)(ci / NA) & 1u);
                    mbar_wait(&full_b[sb], (uint32_t)(ci / NB) & 1u);
                    tc_fence_after();
                    const uint32_t ab = smem_u32(sA + sa * CHUNK), bb = smem_u32(sB + sb * CHUNK);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t off = (uint32_t)ks * 4096u;                  // 8 k = two 2 KB core-matrix columns
                        const uint64_t ah = umma_desc(ab + off), al = umma_desc(ab + 16384u + off);
                        const uint64_t bh = umma_desc(bb + off), bl = umma_desc(bb + 16384u + off);
                        const uint32_t keep = (type == 1 ? (d > 0) : 0) | (c > 0) | (ks > 0);
                        umma_tf32(acc, al, bh, keep);
                        umma_tf32(acc, ah, bl, 1u);
                        umma_tf32(acc, ah, bh, 1u);
                    }
                    umma_commit(&empty_a[sa]);
                    umma_commit(&empty_b[sb]);
                }
                if (type == 0) umma_commit(&g1_done[d]);
                else if (type == 1 && d == D - 1) umma_commit(g2_done);
                else if (type == 2) umma_commit(g3_done);
            }
        }
    } else {
        // ---- row threads ---------------------------------------------------------------------------------------------
        const int h = warp >> 2, r = 32 * (warp & 3) + lane;
        const uint32_t trow = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        auto rt_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory"); };
        // write 32 consecutive k of row r into chunk ci of the A ring (hi / lo split) and publish it
        auto put_chunk = [&](int ci, const float *val) {
            const int s = ci % NA, use = ci / NA;
            if (use > 0) mbar_wait(&empty_a[s], (uint32_t)(use - 1) & 1u);
            float *p = sA + s * CHUNK + r * 4;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 hi, lo;
                hi.x = tf32_rna(val[4 * q + 0]); lo.x = tf32_rna(val[4 * q + 0] - hi.x);
                hi.y = tf32_rna(val[4 * q + 1]); lo.y = tf32_rna(val[4 * q + 1] - hi.y);
                hi.z = tf32_rna(val[4 * q + 2]); lo.z = tf32_rna(val[4 * q + 2] - hi.z);
                hi.w = tf32_rna(val[4 * q + 3]); lo.w = tf32_rna(val[4 * q + 3] - hi.w);
                st4(p + q * 512, hi);
                st4(p + 4096 + q * 512, lo);
            }
            fence_proxy_async();                  // generic-proxy stores -> visible to the tensor core (async proxy)
            tc_fence_before();
            mbar_arrive(&full_a[s]);
        };
        // sum over the 128 columns of the row from the two halves' partial sums
        int exb = 0;
        auto row_sum = [&](float part) {
            float *e = sex + exb * 256;
            exb ^= 1;
            e[h * 128 + r] = part;
            rt_sync();
            return e[r] + e[128 + r];
        };
        // features of dim d: cos -> chunk ci0 + h, sin -> chunk ci0 + 2 + h (freqs 32h .. 32h + 31)
        auto features = [&](int d, int ci0) {
            const float x = sraw[r * 4 + d];
            float cs[32], sn[32];
            const float *fq = a.w.freqs + d * 64 + 32 * h;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 f4 = ldg4(fq + 4 * q);
                const float f[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // x.unsqueeze(-1) * freqs * 2 * math.pi, evaluated left to right in fp32 (layers.py:151)
                    const float arg = __fmul_rn(__fmul_rn(__fmul_rn(x, f[i]), 2.0f), 3.14159265358979323846f);
                    sincosf(arg, &sn[4 * q + i], &cs[4 * q + i]);
                }
            }
            put_chunk(ci0 + h, cs);
            put_chunk(ci0 + 2 + h, sn);
        };
        // LayerNorm + ReLU of the row (this thread: columns 64h .. 64h+63 in val) -> chunks ci0 + 2h, ci0 + 2h + 1
        auto norm_relu_put = [&](float *val, const float *g, const float *b, int ci0) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) s += val[i];
            const float mean = row_sum(s) * (1.0f / HID);
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) { const float c = val[i] - mean; q = fmaf(c, c, q); }
            const float rstd = 1.0f / sqrtf(row_sum(q) * (1.0f / HID) + LN_EPS);
#pragma unroll
            for (int i4 = 0; i4 < 16; ++i4) {
                const float4 g4 = ldg4(g + 64 * h + 4 * i4), b4 = ldg4(b + 64 * h + 4 * i4);
                val[4 * i4 + 0] = fmaxf((val[4 * i4 + 0] - mean) * rstd * g4.x + b4.x, 0.f);
                val[4 * i4 + 1] = fmaxf((val[4 * i4 + 1] - mean) * rstd * g4.y + b4.y, 0.f);
                val[4 * i4 + 2] = fmaxf((val[4 * i4 + 2] - mean) * rstd * g4.z + b4.z, 0.f);
                val[4 * i4 + 3] = fmaxf((val[4 * i4 + 3] - mean) * rstd * g4.w + b4.w, 0.f);
            }
            put_chunk(ci0 + 2 * h, val);
            put_chunk(ci0 + 2 * h + 1, val + 32);
        };
        const float *xrow = a.w.wimg + (size_t)n_chunks * CHUNK;      // [D][128]
        float val[64];
        int ci = 0;
        features(0, ci);
        ci += 4;
        for (int d = 0; d < D; ++d) {
            if (d + 1 < D) { features(d + 1, ci); ci += 4; }
            mbar_wait(&g1_done[d], 0);
            tc_fence_after();
            const uint32_t t = trow + (uint32_t)((d & 1) * 128 + 64 * h);
            tmem_ld32(t, val);
            tmem_ld32(t + 32, val + 32);
            const float x = sraw[r * 4 + d];
#pragma unroll
            for (int i4 = 0; i4 < 16; ++i4) {
                const float4 b4 = ldg4(a.w.b0[d] + 64 * h + 4 * i4), w4 = ldg4(xrow + d * 128 + 64 * h + 4 * i4);
                val[4 * i4 + 0] += fmaf(x, w4.x, b4.x);
                val[4 * i4 + 1] += fmaf(x, w4.y, b4.y);
                val[4 * i4 + 2] += fmaf(x, w4.z, b4.z);
                val[4 * i4 + 3] += fmaf(x, w4.w, b4.w);
            }
            norm_relu_put(val, a.w.ln_g[d], a.w.ln_b[d], ci);
            ci += 4;
        }
        // sum over dims (accumulated in TMEM) + biases -> LN -> ReLU -> A of the output Linear
        mbar_wait(g2_done, 0);
        tc_fence_after();
        tmem_ld32(trow + 256u + (uint32_t)(64 * h), val);
        tmem_ld32(trow + 256u + (uint32_t)(64 * h) + 32, val + 32);
        for (int d = 0; d < D; ++d) {
#pragma unroll
            for (int i4 = 0; i4 < 16; ++i4) {
                const float4 b4 = ldg4(a.w.b3[d] + 64 * h + 4 * i4);
                val[4 * i4 + 0] += b4.x; val[4 * i4 + 1] += b4.y; val[4 * i4 + 2] += b4.z; val[4 * i4 + 3] += b4.w;
            }
        }
        norm_relu_put(val, a.w.out_ln_g, a.w.out_ln_b, ci);
        ci += 4;
        mbar_wait(g3_done, 0);
        tc_fence_after();
        {
            const uint32_t t = trow + (uint32_t)((D & 1) * 128 + 64 * h);
            tmem_ld32(t, val);
            tmem_ld32(t + 32, val + 32);
        }
#pragma unroll
        for (int i4 = 0; i4 < 16; ++i4) {
            const float4 b4 = ldg4(a.w.b_out + 64 * h + 4 * i4);
            val[4 * i4 + 0] += b4.x; val[4 * i4 + 1] += b4.y; val[4 * i4 + 2] += b4.z; val[4 * i4 + 3] += b4.w;
        }
        if (a.normalize) {                        // (y - mean) / sqrt(var + eps): input of every layer's attn_prenorm_r
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) s += val[i];
            const float mean = row_sum(s) * (1.0f / HID);
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 64; ++i) { const float c = val[i] - mean; q = fmaf(c, c, q); }
            const float rstd = 1.0f / sqrtf(row_sum(q) * (1.0f / HID) + LN_EPS);
#pragma unroll
            for (int i = 0; i < 64; ++i) val[i] = (val[i] - mean) * rstd;
        }
        if (s_valid[r]) {
            float *o = a.out + (size_t)(s0 + r) * 128 + 64 * h;
#pragma unroll
            for (int i4 = 0; i4 < 16; ++i4) st4(o + 4 * i4, make_float4(val[4 * i4], val[4 * i4 + 1], val[4 * i4 + 2], val[4 * i4 + 3]));
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

}  // namespace infgen
