// Probe: one tcgen05.mma kind::tf32 tile GEMM (M = 128, N = 128, K = 32 .. 128) with NO-swizzle K-major operands written
// by the threads themselves, accumulator in TMEM, read back with tcgen05.ld 32x32b.  Checks
//   (1) the shared-memory descriptor semantics (LBO = K-direction core-matrix stride, SBO = 8-row-group stride),
//   (2) the TMEM lane/column mapping of the accumulator,
//   (3) the accuracy of the 3xTF32 split (hi*hi + lo*hi + hi*lo) against an fp64 reference,
//   (4) the issue rate of back-to-back MMAs (cycles per 128x128x8 instruction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    return d;                                  // layout type 0 = no swizzle
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
            smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
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

// tile image: element (r, k) of a [128 x K] K-major operand at float offset (k/4)*512 + r*4 + (k%4)
//   -> core matrix (8 rows x 16 B) contiguous 128 B; 8-row groups 128 B apart (SBO); k-chunks of 4 are 2048 B apart (LBO)
constexpr int KMAX = 64;
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);

// mode 0: single tf32 pass;  mode 1: 3xTF32;  swap: exchange LBO / SBO in the descriptor (expected to be wrong)
__global__ void __launch_bounds__(128, 1) k_probe(const float *A, const float *B, float *D, int K, int mode, int swap, int reps,
                                                 long long *cycles) {
    extern __shared__ __align__(128) float smem[];
    float *a_hi = smem, *a_lo = a_hi + 128 * KMAX, *b_hi = a_lo + 128 * KMAX, *b_lo = b_hi + 128 * KMAX;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // operands: thread = row
    for (int k4 = 0; k4 < K / 4; ++k4) {
        float4 a = *reinterpret_cast<const float4 *>(A + (size_t)tid * K + 4 * k4);
        float4 b = *reinterpret_cast<const float4 *>(B + (size_t)tid * K + 4 * k4);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
        float ah[4], al[4], bh[4], bl[4];
        for (int i = 0; i < 4; ++i) {
            ah[i] = __uint_as_float(f2tf32(av[i]));
            al[i] = __uint_as_float(f2tf32(av[i] - ah[i]));
            bh[i] = __uint_as_float(f2tf32(bv[i]));
            bl[i] = __uint_as_float(f2tf32(bv[i] - bh[i]));
        }
        const int o = k4 * 512 + tid * 4;
        *reinterpret_cast<float4 *>(a_hi + o) = make_float4(ah[0], ah[1], ah[2], ah[3]);
        *reinterpret_cast<float4 *>(a_lo + o) = make_float4(al[0], al[1], al[2], al[3]);
        *reinterpret_cast<float4 *>(b_hi + o) = make_float4(bh[0], bh[1], bh[2], bh[3]);
        *reinterpret_cast<float4 *>(b_lo + o) = make_float4(bl[0], bl[1], bl[2], bl[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lbo = swap ? 128u : 2048u, sbo = swap ? 2048u : 128u;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint32_t off = (uint32_t)ks * 4096u;          // two k-chunks of 4 per MMA
                const uint64_t dah = make_desc(smem_u32(a_hi) + off, lbo, sbo), dal = make_desc(smem_u32(a_lo) + off, lbo, sbo);
                const uint64_t dbh = make_desc(smem_u32(b_hi) + off, lbo, sbo), dbl = make_desc(smem_u32(b_lo) + off, lbo, sbo);
                if (mode == 1) {
                    mma_tf32(tmem, dal, dbh, IDESC, (rep | ks) ? 1u : 0u);
                    mma_tf32(tmem, dah, dbl, IDESC, 1u);
                    mma_tf32(tmem, dah, dbh, IDESC, 1u);
                } else {
                    mma_tf32(tmem, dah, dbh, IDESC, (rep | ks) ? 1u : 0u);
                }
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 32; ++i) D[(size_t)tid * 128 + c0 + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

int main() {
    const int K = 64;
    std::vector<float> A(128 * K), B(128 * K), D(128 * 128);
    srand(1);
    for (auto &v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.2f;
    float *dA, *dB, *dD;
    long long *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 4 * 128 * KMAX * 4;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<double> ref(128 * 128);
    std::vector<float> ref32(128 * 128), reft(128 * 128);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            double s = 0; float s32 = 0.f; double st = 0;
            for (int k = 0; k < K; ++k) {
                s += (double)A[m * K + k] * B[n * K + k];
                s32 = fmaf(A[m * K + k], B[n * K + k], s32);
                st += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(B[n * K + k]);
            }
            ref[m * 128 + n] = s; ref32[m * 128 + n] = s32; reft[m * 128 + n] = (float)st;
        }
    for (int mode = 0; mode < 2; ++mode)
        for (int swap = 0; swap < 1; ++swap) {
            cudaMemset(dD, 0, D.size() * 4);
            k_probe<<<1, 128, smem>>>(dA, dB, dD, K, mode, swap, 1, dC);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d swap %d: CUDA error %s\n", mode, swap, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double e64 = 0, e32 = 0, et = 0, mag = 0;
            for (int i = 0; i < 128 * 128; ++i) {
                e64 = fmax(e64, fabs(D[i] - ref[i]));
                e32 = fmax(e32, fabs((double)ref32[i] - ref[i]));
                et = fmax(et, fabs((double)D[i] - reft[i]));
                mag = fmax(mag, fabs(ref[i]));
            }
            printf("mode %d (%s) swap %d: max|D-ref64| = %.3e   max|D-tf32ref| = %.3e   (fp32 FMA chain err %.3e, max|ref| %.3f)\n", mode,
                   mode ? "3xTF32" : "1xTF32", swap, e64, et, e32, mag);
        }
    // issue rate
    for (int mode = 0; mode < 2; ++mode) {
        const int reps = 64;
        k_probe<<<1, 128, smem>>>(dA, dB, dD, K, mode, 0, reps, dC);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
        const double n_mma = (double)reps * (K / 8) * (mode ? 3 : 1);
        printf("mode %d: %lld cycles for %.0f MMAs (128x128x8) = %.1f cycles each -> %.0f MAC/clk/SM\n", mode, c, n_mma, c / n_mma,
               128.0 * 128 * 8 * n_mma / c);
    }
    return 0;
}
