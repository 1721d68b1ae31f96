// FourierEmbedding (reference layers.py:142-160) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// 3xTF32 error-compensated split, accumulators in TMEM.
//
// One CTA embeds a tile of 128 edge slots.  Every Linear of the module is a [128 slots x 128] x [128 x 128] GEMM:
//     per input dim d:  G1_d = [cos(2 pi f x_d) | sin(2 pi f x_d)] W0_d[0:128]     (+ x_d W0_d[128] + b0_d in the epilogue;
//                              K permuted so that a 32-k chunk holds cos and sin of the same 16 frequencies)
//                       G2_d = relu(LN(G1_d)) W3_d, summed over d IN TMEM (the accumulator is simply not cleared)
//     output:           G3   = relu(LN(sum_d G2_d + b3_d)) Wout + bout             (optionally standardised -> rhat)
//
// fp32 fidelity: the reference is fp32 and the greedy argmax of a closed loop has to survive, so every operand is split
// x = hi + lo (both TF32, cvt.rna) and D += A_lo B_hi + A_hi B_lo + A_hi B_hi (fp32 accumulate); measured error of this
// scheme on B200: 1.4e-6 against fp64 for K = 64, |D| <= 2.3 (tools/probe/umma_probe.cu), the same as an fp32 FMA chain.
//
// Operands.  A (features / normalised activations) lives in TENSOR MEMORY: tcgen05.mma with the A matrix in TMEM, lane =
// row, one 32-bit column per k; four 64-column stages [hi 32 | lo 32], written by the row threads with tcgen05.st - no
// shared-memory staging, no proxy fence.  B (weights) comes from shared memory in the NO-swizzle K-major core-matrix
// layout measured in the probe: element (n, k) of a [128 x 32] chunk at float offset (k / 4) * 512 + n * 4 + (k % 4), i.e.
// LBO (K direction) = 2048 B, SBO (8-row groups) = 128 B.  That is exactly the packed [K/4][128][4] layout of the weight
// blob, so the B images are the blob matrices split into hi / lo and cut into 32-k chunks of [hi 16 KB | lo 16 KB], stored
// in consumption order: one 32 KB cp.async.bulk per chunk into a 6-stage ring.
// TMEM columns: [0,128) G1 / G3 accumulator, [128,256) dim-sum accumulator, [256,512) the four A stages.
//
// Roles (576 threads):
//   warp 0     MMA issuer (12 tcgen05.mma per chunk + commits to the ring / accumulator mbarriers).  It is warp 0 - the
//              oldest warp of its scheduler partition - and runs in uniform control flow with elect.sync, because as the
//              youngest warp, issuing from one divergent lane, it was starved by the row warps (110-145 cycles per MMA
//              against 68 at peak; 88-99 now).
//   warps 1-16 row threads: thread (r = 32 (w & 3) + lane, quarter qd = (w - 1) >> 2) owns slot r and 32 of its 128
//              columns - one 32-k chunk of every A operand (TMEM lanes are only visible to warps with the same w & 3).
//              They generate the Fourier features and run the epilogues (TMEM -> registers -> bias / LayerNorm / ReLU
//              -> hi/lo split -> TMEM).  This part, not the tensor core, bounds the kernel (sincosf, LayerNorm).
//   warp 17    streams the weight chunks through the B ring.
// G1_{d+1} is issued as soon as the row threads have loaded the result of G1_d into registers (acc_free), so the tensor
// core works while they normalise.  Every mbarrier wait is bounded (ftc_wait): a protocol bug becomes an error code, not
// a hung GPU.
#pragma once
#include "common.cuh"
#include "ops.cuh"

namespace infgen {
namespace ftc {
constexpr int TM = 128;                          // slots per tile
// A operands live in TMEM (tcgen05.mma with the A matrix in tensor memory): 4 stages of 64 columns [hi 32 | lo 32], one
// per quarter of the row threads.  NA must be 4: quarter q then always writes A stage q, so every parity wait is exactly
// one phase ahead of the last phase that thread observed (with a 3-deep ring a quarter that skipped a stage's previous
// use raced two phases ahead and mbarrier.try_wait.parity aliased - found by the watchdog below).
// B (weights): NB stages of 32 KB in shared memory.
constexpr int NA = 4, NB = 6;
static_assert(NA == 4, "A-ring protocol: one stage per quarter");
constexpr uint32_t TC_ACC_H = 0, TC_ACC_S = 128, TC_A = 256;    // TMEM columns: G1/G3 accumulator, dim-sum accumulator, A stages
constexpr int CHUNK = 8192;                      // floats per chunk: [hi 4096 | lo 4096]
constexpr int RT = 512;                          // row threads
constexpr int THREADS = RT + 64;
constexpr int SM_B = 0;
constexpr int SM_EX = SM_B + NB * CHUNK;         // [2 buffers][4 quarters][128] LayerNorm partials
constexpr int SM_RAW = SM_EX + 1024;              // [128][4]
constexpr int SM_VALID = SM_RAW + 512;           // [128] int
constexpr int SM_SLOT = SM_VALID + 128;          // [128] int: slot of every row of the tile
constexpr int SM_BAR = SM_SLOT + 128;           // full_a[NA] empty_a[NA] full_b[NB] empty_b[NB] g1[4] g2 g3 acc_free (uint64 each)
constexpr int N_BAR = 2 * NA + 2 * NB + 7;
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
// A from tensor memory (lane = row, one 32-bit column per k), B from shared memory.  Called by the WHOLE warp with
// warp-uniform operands; one elected lane issues (keeps the issue loop in uniform control flow: no per-lane waterfall).
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(ftc::IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *b) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(b))
        : "memory");
}
// 32 consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
        "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]),
        "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]),
        "f"(v[31])
        : "memory");
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
// x = hi + lo with hi = x rounded to TF32 (nearest, ties away: integer add + mask, full-rate ALU instead of cvt.rna);
// lo = x - hi is exact in fp32 and the tensor core ignores its low 13 mantissa bits (2^-21 |x| at most)
__device__ __forceinline__ void split_tf32x2(float x, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Bounded mbarrier wait: a protocol bug must not hang the GPU.  After ~10 ms (2e7 cycles) the wait gives up, records
// the first (code, thread) of each wait class in g_ftc_hang and execution continues with whatever is there, so the kernel always terminates;
// the engine reports a non-zero g_ftc_hang[0] as an error (infgen_op_fourier_embedding, tests).
__device__ int g_ftc_hang[8];
template <bool BACKOFF = false>
__device__ __forceinline__ void ftc_wait(uint64_t *b, uint32_t parity, int code) {
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
        if (BACKOFF) __nanosleep(64);             // pollers must not take issue slots from the MMA warp
        if (clock64() - t0 > 20000000ll) {
            atomicAdd(&g_ftc_hang[0], 1);
            atomicCAS(&g_ftc_hang[code / 100], 0, code * 1000 + (int)threadIdx.x);   // first of each class: code, thread
            return;
        }
    }
}

#ifdef INFGEN_FTC_TRACE
// debug: clock64 stamps of [traced block 0/1][stream: 0 = row thread 0, 1 = MMA lane, 2 = row thread 128][64]
__device__ long long g_ftc_trace[2][3][64];
#define FTC_STAMP(stream)                                                                                     \
    do {                                                                                                      \
        if (trace_blk >= 0 && trace_n < 64) g_ftc_trace[trace_blk][stream][trace_n++] = clock64();           \
    } while (0)
#else
#define FTC_STAMP(stream) do {} while (0)
#endif

// Compact list of the valid slots of a strided slot space [n_rows][stride] (slot valid iff k < cnt[row]): one CTA, block
// scan over the rows.  list[off(row) + k] = row * stride + k, *n_list = total.  Batches only: a2a rows hold ~40 valid of 64
// slots, so the strided walk spends a third of the 128-slot tiles on padding.
__global__ void __launch_bounds__(1024) k_slot_compact(const int *__restrict__ cnt, int n_rows, int stride, int *__restrict__ list,
                                                        int *__restrict__ n_list, int *__restrict__ row_off) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < n_rows; r0 += 1024) {             // exclusive scan of the row counts -> row_off
        const int r = r0 + tid;
        const int c = r < n_rows ? min(cnt[r], stride) : 0;
        int x = c;                                           // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;                                // inclusive prefix of the warp totals
        }
        __syncthreads();
        if (r < n_rows) row_off[r] = s_base + (warp ? s_warp[warp - 1] : 0) + x - c;
        __syncthreads();
        if (tid == 0) s_base += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) *n_list = s_base;
    // one warp per row writes its slots (a serial loop per thread cost 32 us at 2,048 rows)
    for (int r = warp; r < n_rows; r += 32) {
        const int c = min(cnt[r], stride), off = row_off[r];
        for (int k = lane; k < c; k += 32) list[off + k] = r * stride + k;
    }
}

// Per-dim MLP output of input dim d (layers.py:152-155: Linear(129,128) -> LN -> ReLU -> Linear(128,128)) for the inputs
// x = -1, -2, .. (one block per value), plain fp32 FMA from the packed [K/4][128][4] weights.  Run once at infgen_create.
__global__ void __launch_bounds__(128) k_fourier_dim_table(const FourierW w, int d, float *__restrict__ out) {
    __shared__ float feat[132], hid[128], red[8];
    const int n = threadIdx.x, lane = n & 31, warp = n >> 5;
    const float x = -(float)(blockIdx.x + 1);
    if (n < 64) {
        const float arg = __fmul_rn(__fmul_rn(__fmul_rn(x, __ldg(w.freqs + d * 64 + n)), 2.0f), 3.14159265358979323846f);
        float sn, cs;
        sincosf(arg, &sn, &cs);
        feat[n] = cs;
        feat[64 + n] = sn;
    }
    if (n == 0) { feat[128] = x; feat[129] = 0.f; feat[130] = 0.f; feat[131] = 0.f; }
    __syncthreads();
    float y = __ldg(w.b0[d] + n);
    for (int k = 0; k < 129; ++k) y = fmaf(feat[k], __ldg(w.w0[d] + ((size_t)(k >> 2) * 128 + n) * 4 + (k & 3)), y);
    // LayerNorm over the 128 outputs (two passes, as torch)
    float s = warp_sum(y);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    const float mean = (red[0] + red[1] + red[2] + red[3]) * (1.0f / HID);
    const float c = y - mean;
    float q = warp_sum(c * c);
    if (lane == 0) red[4 + warp] = q;
    __syncthreads();
    const float rstd = 1.0f / sqrtf((red[4] + red[5] + red[6] + red[7]) * (1.0f / HID) + LN_EPS);
    hid[n] = fmaxf(c * rstd * __ldg(w.ln_g[d] + n) + __ldg(w.ln_b[d] + n), 0.f);
    __syncthreads();
    float z = __ldg(w.b3[d] + n);
    for (int k = 0; k < 128; ++k) z = fmaf(hid[k], __ldg(w.w3[d] + ((size_t)(k >> 2) * 128 + n) * 4 + (k & 3)), z);
    out[(size_t)blockIdx.x * 128 + n] = z;
}

// weight image builder: 4 chunks of one packed [32 k4][128][4] matrix -> [hi | lo] chunks.  fourier_order: the K order of
// G1 (chunk c = k4 rows 4c .. 4c+3 (cos of 16 freqs) then 16+4c .. 16+4c+3 (their sin)); else K in natural order
__global__ void k_wimg_split(const float *__restrict__ src, float *__restrict__ dst, int fourier_order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * 4096) return;
    const int c = i >> 12, o = i & 4095;
    int si = i;
    if (fourier_order) {
        const int q = o >> 9;                     // k4 group inside the chunk
        si = ((q < 4 ? 4 * c + q : 16 + 4 * c + (q - 4)) << 9) + (o & 511);
    }
    const float x = src[si], hi = tf32_rna(x), lo = tf32_rna(x - hi);
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
    float *sB = smem + SM_B, *sex = smem + SM_EX, *sraw = smem + SM_RAW;
    int *s_valid = reinterpret_cast<int *>(smem + SM_VALID);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *full_a = bars, *empty_a = full_a + NA, *full_b = empty_a + NA, *empty_b = full_b + NB, *g1_done = empty_b + NB,
             *g2_done = g1_done + 4, *g3_done = g2_done + 1, *acc_free = g3_done + 1;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + SM_TMEM);

    int j = 0;
    while (j + 1 < fb.n_jobs && (int)blockIdx.x >= fb.tile0[j + 1]) ++j;
    const FourierArgs &a = fb.job[j];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = ((int)blockIdx.x - fb.tile0[j]) * TM;
    const int D = a.dim;
#ifdef INFGEN_FTC_TRACE
    const int trace_blk = blockIdx.x == 0 ? 0 : ((fb.n_jobs > 1 && (int)blockIdx.x == fb.tile0[1]) ? 1 : -1);
    int trace_n = 0;
#endif
    int v = 0;
    int *s_slot = reinterpret_cast<int *>(smem + SM_SLOT);
    if (tid < TM) {
        int s = s0 + tid;
        if (a.slot_list) {                        // compact list of valid slots: no tile is spent on padding
            v = s < *a.n_list;
            s = v ? a.slot_list[s] : 0;
        } else if (s < a.n_slots) {
            v = a.cnt ? ((s % a.stride) < a.cnt[s / a.stride]) : 1;
        }
        s_valid[tid] = v;
        s_slot[tid] = s;
        const int rs = a.raw_stride ? a.raw_stride : D;
        for (int d = 0; d < 4; ++d) sraw[tid * 4 + d] = (v && d < rs) ? a.raw[(size_t)s * rs + d] : 0.f;
    }
    if (!__syncthreads_or(v)) return;

    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(&full_a[i], 4); mbar_init(&empty_a[i], 1); }   // one arrival per warp of a quarter
        for (int i = 0; i < NB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
        for (int i = 0; i < 6; ++i) mbar_init(&g1_done[i], 1);
        mbar_init(acc_free, RT / 32);
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

    if (warp == RT / 32 + 1) {
        // ---- weight producer (warp 17) ------------------------------------------------------------------------------------------
        // (the whole warp walks the loop and one lane issues, so that the warp reaches the final barrier converged)
        for (int i = 0; i < n_chunks; ++i) {
            const int st = i % NB, use = i / NB;
            if (lane == 0) {
                if (use > 0) ftc_wait(&empty_b[st], (uint32_t)(use - 1) & 1u, 100 + i);
                mbar_expect_tx(&full_b[st], CHUNK * 4u);
                bulk_g2s(sB + st * CHUNK, a.w.wimg + (size_t)i * CHUNK, CHUNK * 4u, &full_b[st]);
            }
            __syncwarp();
        }
    } else if (warp == 0) {
        // ---- MMA issuer: warp 0, the oldest warp of its scheduler partition, so that it wins the issue slot whenever it
        // is ready (as the youngest warp it was starved by the row warps: 110-145 cycles per MMA instead of 68).  The whole
        // warp runs the loop in uniform control flow; one elected lane issues each tcgen05 instruction.
        int jt[9], jd[9];
        job_list(D, jt, jd);
        int ci = 0, n_h = 0;                      // n_h: uses of the G1/G3 accumulator so far
        if (lane == 0) FTC_STAMP(1);
        for (int jb = 0; jb < n_jobs; ++jb) {
            const int type = jt[jb], d = jd[jb];
            const uint32_t acc = tmem + (type == 1 ? TC_ACC_S : TC_ACC_H);
            if (type != 1) {
                // the single G1/G3 accumulator is free again once every row warp has loaded the previous G1 result
                if (jb > 0) ftc_wait(acc_free, (uint32_t)(n_h - 1) & 1u, 600 + n_h);
                ++n_h;
            }
            for (int c = 0; c < 4; ++c, ++ci) {
                const int sb = ci % NB;
                ftc_wait(&full_a[c], (uint32_t)(ci / NA) & 1u, 200 + ci);
                ftc_wait(&full_b[sb], (uint32_t)(ci / NB) & 1u, 300 + ci);
                tc_fence_after();
                const uint32_t at = tmem + TC_A + 64u * (uint32_t)c;
                const uint64_t b0 = umma_desc(smem_u32(sB + sb * CHUNK));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ah = at + 8u * ks, al = at + 32u + 8u * ks;      // 8 k = 8 TMEM columns
                    const uint64_t bh = b0 + (uint64_t)(ks * (4096 >> 4)), bl = bh + (16384 >> 4);   // 8 k = 4 KB of B
                    const uint32_t keep = (type == 1 ? (d > 0) : 0) | (c > 0) | (ks > 0);
                    umma_tf32_ts(acc, al, bh, keep);
                    umma_tf32_ts(acc, ah, bl, 1u);
                    umma_tf32_ts(acc, ah, bh, 1u);
                }
                umma_commit_elect(&empty_a[c]);
                umma_commit_elect(&empty_b[sb]);
                if (c == 3) {
                    if (type == 0) umma_commit_elect(&g1_done[d]);
                    else if (type == 1 && d == D - 1) umma_commit_elect(g2_done);
                    else if (type == 2) umma_commit_elect(g3_done);
                }
                if (lane == 0) FTC_STAMP(1);
            }
        }
    } else {
        // ---- row threads ---------------------------------------------------------------------------------------------
        const int qd = (warp - 1) >> 2, r = 32 * (warp & 3) + lane;    // warps 1..16; TMEM lane group = warp % 4
        const uint32_t trow = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(32 * qd);
#ifdef INFGEN_FTC_TRACE
        const int rt_stream = tid == 32 ? 0 : 2;
#define RT_STAMP() do { if (tid == 32 || tid == 160) FTC_STAMP(rt_stream); } while (0)
#else
#define RT_STAMP() do {} while (0)
#endif
        auto rt_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory"); };
        const uint32_t ta = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + TC_A + 64u * (uint32_t)qd;   // this thread's A stage
        // A stage qd is free again once the MMAs of its previous chunk have completed
        // one lane per warp polls (with back-off): 512 spinning threads took the issue slots the MMA warp needs
        auto warp_wait = [&](uint64_t *b, uint32_t parity, int code) {
            if (lane == 0) ftc_wait<true>(b, parity, code);
            __syncwarp();
            tc_fence_after();
        };
        auto a_acquire = [&](int ci) {
            const int use = ci / NA;
            if (use > 0) warp_wait(&empty_a[qd], (uint32_t)(use - 1) & 1u, 400 + ci);
        };
        auto a_publish = [&]() {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[qd]);
        };
        // this thread's 32 k of the row -> [hi | lo] columns of A stage qd
        auto put_chunk = [&](int ci, const float *val) {
            a_acquire(ci);
            float t[32];                          // hi, then lo (one at a time: 96 live values would spill)
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __uint_as_float((__float_as_uint(val[i]) + 0x1000u) & 0xFFFFE000u);
            tmem_st32(ta, t);
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = val[i] - t[i];
            tmem_st32(ta + 32u, t);
            a_publish();
        };
        // sum over the 128 columns of the row from the four quarters' partial sums
        int exb = 0;
        auto row_sum = [&](float part) {
            float *e = sex + exb * 512;
            exb ^= 1;
            e[qd * 128 + r] = part;
            rt_sync();
            return (e[r] + e[128 + r]) + (e[256 + r] + e[384 + r]);
        };
        // features of dim d.  K order of G1: chunk c = [cos(f_16c .. f_16c+15) | sin(f_16c .. f_16c+15)], so a chunk is
        // 16 sincosf of one thread (quarter qd fills chunk ci0 + qd).  The sincosf loop stays rolled: with all of them
        // inlined the kernel was 168 KB of SASS and the row warps thrashed the instruction cache.
        auto features = [&](int d, int ci0, float *val) {
            const float x = sraw[r * 4 + d];
            const float *fq = a.w.freqs + d * 64 + 16 * qd;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const float4 f4 = ldg4(fq + 4 * q);
                const float f[4] = {f4.x, f4.y, f4.z, f4.w};
                float cs[4], sn[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // x.unsqueeze(-1) * freqs * 2 * math.pi, evaluated left to right in fp32 (layers.py:151)
                    const float arg = __fmul_rn(__fmul_rn(__fmul_rn(x, f[i]), 2.0f), 3.14159265358979323846f);
                    sincosf(arg, &sn[i], &cs[i]);
                }
                // (the switch keeps val[] in registers although q is a loop variable)
                switch (q) {
                case 0: val[0] = cs[0]; val[1] = cs[1]; val[2] = cs[2]; val[3] = cs[3]; val[16] = sn[0]; val[17] = sn[1]; val[18] = sn[2]; val[19] = sn[3]; break;
                case 1: val[4] = cs[0]; val[5] = cs[1]; val[6] = cs[2]; val[7] = cs[3]; val[20] = sn[0]; val[21] = sn[1]; val[22] = sn[2]; val[23] = sn[3]; break;
                case 2: val[8] = cs[0]; val[9] = cs[1]; val[10] = cs[2]; val[11] = cs[3]; val[24] = sn[0]; val[25] = sn[1]; val[26] = sn[2]; val[27] = sn[3]; break;
                default: val[12] = cs[0]; val[13] = cs[1]; val[14] = cs[2]; val[15] = cs[3]; val[28] = sn[0]; val[29] = sn[1]; val[30] = sn[2]; val[31] = sn[3]; break;
                }
            }
            put_chunk(ci0 + qd, val);
        };
        // two-pass mean / rstd of the row (this thread holds 32 of its 128 values)
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
        // LayerNorm + ReLU of the row -> chunk ci0 + qd
        auto norm_relu_put = [&](float *val, const float *g, const float *b, int ci0) {
            float mean, rstd;
            row_stats(val, mean, rstd);
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 g4 = ldg4(g + 32 * qd + 4 * i4), b4 = ldg4(b + 32 * qd + 4 * i4);
                val[4 * i4 + 0] = fmaxf((val[4 * i4 + 0] - mean) * rstd * g4.x + b4.x, 0.f);
                val[4 * i4 + 1] = fmaxf((val[4 * i4 + 1] - mean) * rstd * g4.y + b4.y, 0.f);
                val[4 * i4 + 2] = fmaxf((val[4 * i4 + 2] - mean) * rstd * g4.z + b4.z, 0.f);
                val[4 * i4 + 3] = fmaxf((val[4 * i4 + 3] - mean) * rstd * g4.w + b4.w, 0.f);
            }
            put_chunk(ci0 + qd, val);
        };
        const float *xrow = a.w.wimg + (size_t)n_chunks * CHUNK;      // [D][128]
        float val[32];
        int ci = 0;
        RT_STAMP();
        features(0, ci, val);
        ci += 4;
        RT_STAMP();
        for (int d = 0; d < D; ++d) {
            if (d + 1 < D) { features(d + 1, ci, val); ci += 4; }
            RT_STAMP();
            warp_wait(&g1_done[d], 0, 500 + d);
            RT_STAMP();
            tmem_ld32(trow + TC_ACC_H, val);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);  // G1_{d+1} / G3 may overwrite the accumulator
            const float x = sraw[r * 4 + d];
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 b4 = ldg4(a.w.b0[d] + 32 * qd + 4 * i4), w4 = ldg4(xrow + d * 128 + 32 * qd + 4 * i4);
                val[4 * i4 + 0] += fmaf(x, w4.x, b4.x);
                val[4 * i4 + 1] += fmaf(x, w4.y, b4.y);
                val[4 * i4 + 2] += fmaf(x, w4.z, b4.z);
                val[4 * i4 + 3] += fmaf(x, w4.w, b4.w);
            }
            norm_relu_put(val, a.w.ln_g[d], a.w.ln_b[d], ci);
            ci += 4;
            RT_STAMP();
        }
        // sum over dims (accumulated in TMEM) + biases -> LN -> ReLU -> A of the output Linear
        warp_wait(g2_done, 0, 504);
        RT_STAMP();
        tmem_ld32(trow + TC_ACC_S, val);
        for (int d = 0; d < D; ++d) {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 b4 = ldg4(a.w.b3[d] + 32 * qd + 4 * i4);
                val[4 * i4 + 0] += b4.x; val[4 * i4 + 1] += b4.y; val[4 * i4 + 2] += b4.z; val[4 * i4 + 3] += b4.w;
            }
        }
        if (a.dim_table) {
            // input dim D takes only the values -1 .. -table_n (the column offset of a temporal edge, agent_decoder.py:607):
            // its per-dim MLP output comes from a table built at infgen_create instead of two more GEMMs per tile
            const int k = min(max((int)lrintf(-sraw[r * 4 + D]) - 1, 0), a.table_n - 1);
            const float *tb = a.dim_table + (size_t)k * 128 + 32 * qd;
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 t4 = ldg4(tb + 4 * i4);
                val[4 * i4 + 0] += t4.x; val[4 * i4 + 1] += t4.y; val[4 * i4 + 2] += t4.z; val[4 * i4 + 3] += t4.w;
            }
        }
        norm_relu_put(val, a.w.out_ln_g, a.w.out_ln_b, ci);
        ci += 4;
        RT_STAMP();
        warp_wait(g3_done, 0, 505);
        RT_STAMP();
        tmem_ld32(trow + TC_ACC_H, val);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
            const float4 b4 = ldg4(a.w.b_out + 32 * qd + 4 * i4);
            val[4 * i4 + 0] += b4.x; val[4 * i4 + 1] += b4.y; val[4 * i4 + 2] += b4.z; val[4 * i4 + 3] += b4.w;
        }
        if (a.normalize) {                        // (y - mean) / sqrt(var + eps): input of every layer's attn_prenorm_r
            float mean, rstd;
            row_stats(val, mean, rstd);
#pragma unroll
            for (int i = 0; i < 32; ++i) val[i] = (val[i] - mean) * rstd;
        }
        if (s_valid[r]) {
            float *o = a.out + (size_t)s_slot[r] * 128 + 32 * qd;
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) st4(o + 4 * i4, make_float4(val[4 * i4], val[4 * i4 + 1], val[4 * i4 + 2], val[4 * i4 + 3]));
        }
        RT_STAMP();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

}  // namespace infgen
