// Probe: how fast can one warp per destination row gather 1 KB K|V rows + 512 B rhat rows (the access pattern of k_attn)?
//   mode 0  plain LDG.128, E_IN_FLIGHT edges in registers (what k_attn did in round 1: 2 in flight)
//   mode 1  cp.async.bulk (TMA bulk copy) issued by lane 0 of the warp into a per-warp shared-memory ring, mbarrier completion
//   mode 2  cp.async.bulk issued by lane (edge % 32): up to 32 issuing lanes per warp
//   mode 3  cp.async.cg 16 B per lane (LDGSTS) into the ring, cp.async.wait_group
// Reports GB/s of gathered bytes.  Working sets: kv rows 3 MB (L2 resident) / 384 MB; rhat streams 64 MB / 512 MB.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu && ./gather_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

constexpr int WARPS = 8;
template <int MODE, int NS>
__global__ void __launch_bounds__(WARPS * 32) k_gather(const float *__restrict__ kv, const float *__restrict__ rhat, const int *__restrict__ src,
                                                         const int *__restrict__ cnt, int stride, int n_rows, float *out) {
    extern __shared__ __align__(128) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + warp;
    float *ring = smem + (size_t)warp * NS * 384;                      // [NS][rhat 128 | k 128 | v 128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * NS * 384) + warp * NS;
    if (MODE == 1 || MODE == 2) {
        if (lane < NS) mbar_init(&bars[lane], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    if (r >= n_rows) return;
    const int n = cnt[r];
    const int e0 = r * stride;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 0) {
        constexpr int F = NS;                                          // edges in flight
        float4 a[F], b[F], c[F];
        for (int i0 = 0; i0 < n; i0 += F) {
#pragma unroll
            for (int j = 0; j < F; ++j)
                if (i0 + j < n) {
                    const int s = src[e0 + i0 + j];
                    a[j] = *reinterpret_cast<const float4 *>(rhat + (size_t)(e0 + i0 + j) * 128 + 4 * lane);
                    b[j] = __ldcg(reinterpret_cast<const float4 *>(kv + (size_t)s * 256 + 4 * lane));
                    c[j] = __ldcg(reinterpret_cast<const float4 *>(kv + (size_t)s * 256 + 128 + 4 * lane));
                }
#pragma unroll
            for (int j = 0; j < F; ++j)
                if (i0 + j < n) { acc.x += a[j].x * b[j].x + c[j].x; acc.y += a[j].y * b[j].y + c[j].y; acc.z += a[j].z + b[j].z * c[j].z; acc.w += a[j].w + b[j].w + c[j].w; }
        }
    } else if (MODE == 1 || MODE == 2) {
        auto issue = [&](int i) {                                       // edge i -> stage i % NS
            const int st = i % NS;
            const int s = src[e0 + i];
            float *d = ring + st * 384;
            mbar_expect(&bars[st], 1536);
            bulk(d, rhat + (size_t)(e0 + i) * 128, 512, &bars[st]);
            bulk(d + 128, kv + (size_t)s * 256, 1024, &bars[st]);
        };
        const int pre = n < NS ? n : NS;
        if (MODE == 1) { if (lane == 0) for (int i = 0; i < pre; ++i) issue(i); }
        else if (lane < pre) issue(lane);
        for (int i = 0; i < n; ++i) {
            const int st = i % NS;
            mbar_wait(&bars[st], (uint32_t)(i / NS) & 1u);
            const float *d = ring + st * 384;
            const float4 a = *reinterpret_cast<const float4 *>(d + 4 * lane), b = *reinterpret_cast<const float4 *>(d + 128 + 4 * lane),
                         c = *reinterpret_cast<const float4 *>(d + 256 + 4 * lane);
            acc.x += a.x * b.x + c.x; acc.y += a.y * b.y + c.y; acc.z += a.z + b.z * c.z; acc.w += a.w + b.w + c.w;
            __syncwarp();
            if (i + NS < n) {
                if (MODE == 1) { if (lane == 0) issue(i + NS); }
                else if (lane == (i + NS) % 32) issue(i + NS);
            }
        }
    } else {
        auto issue = [&](int i) {
            const int st = i % NS;
            const int s = src[e0 + i];
            float *d = ring + st * 384 + 4 * lane;
            const float *g0 = rhat + (size_t)(e0 + i) * 128 + 4 * lane, *g1 = kv + (size_t)s * 256 + 4 * lane;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(g0) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + 128)), "l"(g1) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + 256)), "l"(g1 + 128) : "memory");
        };
        for (int i = 0; i < NS; ++i) {
            if (i < n) issue(i);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int i = 0; i < n; ++i) {
            asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
            __syncwarp();
            const float *d = ring + (i % NS) * 384;
            const float4 a = *reinterpret_cast<const float4 *>(d + 4 * lane), b = *reinterpret_cast<const float4 *>(d + 128 + 4 * lane),
                         c = *reinterpret_cast<const float4 *>(d + 256 + 4 * lane);
            acc.x += a.x * b.x + c.x; acc.y += a.y * b.y + c.y; acc.z += a.z + b.z * c.z; acc.w += a.w + b.w + c.w;
            __syncwarp();
            if (i + NS < n) issue(i + NS);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    *reinterpret_cast<float4 *>(out + (size_t)r * 128 + 4 * lane) = acc;
}

template <int MODE, int NS>
static void run(const char *name, const float *kv, const float *rhat, const int *src, const int *cnt, int stride, int n_rows, float *out,
                double bytes, float *flush, size_t flush_n) {
    const size_t smem = (size_t)WARPS * NS * 1536 + WARPS * NS * 8 + 128;
    CK(cudaFuncSetAttribute(k_gather<MODE, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
        if (flush) CK(cudaMemsetAsync(flush, rep, flush_n));
        CK(cudaEventRecord(a));
        k_gather<MODE, NS><<<(n_rows + WARPS - 1) / WARPS, WARPS * 32, MODE == 0 ? 0 : smem>>>(kv, rhat, src, cnt, stride, n_rows, out);
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    printf("  %-44s %8.1f us  %7.1f GB/s\n", name, best * 1e3, bytes / (best * 1e-3) / 1e9);
}

int main() {
    for (int big = 0; big < 2; ++big) {
        // big = 0: 3,072 rows x 40 edges (configs[2]-like: kv 3 MB, rhat 63 MB);  big = 1: 32k rows x 48 edges, kv over 384 MB
        const int n_rows = big ? 32768 : 3072, stride = 64, deg = big ? 48 : 40;
        const size_t n_src = big ? 393216 : 3072;
        float *kv, *rhat, *out, *flush = nullptr;
        int *src, *cnt;
        CK(cudaMalloc(&kv, n_src * 1024)); CK(cudaMalloc(&rhat, (size_t)n_rows * stride * 512)); CK(cudaMalloc(&out, (size_t)n_rows * 512));
        CK(cudaMalloc(&src, (size_t)n_rows * stride * 4)); CK(cudaMalloc(&cnt, n_rows * 4));
        CK(cudaMemset(kv, 0, n_src * 1024)); CK(cudaMemset(rhat, 0, (size_t)n_rows * stride * 512));
        std::vector<int> hs((size_t)n_rows * stride), hc(n_rows);
        srand(1);
        for (int r = 0; r < n_rows; ++r) {
            hc[r] = deg - 8 + rand() % 17;
            // sources: rows of the same "scene" (64-row neighbourhood) for the small case, random for the big one
            for (int k = 0; k < stride; ++k) hs[(size_t)r * stride + k] = big ? (int)(((size_t)rand() * 7919 + rand()) % n_src) : (r / 96) * 96 + rand() % 96;
        }
        CK(cudaMemcpy(src, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(cnt, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
        double edges = 0;
        for (int r = 0; r < n_rows; ++r) edges += hc[r];
        const double bytes = edges * 1536.0;
        const size_t flush_n = 512u << 20;
        if (big) CK(cudaMalloc(&flush, flush_n));
        printf("%s: %d rows, %.0f edges, %.1f MB gathered%s\n", big ? "HBM-resident" : "L2-resident (warm)", n_rows, edges, bytes / 1e6, big ? " (L2 flushed)" : "");
        run<0, 2>("LDG.128, 2 edges in flight", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<0, 4>("LDG.128, 4 edges in flight", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<0, 8>("LDG.128, 8 edges in flight", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<1, 4>("cp.async.bulk, lane 0 issues, ring 4", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<1, 8>("cp.async.bulk, lane 0 issues, ring 8", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<2, 8>("cp.async.bulk, lane e%32 issues, ring 8", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<2, 16>("cp.async.bulk, lane e%32 issues, ring 16", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<3, 4>("cp.async 16 B (LDGSTS), ring 4", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        run<3, 8>("cp.async 16 B (LDGSTS), ring 8", kv, rhat, src, cnt, stride, n_rows, out, bytes, flush, flush_n);
        cudaFree(kv); cudaFree(rhat); cudaFree(out); cudaFree(src); cudaFree(cnt); if (flush) cudaFree(flush);
    }
    return 0;
}
