// Probe: throughput / latency of cp.async.bulk global->shared rings on one SM and on all SMs.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *d, const void *s, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)), "l"(s), "r"(n), "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(b)), "r"(par) : "memory"); }

// mode 0: dedicated producer warp (warp 8), 8 consumer warps wait+release.  mode 1: like 0 but consumers also read the stage.
__global__ void __launch_bounds__(288) k_ring(const float *src, size_t span_floats, int n_chunks, int chunk_bytes, int stages, long long *out, float *sink, int mode) {
    extern __shared__ __align__(128) float sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm), *empty = full + 16;
    float *ring = sm + 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const long long t0 = clock64();
    const float *base = src + ((size_t)blockIdx.x * 1237 * 4096) % span_floats;
    if (warp == 8) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&empty[st], ph ^ 1u);
                mbar_expect_tx(&full[st], chunk_bytes);
                bulk_g2s(ring + (size_t)st * (chunk_bytes / 4), base + ((size_t)c * (chunk_bytes / 4)) % (span_floats / 2), chunk_bytes, &full[st]);
                if (++st == stages) { st = 0; ph ^= 1u; }
            }
        }
        return;
    }
    int st = 0; uint32_t ph = 0; float acc = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(&full[st], ph);
        if (mode == 1) { const float *p = ring + (size_t)st * (chunk_bytes / 4); for (int i = threadIdx.x * 4; i < chunk_bytes / 4; i += 1024) { float4 v = *reinterpret_cast<const float4 *>(p + i); acc += v.x + v.y + v.z + v.w; } }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == stages) { st = 0; ph ^= 1u; }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
    if (acc == 123.456f) sink[0] = acc;
}
int main() {
    size_t span = (size_t)64 << 20;   // 256 MB of floats? no: 64M floats = 256 MB; use 16M floats = 64 MB (L2 resident)
    span = (size_t)16 << 20;
    float *src; cudaMalloc(&src, span * 4); cudaMemset(src, 0, span * 4);
    long long *out; cudaMalloc(&out, 148 * 8); float *sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[148];
    for (int grid : {1, 148}) for (int mode : {0, 1}) for (int chunk : {2048, 16384, 32768}) for (int stages : {1, 2, 3, 4, 6}) {
        if ((size_t)chunk * stages > 190 * 1024) continue;
        int n = 256;
        size_t smem = 256 + (size_t)chunk * stages;
        for (int rep = 0; rep < 3; ++rep) k_ring<<<grid, 288, smem>>>(src, span, n, chunk, stages, out, sink, mode);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d mode %d chunk %5d stages %d: %8lld cycles, %.1f cycles/chunk, %.2f B/cycle/SM  (%s)\n", grid, mode, chunk, stages, mx, (double)mx / n, (double)chunk * n / mx, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
