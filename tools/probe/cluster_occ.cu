// Probe: how many 8-CTA clusters with ~211 KB of dynamic shared memory are co-resident on this GPU, and what a
// back-to-back launch of such a kernel costs.  nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occ cluster_occ.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
namespace cg = cooperative_groups;
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1) k8(int *out, int spin) {
    extern __shared__ float sm[];
    cg::cluster_group cl = cg::this_cluster();
    sm[threadIdx.x] = threadIdx.x;
    cl.sync();
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    cl.sync();
    if (threadIdx.x == 0 && out) out[blockIdx.x] = (int)sm[1];
}
__global__ void __launch_bounds__(256, 1) k1(int *out, int spin) {
    extern __shared__ float sm[];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (threadIdx.x == 0 && out) out[blockIdx.x] = (int)sm[1];
}
int main() {
    int smem = 211 * 1024;
    cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cs : {2, 4, 8}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k1, &cfg);
        printf("cluster size %d smem %d: max active clusters %d (%s)\n", cs, smem, n, cudaGetErrorString(e));
    }
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k8, &cfg);
        printf("k8 (__cluster_dims__ 8): max active clusters %d (%s)\n", n, cudaGetErrorString(e));
    }
    int *d; cudaMalloc(&d, 4096 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int grid : {64, 96, 104, 112, 120, 128, 136, 144}) {
        for (int spin : {0, 20000}) {
            for (int i = 0; i < 5; ++i) k8<<<grid, 256, smem>>>(d, spin);
            cudaEventRecord(a);
            for (int i = 0; i < 50; ++i) k8<<<grid, 256, smem>>>(d, spin);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("k8 grid %3d spin %5d: %.2f us/launch (%s)\n", grid, spin, ms * 1000 / 50, cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int grid : {64, 128, 148}) {
        for (int spin : {0, 20000}) {
            for (int i = 0; i < 5; ++i) k1<<<grid, 256, smem>>>(d, spin);
            cudaEventRecord(a);
            for (int i = 0; i < 50; ++i) k1<<<grid, 256, smem>>>(d, spin);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("k1 grid %3d spin %5d: %.2f us/launch\n", grid, spin, ms * 1000 / 50);
        }
    }
    // alternate big-smem cluster kernel with a small-smem plain kernel (carve-out reconfiguration?)
    for (int i = 0; i < 5; ++i) { k8<<<64, 256, smem>>>(d, 0); k1<<<64, 256, 1024>>>(d, 0); }
    cudaEventRecord(a);
    for (int i = 0; i < 50; ++i) { k8<<<64, 256, smem>>>(d, 0); k1<<<64, 256, 1024>>>(d, 0); }
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("alternating k8(211KB) + k1(1KB): %.2f us/pair\n", ms * 1000 / 50);
    return 0;
}
