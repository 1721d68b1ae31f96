// Probe: CUDA-graph conditional nodes (WHILE with a nested IF) populated by stream capture, with a thread-block-cluster
// kernel, a device-to-device memcpy and a memset inside the loop body - the control structure the insertion stage needs
// (a data-dependent number of passes per decode iteration, agent_decoder.py:1773-2105) without any host round trip.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cond_graph cond_graph.cu && ./cond_graph
//
// Expected output: "ok" lines for 3 replays with different trip counts, and the measured cost of a loop trip.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <vector>
#include <functional>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)

struct State { int *trips_left, *n_pass, *n_heading, *data, *copy; };

__global__ void k_begin(State s, cudaGraphConditionalHandle h_while) {
    if (threadIdx.x == 0) { *s.n_pass = 0; *s.n_heading = 0; cudaGraphSetConditional(h_while, *s.trips_left > 0); }
}
__global__ void __cluster_dims__(2, 1, 1) k_cluster_body(State s) {
    cg::cluster_group cl = cg::this_cluster();
    __shared__ int v;
    if (threadIdx.x == 0) v = (int)cl.block_rank() + 1;
    cl.sync();
    int *peer = cl.map_shared_rank(&v, cl.block_rank() ^ 1);
    if (threadIdx.x == 0) atomicAdd(s.data, *peer);          // +3 per pass (1 + 2)
    cl.sync();
}
__global__ void k_decide(State s, cudaGraphConditionalHandle h_while, cudaGraphConditionalHandle h_if) {
    if (threadIdx.x == 0) {
        const int left = --(*s.trips_left);
        (*s.n_pass)++;
        cudaGraphSetConditional(h_while, left > 0);
        cudaGraphSetConditional(h_if, (left & 1) == 0);      // heading stage on every other pass
    }
}
__global__ void k_heading(State s) { if (threadIdx.x == 0) (*s.n_heading)++; }
__global__ void k_after(State s, int *out) { if (threadIdx.x == 0) { out[0] = *s.n_pass; out[1] = *s.n_heading; out[2] = *s.data; out[3] = *s.copy; } }

struct Pending { cudaGraph_t body; std::function<int()> fn; };
static std::vector<Pending> g_pending;

// inside a capture on `st`: add a conditional node after the captured work, queue its body
static int add_cond(cudaStream_t st, cudaGraphConditionalHandle h, cudaGraphConditionalNodeType type, std::function<int()> fn) {
    cudaStreamCaptureStatus status;
    cudaGraph_t g;
    const cudaGraphNode_t *deps;
    size_t n_deps;
    CK(cudaStreamGetCaptureInfo(st, &status, nullptr, &g, &deps, &n_deps));
    if (status != cudaStreamCaptureStatusActive) { printf("FAIL: stream not capturing\n"); return 1; }
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = type;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, g, deps, n_deps, &p));
    CK(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
    g_pending.push_back(Pending{p.conditional.phGraph_out[0], fn});
    return 0;
}
static int drain(cudaStream_t st) {
    while (!g_pending.empty()) {
        Pending p = g_pending.back();
        g_pending.pop_back();
        CK(cudaStreamBeginCaptureToGraph(st, p.body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        if (p.fn()) return 1;
        CK(cudaStreamEndCapture(st, nullptr));
    }
    return 0;
}

int main() {
    int drv = 0, rt = 0;
    cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    printf("driver %d runtime %d\n", drv, rt);
    State s;
    int *out;
    CK(cudaMalloc(&s.trips_left, 4)); CK(cudaMalloc(&s.n_pass, 4)); CK(cudaMalloc(&s.n_heading, 4));
    CK(cudaMalloc(&s.data, 4)); CK(cudaMalloc(&s.copy, 4)); CK(cudaMalloc(&out, 16));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));

    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    cudaStreamCaptureStatus status;
    cudaGraph_t top;
    CK(cudaStreamGetCaptureInfo(st, &status, nullptr, &top, nullptr, nullptr));
    cudaGraphConditionalHandle h_while, h_if;
    CK(cudaGraphConditionalHandleCreate(&h_while, top, 0, cudaGraphCondAssignDefault));
    CK(cudaGraphConditionalHandleCreate(&h_if, top, 0, cudaGraphCondAssignDefault));
    k_begin<<<1, 32, 0, st>>>(s, h_while);
    if (add_cond(st, h_while, cudaGraphCondTypeWhile, [&]() -> int {
            CK(cudaMemsetAsync(s.copy, 0, 4, st));
            k_cluster_body<<<2, 32, 0, st>>>(s);
            CK(cudaMemcpyAsync(s.copy, s.data, 4, cudaMemcpyDeviceToDevice, st));
            k_decide<<<1, 32, 0, st>>>(s, h_while, h_if);
            return add_cond(st, h_if, cudaGraphCondTypeIf, [&]() -> int {
                k_heading<<<1, 32, 0, st>>>(s);
                k_heading<<<1, 32, 0, st>>>(s);
                return 0;
            });
        })) return 1;
    k_after<<<1, 32, 0, st>>>(s, out);
    cudaGraph_t g;
    CK(cudaStreamEndCapture(st, &g));
    if (drain(st)) return 1;
    cudaGraphExec_t ge;
    CK(cudaGraphInstantiate(&ge, g, 0));
    printf("instantiated\n");

    int total = 0;
    for (int trips : {0, 1, 4, 10}) {
        CK(cudaMemcpyAsync(s.trips_left, &trips, 4, cudaMemcpyHostToDevice, st));
        CK(cudaGraphLaunch(ge, st));
        int h[4];
        CK(cudaMemcpyAsync(h, out, 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        total += trips;
        int want_head = 0;
        for (int left = trips - 1; left >= 0; --left) want_head += ((left & 1) == 0) ? 2 : 0;
        const bool ok = h[0] == trips && h[1] == want_head && h[2] == 3 * total && (trips == 0 || h[3] == 3 * total);
        printf("%s trips %d: passes %d headings %d (want %d) data %d copy %d\n", ok ? "ok" : "MISMATCH", trips, h[0], h[1], want_head, h[2], h[3]);
    }
    // cost of a loop trip
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int trips : {0, 1, 10}) {
        float best = 1e9f;
        for (int rep = 0; rep < 20; ++rep) {
            CK(cudaMemcpyAsync(s.trips_left, &trips, 4, cudaMemcpyHostToDevice, st));
            CK(cudaEventRecord(a, st));
            CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(b, st));
            CK(cudaStreamSynchronize(st));
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (ms < best) best = ms;
        }
        printf("graph with %d trips: %.1f us\n", trips, best * 1e3f);
    }
    return 0;
}
