// Map encoder `InfGenMapDecoder.forward` (reference infgen/modules/map_decoder.py:70-130) - SURVEY.md section 8f row f1: the
// kernels that are not AttentionLayer / FourierEmbedding / MLP calls.  The three pt2pt AttentionLayers run on the row-tile
// path (k_attn + k_node, node.cuh), the relative embedding on k_fourier_tc, the token-vocabulary table on k_mlp_embed.
//
//   k_map_embed   x[p] = token_emb(traj_src)[token_idx[p]] + type_pt_emb[type] + polygon_type_emb[pl_type]
//                        + light_pl_emb[light_type]                                              (map_decoder.py:76-90)
//   k_map_edges   radius_graph(x = pos[:, :2], r = pl2pl_radius, loop = False, max_num_neighbors = 100) (:91-93) and the
//                 raw relative features of every edge (:96-104).  Third-party semantics (torch_cluster 1.6.3, as fixed by
//                 oracle/shims/cluster.py): strict dist^2 < r^2, per target the first max_num_neighbors + 1 candidates by
//                 ascending index - the target itself included - with the self loop dropped afterwards; edges ordered by
//                 target, then source.  Tokens of different scenes never connect.
#pragma once
#include "common.cuh"
#include "decode.cuh"

namespace infgen {

constexpr int MAP_MAX_NB = 100;                 // max_num_neighbors (map_decoder.py:93)
constexpr int MAP_STRIDE = MAP_MAX_NB + 1;      // edge slots per token (101 survive when the target is not among the first 101)

struct MapState {
    int n_scenes, P;
    float r2;                                   // pl2pl_radius squared
    const int *pt_ptr;                          // [n_scenes + 1]
    const float *pos, *ori;                     // [P][2], [P]
    const int *type, *pl_type, *light_type, *token_idx;   // [P]
    const int *scene_of;                        // [P] scene of every token
    // edges: slot p * MAP_STRIDE + k
    int *cnt, *start, *src;                     // [P], [P], [P * MAP_STRIDE]
    float *raw;                                 // [P * MAP_STRIDE][3]
};

__global__ void k_map_scene_of(const int *pt_ptr, int n_scenes, int *scene_of) {
    const int b = blockIdx.x;
    for (int p = pt_ptr[b] + threadIdx.x; p < pt_ptr[b + 1]; p += blockDim.x) scene_of[p] = b;
}

// one warp per token
__global__ void __launch_bounds__(NT) k_map_embed(const MapState m, const float *__restrict__ tok_tab, const float *__restrict__ type_emb,
                                                  const float *__restrict__ pl_emb, const float *__restrict__ light_emb,
                                                  float *__restrict__ x) {
    const int p = blockIdx.x * NWARP + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= m.P) return;
    const float4 t = ldg4(tok_tab + (size_t)m.token_idx[p] * 128 + 4 * lane);
    const float4 a = ldg4(type_emb + (size_t)m.type[p] * 128 + 4 * lane);
    const float4 b = ldg4(pl_emb + (size_t)m.pl_type[p] * 128 + 4 * lane);
    const float4 c = ldg4(light_emb + (size_t)m.light_type[p] * 128 + 4 * lane);
    // torch.stack([type, polygon type, light]).sum(0), then x + that (map_decoder.py:87-90)
    const float4 cat = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
    st4(x + (size_t)p * 128 + 4 * lane, make_float4(t.x + cat.x, t.y + cat.y, t.z + cat.z, t.w + cat.w));
}

// one warp per target token
__global__ void __launch_bounds__(NT) k_map_edges(const MapState m) {
    const int p = blockIdx.x * NWARP + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= m.P) return;
    const int b = m.scene_of[p], p0 = m.pt_ptr[b], p1 = m.pt_ptr[b + 1];
    const float px = m.pos[(size_t)p * 2], py = m.pos[(size_t)p * 2 + 1], od = m.ori[p];
    const float ox = cosf(od), oy = sinf(od);
    const int base = p * MAP_STRIDE;
    int cnt = 0, seen = 0;
    for (int j0 = p0; j0 < p1 && seen <= MAP_MAX_NB; j0 += 32) {
        const int j = j0 + lane;
        bool within = false;
        float rx = 0.f, ry = 0.f;
        if (j < p1) {
            rx = __fsub_rn(m.pos[(size_t)j * 2], px);              // pos[src] - pos[dst] (:96)
            ry = __fsub_rn(m.pos[(size_t)j * 2 + 1], py);
            const float dx = __fsub_rn(px, m.pos[(size_t)j * 2]), dy = __fsub_rn(py, m.pos[(size_t)j * 2 + 1]);
            within = dist2(dx, dy) < m.r2;
        }
        const unsigned wm = __ballot_sync(0xffffffffu, within);
        const bool in_first = within && (seen + __popc(wm & lanemask_lt())) <= MAP_MAX_NB;
        const bool ok = in_first && j != p;
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const int slot = base + cnt + __popc(mask & lanemask_lt());
            m.src[slot] = j;
            m.raw[(size_t)slot * 3 + 0] = norm2(rx, ry);
            m.raw[(size_t)slot * 3 + 1] = angle_between(ox, oy, rx, ry);
            m.raw[(size_t)slot * 3 + 2] = wrap_angle(__fsub_rn(m.ori[j], od));
        }
        cnt += __popc(mask);
        seen += __popc(wm);
    }
    if (lane == 0) { m.cnt[p] = cnt; m.start[p] = base; }
}

}  // namespace infgen
