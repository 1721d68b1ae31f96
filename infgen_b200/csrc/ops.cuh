// Operator kernels: the reference's layers.py modules as sm_100a CUDA.
//
//   k_node_update   AttentionLayer minus the edge part (layers.py:61-76, 94-113): the "post" half of one layer
//                   (relative-value projection, gate, to_out, LayerNorm, FFN) fused with the "pre" half of the
//                   next one (LayerNorm, q/s/k/v projections, relative-query fold), one CTA per tile of rows.
//   k_edge_attn     AttentionLayer.message + segment softmax + aggregation (layers.py:78-92), one warp per
//                   destination row, K/V rows gathered from the KV caches.
//   k_kv_project    LayerNorm + to_k/to_v of source-only nodes (map tokens; layers.py:65-71, 107-108).
//   k_fourier       FourierEmbedding (layers.py:142-160) over edge/agent tiles, output optionally standardised.
//   k_mlp_embed     MLPEmbedding (layers.py:170-189), optionally gathering the 4x128 fusion input.
//   k_heads         MLPLayer token head + per-slice top-k / softmax statistics, state head (layers.py:206-215,
//                   agent_decoder.py:2160-2167).
//
// Algebra used by node_update/edge_attn: with rn_e = LN_r(r_e) = g*rhat_e + b,
//   q_h.(k_j,h + Wkr_h rn_e)       = q_h.k_j,h + (g * Wkr_h^T q_h).rhat_e + const(i,h)   (const drops out of softmax)
//   sum_e a_e (v_j + Wvr rn_e + b) = sum_e a_e v_j + Wvr_h (g * sum_e a_e rhat_e + b * sum_e a_e) + bvr * sum_e a_e
// so the per-edge 128x128 projections become per-node ones and edges only see dot products with rhat.
#pragma once
#include "common.cuh"

namespace infgen {

struct AttnW {                 // one AttentionLayer, pointers into the packed weight blob
    const float *ln_src_g, *ln_src_b, *ln_dst_g, *ln_dst_b;
    const float *w_qs, *b_qs;  // [128 -> 256] to_q | to_s
    const float *w_kv, *b_kv;  // [128 -> 256] to_k | to_v (k bias = 0)
    const float *w_kr;         // to_k_r.weight [128 out][128 in] row-major
    const float *ln_r_g, *ln_r_b;
    const float *w_vr, *b_vr;  // to_v_r packed [32][128][4]
    const float *w_g, *b_g;    // [256 -> 128]
    const float *w_out, *b_out;
    const float *ln_post_g, *ln_post_b, *ln_ffpre_g, *ln_ffpre_b;
    const float *w_ff1, *b_ff1, *w_ff2, *b_ff2;
    const float *ln_ffpost_g, *ln_ffpost_b;
    int has_pos;               // has_pos_emb
    const float *cs_post, *cs_pre;   // cluster-sliced chunks of this layer (layer.cuh), [8][FLOATS] each
};

struct FourierW {              // one FourierEmbedding
    const float *freqs;        // [D][64]
    const float *w0[4], *b0[4], *ln_g[4], *ln_b[4], *w3[4], *b3[4];
    const float *out_ln_g, *out_ln_b, *w_out, *b_out;
};

struct MlpEmbW {               // one MLPEmbedding
    const float *w0, *b0, *ln1_g, *ln1_b, *w3, *b3, *ln4_g, *ln4_b, *w6, *b6;
};

struct MlpHeadW {              // one MLPLayer (in -> 128 -> out)
    const float *w0, *b0, *ln_g, *ln_b, *w3, *b3;
    int k4_in;                 // packed K4 of the first Linear
    int n_out, n_pad;          // real / padded output width
};

// Rows of a batch live in a capacity row space: scene b owns rows [b*cap, b*cap + n_rows[b]).
struct RowSpace {
    int n_total;               // rows in the space (n_scenes * cap, or n for operator-level calls)
    int cap;                   // 0: all rows < n_total are active
    const int *n_rows;         // [n_scenes] device
    __device__ __forceinline__ bool active(int r) const {
        if (r >= n_total) return false;
        if (cap == 0) return true;
        return (r % cap) < __ldg(n_rows + r / cap);
    }
};

// ===============================================================================================================
// K|V projection of source-only nodes: out[l][n][256] = [Wk LN_src(x[n]) | Wv LN_src(x[n]) + bv], blockIdx.y = l
// ===============================================================================================================
struct KvArgs {
    int n;
    const float *x;            // [n][128]
    AttnW w[6];
    float *out[6];             // [n][256] each
};

template <int M>
__global__ void __launch_bounds__(NT) k_kv_project(const KvArgs a) {
    __shared__ __align__(16) float sx[M * 128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * M;
    const AttnW &w = a.w[blockIdx.y];
    float *out = a.out[blockIdx.y];
    for (int m = warp; m < M; m += NWARP) {
        const int r = row0 + m;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < a.n) x = ld4(a.x + (size_t)r * 128 + 4 * lane);
        st4(sx + m * 128 + 4 * lane, ln128(x, w.ln_src_g, w.ln_src_b, lane));
    }
    __syncthreads();
    block_gemm<M, 256>(sx, 128, w.w_kv, 256, 32, nullptr, [&](int m, int n, float v) {
        const int r = row0 + m;
        if (r < a.n) out[(size_t)r * 256 + n] = v + __ldg(w.b_kv + n);
    });
}

// ===============================================================================================================
// FourierEmbedding over tiles of FM slots.  slot s is valid iff cnt == NULL ? s < n_slots : (s % stride) < cnt[s / stride]
// ===============================================================================================================
constexpr int FM = 32;
constexpr int FLD = 132;       // 129 Fourier features padded to a multiple of 4

struct FourierArgs {
    int n_slots;
    const int *cnt;
    int stride;
    const float *raw;          // [slots][D]
    FourierW w;
    const float *cat_tab;      // optional categorical sum rows [.][128]
    const int *cat_idx;        // [slots] row of cat_tab (when cat_tab != NULL; NULL -> row = slot)
    float *out;                // [slots][128]
    int normalize;             // 1: store (y - mean) / std of the output (input of every layer's attn_prenorm_r)
};

template <int D>
__global__ void __launch_bounds__(NT) k_fourier(const FourierArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *sF = smem;                       // [FM][132]
    float *sH = sF + FM * FLD;              // [FM][128]
    float *sA = sH + FM * 128;              // [FM][128]
    float *sred = sA + FM * 128;            // [FM][128]
    float *sraw = sred + FM * 128;          // [FM][4]
    __shared__ int s_valid[FM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = blockIdx.x * FM;
    int v = 0;
    if (tid < FM) {
        const int s = s0 + tid;
        if (s < a.n_slots) v = a.cnt ? ((s % a.stride) < a.cnt[s / a.stride]) : 1;
        s_valid[tid] = v;
#pragma unroll
        for (int d = 0; d < D; ++d) sraw[tid * 4 + d] = v ? a.raw[(size_t)s * D + d] : 0.f;
    }
    if (!__syncthreads_or(v)) return;
    // categorical sum seeds the accumulator (layers.py:156-159)
    for (int m = warp; m < FM; m += NWARP) {
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.cat_tab && s_valid[m]) {
            const int row = a.cat_idx ? a.cat_idx[s0 + m] : (s0 + m);
            c = ld4(a.cat_tab + (size_t)row * 128 + 4 * lane);
        }
        st4(sA + m * 128 + 4 * lane, c);
    }
    const FourierW &w = a.w;
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        __syncthreads();
        for (int i = tid; i < FM * 64; i += NT) {
            const int m = i >> 6, j = i & 63;
            const float x = sraw[m * 4 + d];
            // x.unsqueeze(-1) * freqs * 2 * math.pi, evaluated left to right in fp32 (layers.py:151)
            const float arg = __fmul_rn(__fmul_rn(__fmul_rn(x, __ldg(w.freqs + d * 64 + j)), 2.0f), 3.14159265358979323846f);
            float sn, cs;
            sincosf(arg, &sn, &cs);
            sF[m * FLD + j] = cs;
            sF[m * FLD + 64 + j] = sn;
        }
        if (tid < FM) {
            sF[tid * FLD + 128] = sraw[tid * 4 + d];
            sF[tid * FLD + 129] = 0.f; sF[tid * FLD + 130] = 0.f; sF[tid * FLD + 131] = 0.f;
        }
        __syncthreads();
        block_gemm<FM, 128>(sF, FLD, w.w0[d], 128, 33, sred,
                            [&](int m, int n, float v2) { sH[m * 128 + n] = v2 + __ldg(w.b0[d] + n); });
        __syncthreads();
        rows_layernorm<FM, true>(sH, 128, w.ln_g[d], w.ln_b[d]);
        __syncthreads();
        block_gemm<FM, 128>(sH, 128, w.w3[d], 128, 32, sred,
                            [&](int m, int n, float v2) { sA[m * 128 + n] += v2 + __ldg(w.b3[d] + n); });
    }
    __syncthreads();
    rows_layernorm<FM, true>(sA, 128, w.out_ln_g, w.out_ln_b);
    __syncthreads();
    block_gemm<FM, 128>(sA, 128, w.w_out, 128, 32, sred,
                        [&](int m, int n, float v2) { sH[m * 128 + n] = v2 + __ldg(w.b_out + n); });
    __syncthreads();
    for (int m = warp; m < FM; m += NWARP) {
        if (!s_valid[m]) continue;
        float4 y = ld4(sH + m * 128 + 4 * lane);
        if (a.normalize) {
            float mean, rstd;
            ln_stats(y, mean, rstd);
            y = make_float4((y.x - mean) * rstd, (y.y - mean) * rstd, (y.z - mean) * rstd, (y.w - mean) * rstd);
        }
        st4(a.out + (size_t)(s0 + m) * 128 + 4 * lane, y);
    }
}
constexpr size_t FOURIER_SMEM = (size_t)(FM * FLD + 3 * FM * 128 + FM * 4) * sizeof(float);

// ===============================================================================================================
// MLPEmbedding: x[n][kin] -> 128 (LN, ReLU) -> 128 (LN, ReLU) -> 128.  Fusion mode gathers the four 128-blocks
// [token_emb | x_a_emb | state_emb | grid_emb] of agent_decoder.py:503-507 / 2282-2286 instead of reading x.
// ===============================================================================================================
constexpr int EM = 16;

struct MlpEmbArgs {
    RowSpace rows;
    MlpEmbW w;
    int kin;                   // real input width
    int k4;                    // packed K4 of the first Linear
    const float *x;            // [n][kin] (plain mode), row stride x_ld floats
    int x_ld;
    // fusion mode
    int fusion;
    const float *tok_tab;      // [3][token_size+2][128]
    const int *tok_row;        // [R] row in tok_tab (type*(token_size+2) + index)
    const float *xa;           // [R][128]
    const float *state_tab;    // [4][128]
    const int *state_idx;      // [R]
    const float *grid_tab;     // [grid_size+1][128]
    const int *grid_row;       // [R]
    float *out;                // [n][128]
    int out_ld;                // floats between output rows (128, or more to scatter into a wider table)
};

__global__ void __launch_bounds__(NT) k_mlp_embed(const MlpEmbArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int ldx = a.k4 * 4;
    float *sX = smem;                        // [EM][ldx]
    float *sH = sX + EM * ldx;               // [EM][128]
    float *sG = sH + EM * 128;               // [EM][128]
    float *sred = sG + EM * 128;             // [EM][128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * EM;
    bool any = false;
    for (int m = 0; m < EM; ++m) any |= a.rows.active(row0 + m);
    if (!any) return;
    if (a.fusion) {
        for (int m = warp; m < EM; m += NWARP) {
            const int r = row0 + m;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f), x = t, s = t, g = t;
            if (a.rows.active(r)) {
                t = ld4(a.tok_tab + (size_t)a.tok_row[r] * 128 + 4 * lane);
                x = ld4(a.xa + (size_t)r * 128 + 4 * lane);
                s = ld4(a.state_tab + (size_t)a.state_idx[r] * 128 + 4 * lane);
                g = ld4(a.grid_tab + (size_t)a.grid_row[r] * 128 + 4 * lane);
            }
            st4(sX + m * ldx + 4 * lane, t);
            st4(sX + m * ldx + 128 + 4 * lane, x);
            st4(sX + m * ldx + 256 + 4 * lane, s);
            st4(sX + m * ldx + 384 + 4 * lane, g);
        }
    } else {
        for (int i = tid; i < EM * ldx; i += NT) {
            const int m = i / ldx, k = i % ldx;
            const int r = row0 + m;
            sX[i] = (k < a.kin && a.rows.active(r)) ? a.x[(size_t)r * a.x_ld + k] : 0.f;
        }
    }
    __syncthreads();
    const MlpEmbW &w = a.w;
    block_gemm<EM, 128>(sX, ldx, w.w0, 128, a.k4, sred,
                        [&](int m, int n, float v) { sH[m * 128 + n] = v + __ldg(w.b0 + n); });
    __syncthreads();
    rows_layernorm<EM, true>(sH, 128, w.ln1_g, w.ln1_b);
    __syncthreads();
    block_gemm<EM, 128>(sH, 128, w.w3, 128, 32, sred,
                        [&](int m, int n, float v) { sG[m * 128 + n] = v + __ldg(w.b3 + n); });
    __syncthreads();
    rows_layernorm<EM, true>(sG, 128, w.ln4_g, w.ln4_b);
    __syncthreads();
    block_gemm<EM, 128>(sG, 128, w.w6, 128, 32, sred, [&](int m, int n, float v) {
        const int r = row0 + m;
        if (a.rows.active(r)) a.out[(size_t)r * a.out_ld + n] = v + __ldg(w.b6 + n);
    });
}
static inline size_t mlp_embed_smem(int k4) { return (size_t)(EM * k4 * 4 + 3 * EM * 128) * sizeof(float); }

// ===============================================================================================================
// heads: token_predict_head logits for one 256-wide vocabulary slice + per-slice top-KTOP / max / sum-exp, and
// (slice 0) the state head.  grid = (row tiles, vocab slices)
// ===============================================================================================================
constexpr int HM = 8;          // rows per CTA
constexpr int KTOP = 5;        // candidates kept per slice (>= motion_beam_size)
constexpr int NSLICE = 8;      // 2048 / 256

struct HeadArgs {
    RowSpace rows;
    const float *x;            // [R][128] last layer output at the current column
    MlpHeadW tok, st;
    float *part_v;             // [R][NSLICE][KTOP]
    int *part_i;               // [R][NSLICE][KTOP]
    float *part_m, *part_s;    // [R][NSLICE] slice max / sum exp(l - max)
    float *state_logits;       // [R][4]
    float *trace_head_in;      // optional [R][128]
    float *trace_logits;       // optional [R][2048]
    float *trace_state;        // optional [R][3]
};

__global__ void __launch_bounds__(NT) k_heads(const HeadArgs a) {
    __shared__ __align__(16) float sx[HM * 128];
    __shared__ __align__(16) float sh[HM * 128];
    __shared__ __align__(16) float sred[HM * 128];
    __shared__ __align__(16) float slog[HM * 256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * HM, slice = blockIdx.y;
    bool any = false;
    for (int m = 0; m < HM; ++m) any |= a.rows.active(row0 + m);
    if (!any) return;
    for (int m = warp; m < HM; m += NWARP) {
        const int r = row0 + m;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.rows.active(r)) {
            x = ld4(a.x + (size_t)r * 128 + 4 * lane);
            if (slice == 0 && a.trace_head_in) st4(a.trace_head_in + (size_t)r * 128 + 4 * lane, x);
        }
        st4(sx + m * 128 + 4 * lane, x);
    }
    __syncthreads();
    block_gemm<HM, 128>(sx, 128, a.tok.w0, 128, 32, sred,
                        [&](int m, int n, float v) { sh[m * 128 + n] = v + __ldg(a.tok.b0 + n); });
    __syncthreads();
    rows_layernorm<HM, true>(sh, 128, a.tok.ln_g, a.tok.ln_b);
    __syncthreads();
    block_gemm<HM, 256>(sh, 128, a.tok.w3 + (size_t)slice * 256 * 4, a.tok.n_pad, 32, sred, [&](int m, int n, float v) {
        v += __ldg(a.tok.b3 + slice * 256 + n);
        slog[m * 256 + n] = v;
        const int r = row0 + m;
        if (a.trace_logits && a.rows.active(r)) a.trace_logits[(size_t)r * a.tok.n_out + slice * 256 + n] = v;
    });
    __syncthreads();
    // per-row top-KTOP of the slice: one warp per row, 8 values per lane
    for (int m = warp; m < HM; m += NWARP) {
        const int r = row0 + m;
        if (!a.rows.active(r)) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = slog[m * 256 + lane + 32 * j];
        float smax = 0.f, ssum = 0.f;
#pragma unroll 1
        for (int k = 0; k < KTOP; ++k) {
            float bv = v[0];
            int bi = lane;
#pragma unroll
            for (int j = 1; j < 8; ++j)
                if (v[j] > bv) { bv = v[j]; bi = lane + 32 * j; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (k == 0) {
                smax = bv;
                float e = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) e += expf(v[j] - smax);
                ssum = warp_sum(e);
            }
            if ((bi & 31) == lane) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (bi == lane + 32 * j) v[j] = -INFINITY;
            }
            if (lane == 0) {
                a.part_v[((size_t)r * NSLICE + slice) * KTOP + k] = bv;
                a.part_i[((size_t)r * NSLICE + slice) * KTOP + k] = slice * 256 + bi;
            }
        }
        if (lane == 0) {
            a.part_m[(size_t)r * NSLICE + slice] = smax;
            a.part_s[(size_t)r * NSLICE + slice] = ssum;
        }
    }
    if (slice != 0) return;
    // state head (agent_decoder.py:2166)
    __syncthreads();
    block_gemm<HM, 128>(sx, 128, a.st.w0, 128, 32, sred,
                        [&](int m, int n, float v) { sh[m * 128 + n] = v + __ldg(a.st.b0 + n); });
    __syncthreads();
    rows_layernorm<HM, true>(sh, 128, a.st.ln_g, a.st.ln_b);
    __syncthreads();
    block_gemm<HM, 128>(sh, 128, a.st.w3, 128, 32, sred, [&](int m, int n, float v) {
        const int r = row0 + m;
        if (n < a.st.n_out && a.rows.active(r)) {
            v += __ldg(a.st.b3 + n);
            a.state_logits[(size_t)r * 4 + n] = v;
            if (a.trace_state) a.trace_state[(size_t)r * 3 + n] = v;
        }
    });
}

// rhat = (r - mean) / sqrt(var + eps): the layer-independent part of every attn_prenorm_r (one warp per row)
__global__ void k_standardize(const float *r, float *out, int n) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    float4 y = ld4(r + (size_t)row * 128 + 4 * lane);
    float mean, rstd;
    ln_stats(y, mean, rstd);
    st4(out + (size_t)row * 128 + 4 * lane,
        make_float4((y.x - mean) * rstd, (y.y - mean) * rstd, (y.z - mean) * rstd, (y.w - mean) * rstd));
}

// Generic MLPLayer for operator-level parity (any n_out that is a multiple of 128 after padding)
struct MlpLayerArgs {
    int n;
    const float *x;
    MlpHeadW w;
    float *out;                // [n][n_out]
};
__global__ void __launch_bounds__(NT) k_mlp_layer(const MlpLayerArgs a) {
    __shared__ __align__(16) float sx[HM * 128];
    __shared__ __align__(16) float sh[HM * 128];
    __shared__ __align__(16) float sred[HM * 128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * HM;
    for (int m = warp; m < HM; m += NWARP) {
        const int r = row0 + m;
        st4(sx + m * 128 + 4 * lane, r < a.n ? ld4(a.x + (size_t)r * 128 + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    __syncthreads();
    block_gemm<HM, 128>(sx, 128, a.w.w0, 128, a.w.k4_in, sred,
                        [&](int m, int n, float v) { sh[m * 128 + n] = v + __ldg(a.w.b0 + n); });
    __syncthreads();
    rows_layernorm<HM, true>(sh, 128, a.w.ln_g, a.w.ln_b);
    __syncthreads();
    for (int n0 = blockIdx.y * 128; n0 < a.w.n_pad; n0 += gridDim.y * 128) {
        block_gemm<HM, 128>(sh, 128, a.w.w3 + (size_t)n0 * 4, a.w.n_pad, 32, sred, [&](int m, int n, float v) {
            const int r = row0 + m;
            if (r < a.n && n0 + n < a.w.n_out) a.out[(size_t)r * a.w.n_out + n0 + n] = v + __ldg(a.w.b3 + n0 + n);
        });
        __syncthreads();
    }
}

}  // namespace infgen
