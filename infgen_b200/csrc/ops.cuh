// Operator kernels: the reference's layers.py modules as sm_100a CUDA (the AttentionLayer itself lives in layer.cuh).
//
//   k_kv_project    LayerNorm + to_k/to_v of source-only nodes (map tokens; layers.py:65-71, 107-108).
//   k_fourier       FourierEmbedding (layers.py:142-160) over tiles of edge slots, several embeddings per launch, output
//                   optionally standardised (the layer-independent part of every attn_prenorm_r).
//   k_mlp_embed     MLPEmbedding (layers.py:170-189).
//   k_heads         MLPLayer token head + per-slice top-k / softmax statistics, state head (layers.py:206-215,
//                   agent_decoder.py:2160-2167).
//   k_mlp_layer     generic MLPLayer (insertion-stage heads, operator-level parity).
// All of them stream their weights through the shared-memory ring of stream.cuh (cp.async.bulk + mbarriers).
//
// Algebra shared with layer.cuh: with rn_e = LN_r(r_e) = g*rhat_e + b,
//   q_h.(k_j,h + Wkr_h rn_e)       = q_h.k_j,h + (g * Wkr_h^T q_h).rhat_e + const(i,h)   (const drops out of softmax)
//   sum_e a_e (v_j + Wvr rn_e + b) = sum_e a_e v_j + Wvr_h (g * sum_e a_e rhat_e + b * sum_e a_e) + bvr * sum_e a_e
// so the per-edge 128x128 projections become per-node ones and edges only see dot products with rhat.
#pragma once
#include "common.cuh"
#include "stream.cuh"

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
    const float *npk;                // node-packed copy (node.cuh), NULL when the layer has none
    const float *vrf;                // folded to_v_r table for the k_attn epilogue (node.cuh vrf::), NULL when the layer has none
    const float *tcimg;              // tensor-core weight image for k_node_tc (node_tc.cuh ntc::), NULL when the layer has none
};

struct FourierW {              // one FourierEmbedding
    const float *freqs;        // [D][64]
    const float *w0[4], *b0[4], *ln_g[4], *ln_b[4], *w3[4], *b3[4];
    const float *out_ln_g, *out_ln_b, *w_out, *b_out;
    const float *wimg;         // tensor-core weight image (fourier_tc.cuh): hi/lo TF32 chunks in consumption order
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
    const int *row_lo;         // optional [n_scenes]: rows below it are skipped ("new rows only" launches)
    // optional compact row list (the rows appended by the last insertion pass, one per scene at most): tile t of a kernel
    // that processes M rows per tile then holds list[t*M .. t*M+M), entries >= *n_list are inactive - a handful of tiles
    // for a whole batch instead of one (mostly empty) tile per M rows of the row space
    const int *list, *n_list;
    int list_cap;              // upper bound of *n_list (grid sizing)
    // global row of position m of tile `tile` (-1: none)
    __device__ __forceinline__ int tile_row(int tile, int m, int M) const {
        if (!list) return tile * M + m;
        const int i = tile * M + m;
        return i < *n_list ? list[i] : -1;
    }
    __device__ __forceinline__ bool active_row(int r) const { return r >= 0 && (list ? true : active(r)); }
    __device__ __forceinline__ bool active(int r) const {
        if (r >= n_total) return false;
        if (cap == 0) return true;
        const int b = r / cap, i = r - b * cap;
        if (row_lo && i < row_lo[b]) return false;
        return i < n_rows[b];
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
// FourierEmbedding (layers.py:142-160) over tiles of FM slots, weights streamed (stream.cuh).
// slot s is valid iff cnt == NULL ? s < n_slots : (s % stride) < cnt[s / stride]
// ===============================================================================================================
constexpr int FM = 16;
constexpr int FLD = 132;       // 129 Fourier features padded to a multiple of 4
constexpr int HLD = 132;       // leading dimension of 128-wide activation tiles (132 % 32 == 4: conflict-free row lanes)

struct FourierArgs {
    int n_slots;
    const int *cnt;
    int stride;
    int dim;                   // input_dim (2..4)
    const float *raw;          // [slots][dim]
    FourierW w;
    const float *cat_tab;      // optional categorical sum rows [.][128]
    const int *cat_idx;        // [slots] row of cat_tab (when cat_tab != NULL; NULL -> row = slot)
    float *out;                // [slots][128]
    int normalize;             // 1: store (y - mean) / std of the output (input of every layer's attn_prenorm_r)
    const int *slot_list;      // optional compact list of the valid slots (k_slot_compact); tiles walk it instead of
    const int *n_list;         //   the strided slot space: *n_list entries
    int ffma;                  // 1: the FFMA kernel even where the tensor-core one applies (a handful of slots: its GEMV path)
    int raw_stride;            // floats per slot in `raw` (0: dim)
    const float *dim_table;    // optional [table_n][128]: the per-dim MLP output of input dim `dim` (the one after the
    int table_n;               //   last tensor-core dim) for the inputs -1, -2, .. -table_n (k_fourier_tc only)
};
// several embeddings in one launch: CTA b serves job j with tile0[j] <= b < tile0[j+1]
struct FourierBatch {
    int n_jobs;
    int tile0[4];
    FourierArgs job[3];
};

__device__ __forceinline__ int fourier_segs(const FourierW &w, int dim, WSeg *segs) {
    int n = 0;
    for (int d = 0; d < dim; ++d) {
        segs[n++] = WSeg{w.w0[d], 33, 512};
        segs[n++] = WSeg{w.w3[d], 32, 512};
    }
    segs[n++] = WSeg{w.w_out, 32, 512};
    return n;
}

// consumers: embedding of M rows whose raw inputs sit in sraw[m][4]; sA ([M][HLD], as sH) must hold the categorical seed (or
// zeros); the result (before any standardisation) is left in sH.  Ends with a csync().
template <int M>
__device__ __forceinline__ void fourier_body(WsCons &ws, const FourierW &w, int dim, const float *sraw, float *sF,
                                             float *sH, float *sA, bool single = false) {
    const int tid = threadIdx.x;
    for (int d = 0; d < dim; ++d) {
        csync();
        for (int i = tid; i < M * 64; i += NT) {
            const int m = i >> 6, j = i & 63;
            const float x = sraw[m * 4 + d];
            // x.unsqueeze(-1) * freqs * 2 * math.pi, evaluated left to right in fp32 (layers.py:151)
            const float arg = __fmul_rn(__fmul_rn(__fmul_rn(x, __ldg(w.freqs + d * 64 + j)), 2.0f), 3.14159265358979323846f);
            float sn, cs;
            sincosf(arg, &sn, &cs);
            sF[m * FLD + j] = cs;
            sF[m * FLD + 64 + j] = sn;
        }
        if (tid < M) {
            sF[tid * FLD + 128] = sraw[tid * 4 + d];
            sF[tid * FLD + 129] = 0.f; sF[tid * FLD + 130] = 0.f; sF[tid * FLD + 131] = 0.f;
        }
        csync();
        tile_gemm<M>(ws, sF, FLD, 33, [&](int m, int n, float v) { sH[m * HLD + n] = v + __ldg(w.b0[d] + n); }, single);
        csync();
        rows_layernorm_c<M, true>(sH, HLD, w.ln_g[d], w.ln_b[d]);
        csync();
        tile_gemm<M>(ws, sH, HLD, 32, [&](int m, int n, float v) { sA[m * HLD + n] += v + __ldg(w.b3[d] + n); }, single);
    }
    csync();
    rows_layernorm_c<M, true>(sA, HLD, w.out_ln_g, w.out_ln_b);
    csync();
    tile_gemm<M>(ws, sA, HLD, 32, [&](int m, int n, float v) { sH[m * HLD + n] = v + __ldg(w.b_out + n); }, single);
    csync();
}

constexpr int FOURIER_SMEM_FLOATS = WS_SMEM_FLOATS + FM * FLD + 2 * FM * HLD + FM * 4 + FM;
constexpr size_t FOURIER_SMEM = (size_t)FOURIER_SMEM_FLOATS * sizeof(float);

__global__ void __launch_bounds__(NT_S) k_fourier(const FourierBatch fb) {
    extern __shared__ __align__(16) float smem[];
    WsSmem wsm(smem);
    float *sF = smem + WS_SMEM_FLOATS;      // [FM][132]
    float *sH = sF + FM * FLD;              // [FM][HLD]
    float *sA = sH + FM * HLD;              // [FM][HLD]
    float *sraw = sA + FM * HLD;            // [FM][4]
    int *s_valid = reinterpret_cast<int *>(sraw + FM * 4);
    int j = 0;
    while (j + 1 < fb.n_jobs && (int)blockIdx.x >= fb.tile0[j + 1]) ++j;
    const FourierArgs &a = fb.job[j];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s0 = ((int)blockIdx.x - fb.tile0[j]) * FM;
    __shared__ int s_slot[FM];              // slot of every tile position (a compact slot list redirects it)
    int v = 0;
    if (tid < FM) {
        int s = s0 + tid;
        if (a.slot_list) {
            v = s < *a.n_list;
            s = v ? a.slot_list[s] : 0;
        } else if (s < a.n_slots) {
            v = a.cnt ? ((s % a.stride) < a.cnt[s / a.stride]) : 1;
        }
        s_valid[tid] = v;
        s_slot[tid] = s;
        for (int d = 0; d < 4; ++d) sraw[tid * 4 + d] = (v && d < a.dim) ? a.raw[(size_t)s * a.dim + d] : 0.f;
    }
    if (!__syncthreads_or(v)) return;
    bool single = s_valid[0] != 0;          // only the first position live: GEMV path of the tile products
    for (int m = 1; m < FM; ++m) single = single && !s_valid[m];
    ws_init(wsm);
    if (warp == NWARP) {
        if (lane < WS_STAGES) {
            WSeg segs[WS_MAX_SEGS];
            const int n = fourier_segs(a.w, a.dim, segs);
            ws_produce(wsm, segs, n);
        }
        return;
    }
    WsCons ws(wsm);
    // categorical sum seeds the accumulator (layers.py:156-159)
    for (int m = warp; m < FM; m += NWARP) {
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.cat_tab && s_valid[m]) {
            const int row = a.cat_idx ? a.cat_idx[s_slot[m]] : s_slot[m];
            c = ld4(a.cat_tab + (size_t)row * 128 + 4 * lane);
        }
        st4(sA + m * HLD + 4 * lane, c);
    }
    fourier_body<FM>(ws, a.w, a.dim, sraw, sF, sH, sA, single);
    for (int m = warp; m < FM; m += NWARP) {
        if (!s_valid[m]) continue;
        float4 y = ld4(sH + m * HLD + 4 * lane);
        if (a.normalize) {
            float mean, rstd;
            ln_stats(y, mean, rstd);
            y = make_float4((y.x - mean) * rstd, (y.y - mean) * rstd, (y.z - mean) * rstd, (y.w - mean) * rstd);
        }
        st4(a.out + (size_t)s_slot[m] * 128 + 4 * lane, y);
    }
}

// ===============================================================================================================
// MLPEmbedding (layers.py:170-189): x[n][kin] -> 128 (LN, ReLU) -> 128 (LN, ReLU) -> 128, tiles of EM rows
// ===============================================================================================================
constexpr int EM = 8;

__device__ __forceinline__ int mlp3_segs(const MlpEmbW &w, int k4, WSeg *segs) {
    segs[0] = WSeg{w.w0, k4, 512};
    segs[1] = WSeg{w.w3, 32, 512};
    segs[2] = WSeg{w.w6, 32, 512};
    return 3;
}
// consumers: sX [M][ldx] -> epi(m, n, value); sH, sG: [M][HLD] scratch.  The caller csync()s after filling sX.
template <int M, typename Epi>
__device__ __forceinline__ void mlp3_body(WsCons &ws, const MlpEmbW &w, const float *sX, int ldx, int k4, float *sH,
                                          float *sG, Epi epi, bool single = false) {
    tile_gemm<M>(ws, sX, ldx, k4, [&](int m, int n, float v) { sH[m * HLD + n] = v + __ldg(w.b0 + n); }, single);
    csync();
    rows_layernorm_c<M, true>(sH, HLD, w.ln1_g, w.ln1_b);
    csync();
    tile_gemm<M>(ws, sH, HLD, 32, [&](int m, int n, float v) { sG[m * HLD + n] = v + __ldg(w.b3 + n); }, single);
    csync();
    rows_layernorm_c<M, true>(sG, HLD, w.ln4_g, w.ln4_b);
    csync();
    tile_gemm<M>(ws, sG, HLD, 32, [&](int m, int n, float v) { epi(m, n, v + __ldg(w.b6 + n)); }, single);
}
// leading dimension of a [rows][4*k4] input tile: padded so that it is 4 (mod 32) floats when wide
__host__ __device__ __forceinline__ int mlp_ldx(int k4) { return k4 * 4 + ((k4 * 4) % 32 == 0 ? 4 : 0); }

struct MlpEmbArgs {
    RowSpace rows;
    MlpEmbW w;
    int kin;                   // real input width
    int k4;                    // packed K4 of the first Linear
    const float *x;            // [n][kin], row stride x_ld floats
    int x_ld;
    float *out;                // [n][128]
    int out_ld;                // floats between output rows (128, or more to scatter into a wider table)
};

__global__ void __launch_bounds__(NT_S) k_mlp_embed(const MlpEmbArgs a) {
    extern __shared__ __align__(16) float smem[];
    WsSmem wsm(smem);
    const int ldx = mlp_ldx(a.k4);
    float *sX = smem + WS_SMEM_FLOATS;       // [EM][ldx]
    float *sH = sX + EM * ldx;               // [EM][HLD]
    float *sG = sH + EM * HLD;               // [EM][HLD]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ int s_row[EM];                // global row of every tile position, -1 = inactive
    if (tid < EM) {
        const int r = a.rows.tile_row(blockIdx.x, tid, EM);
        s_row[tid] = a.rows.active_row(r) ? r : -1;
    }
    __syncthreads();
    bool any = false;
    for (int m = 0; m < EM; ++m) any |= s_row[m] >= 0;
    if (!any) return;
    ws_init(wsm);
    if (warp == NWARP) {
        if (lane < WS_STAGES) {
            WSeg segs[3];
            ws_produce(wsm, segs, mlp3_segs(a.w, a.k4, segs));
        }
        return;
    }
    WsCons ws(wsm);
    for (int i = tid; i < EM * ldx; i += NT) {
        const int m = i / ldx, k = i % ldx;
        const int r = s_row[m];
        sX[i] = (k < a.kin && r >= 0) ? a.x[(size_t)r * a.x_ld + k] : 0.f;
    }
    csync();
    bool single = s_row[0] >= 0;
    for (int m = 1; m < EM; ++m) single = single && s_row[m] < 0;
    mlp3_body<EM>(ws, a.w, sX, ldx, a.k4, sH, sG, [&](int m, int n, float v) {
        const int r = s_row[m];
        if (r >= 0) a.out[(size_t)r * a.out_ld + n] = v;
    }, single);
}
static inline size_t mlp_embed_smem(int k4) { return (size_t)(WS_SMEM_FLOATS + EM * mlp_ldx(k4) + 2 * EM * HLD) * sizeof(float); }

// ===============================================================================================================
// heads (agent_decoder.py:2160-2167): grid = (row tiles of HM, NSLICE + 1).  y < NSLICE: token_predict_head hidden
// layer + the logits of vocabulary slice y (256 wide) + per-slice top-KTOP / max / sum-exp;  y == NSLICE: state head.
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
// HT rows per CTA: 8 for a single scene (latency), 16 for batches (the weight stream and its LSU cost per stage are
// amortised over twice the rows)
template <int HT>
constexpr size_t heads_smem() { return (size_t)(WS_SMEM_FLOATS + 2 * HT * HLD + HT * 256) * sizeof(float); }
constexpr size_t HEADS_SMEM = heads_smem<HM>();

template <int HT>
__global__ void __launch_bounds__(NT_S) k_heads(const HeadArgs a) {
    extern __shared__ __align__(16) float smem[];
    WsSmem wsm(smem);
    float *sx = smem + WS_SMEM_FLOATS;       // [HT][HLD]
    float *sh = sx + HT * HLD;               // [HT][HLD]
    float *slog = sh + HT * HLD;             // [HT][256]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * HT, slice = blockIdx.y;
    const bool is_state = slice == NSLICE;
    const MlpHeadW &hw = is_state ? a.st : a.tok;
    bool any = false;
    for (int m = 0; m < HT; ++m) any |= a.rows.active(row0 + m);
    if (!any) return;
    ws_init(wsm);
    if (warp == NWARP) {
        if (lane < WS_STAGES) {
            WSeg segs[3];
            segs[0] = WSeg{hw.w0, 32, 512};
            int n = 1;
            if (is_state) {
                segs[n++] = WSeg{hw.w3, 32, hw.n_pad * 4};
            } else {
                segs[n++] = WSeg{hw.w3 + (size_t)slice * 256 * 4, 32, hw.n_pad * 4};
                segs[n++] = WSeg{hw.w3 + (size_t)(slice * 256 + 128) * 4, 32, hw.n_pad * 4};
            }
            ws_produce(wsm, segs, n);
        }
        return;
    }
    WsCons ws(wsm);
    for (int m = warp; m < HT; m += NWARP) {
        const int r = row0 + m;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.rows.active(r)) {
            x = ld4(a.x + (size_t)r * 128 + 4 * lane);
            if (slice == 0 && a.trace_head_in) st4(a.trace_head_in + (size_t)r * 128 + 4 * lane, x);
        }
        st4(sx + m * HLD + 4 * lane, x);
    }
    csync();
    tile_gemm<HT>(ws, sx, HLD, 32, [&](int m, int n, float v) { sh[m * HLD + n] = v + __ldg(hw.b0 + n); });
    csync();
    rows_layernorm_c<HT, true>(sh, HLD, hw.ln_g, hw.ln_b);
    csync();
    if (is_state) {                          // agent_decoder.py:2166
        tile_gemm<HT>(ws, sh, HLD, 32, [&](int m, int n, float v) {
            const int r = row0 + m;
            if (n < hw.n_out && a.rows.active(r)) {
                v += __ldg(hw.b3 + n);
                a.state_logits[(size_t)r * 4 + n] = v;
                if (a.trace_state) a.trace_state[(size_t)r * 3 + n] = v;
            }
        });
        return;
    }
    for (int half = 0; half < 2; ++half) {
        tile_gemm<HT>(ws, sh, HLD, 32, [&](int m, int n, float v) {
            const int col = half * 128 + n;
            v += __ldg(hw.b3 + slice * 256 + col);
            slog[m * 256 + col] = v;
            const int r = row0 + m;
            if (a.trace_logits && a.rows.active(r)) a.trace_logits[(size_t)r * hw.n_out + slice * 256 + col] = v;
        });
    }
    csync();
    // per-row top-KTOP of the slice: one warp per row, 8 values per lane
    for (int m = warp; m < HT; m += NWARP) {
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

// Generic MLPLayer (layers.py:206-215) for operator-level parity: grid = (row tiles of HM, 128-column tiles)
struct MlpLayerArgs {
    int n;
    const float *x;
    MlpHeadW w;
    float *out;                // [n][n_out]
    // up to six heads over the same input rows in one launch (blockIdx.z): the heads of the seed query
    MlpHeadW w2, w3, w4, w5, w6;
    float *out2, *out3, *out4, *out5, *out6;
    int single_stride;         // > 0: only rows 0, single_stride, 2 * single_stride, .. are live; a tile that holds one live
                               //      row (its first) takes the GEMV path of the tile products
};
constexpr size_t MLP_LAYER_SMEM = (size_t)(WS_SMEM_FLOATS + 2 * HM * HLD) * sizeof(float);
__global__ void __launch_bounds__(NT_S) k_mlp_layer(const MlpLayerArgs a) {
    extern __shared__ __align__(16) float smem[];
    WsSmem wsm(smem);
    float *sx = smem + WS_SMEM_FLOATS, *sh = sx + HM * HLD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * HM, n0 = blockIdx.y * 128;
    // (selected by value: indexing the parameter block dynamically would move it to a local-memory frame)
    MlpHeadW hw;
    float *out;
    switch (blockIdx.z) {
        case 0: hw = a.w; out = a.out; break;
        case 1: hw = a.w2; out = a.out2; break;
        case 2: hw = a.w3; out = a.out3; break;
        case 3: hw = a.w4; out = a.out4; break;
        case 4: hw = a.w5; out = a.out5; break;
        default: hw = a.w6; out = a.out6; break;
    }
    const bool single = a.single_stride >= HM || (a.single_stride > 0 && row0 + a.single_stride >= a.n);
    if (n0 >= hw.n_pad) return;
    ws_init(wsm);
    if (warp == NWARP) {
        if (lane < WS_STAGES) {
            WSeg segs[2];
            segs[0] = WSeg{hw.w0, hw.k4_in, 512};
            segs[1] = WSeg{hw.w3 + (size_t)n0 * 4, 32, hw.n_pad * 4};
            ws_produce(wsm, segs, 2);
        }
        return;
    }
    WsCons ws(wsm);
    for (int m = warp; m < HM; m += NWARP) {
        const int r = row0 + m;
        st4(sx + m * HLD + 4 * lane, r < a.n ? ld4(a.x + (size_t)r * 128 + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    csync();
    tile_gemm<HM>(ws, sx, HLD, hw.k4_in, [&](int m, int n, float v) { sh[m * HLD + n] = v + __ldg(hw.b0 + n); }, single);
    csync();
    rows_layernorm_c<HM, true>(sh, HLD, hw.ln_g, hw.ln_b);
    csync();
    tile_gemm<HM>(ws, sh, HLD, 32, [&](int m, int n, float v) {
        const int r = row0 + m;
        if (r < a.n && n0 + n < hw.n_out) out[(size_t)r * hw.n_out + n0 + n] = v + __ldg(hw.b3 + n0 + n);
    }, single);
}

}  // namespace infgen
