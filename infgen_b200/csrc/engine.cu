// Host side of libinfgen_b200.so: weight layout registry, engine state, launch sequences, CUDA-graph replay, C ABI.
// See include/infgen_b200.h for the contract of every exported function.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <unordered_map>
#include <algorithm>

#include "../../include/infgen_b200.h"
#include "common.cuh"
#include "ops.cuh"
#include "layer.cuh"
#include "fourier_tc.cuh"
#include "node.cuh"
#include "node_tc.cuh"
#include "decode.cuh"
#include "insert.cuh"
#include "map.cuh"
#include "prep.cuh"

using namespace infgen;

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t _e = (call);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail(INFGEN_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                 \
                        cudaGetErrorString(_e));                                                              \
    } while (0)
#define CKL()                                                                                                 \
    do {                                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                                  \
        if (_e != cudaSuccess)                                                                                \
            return fail(INFGEN_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__,             \
                        cudaGetErrorString(_e));                                                              \
    } while (0)
#define RET(call)                                                                                             \
    do {                                                                                                      \
        int _r = (call);                                                                                      \
        if (_r != 0) return _r;                                                                               \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// packed weight layout (single source of truth; the host packs by name, see infgen_b200/weights.py)
// ---------------------------------------------------------------------------------------------------------------
struct WEntry { std::string name; int64_t numel, offset; };
static std::vector<WEntry> g_layout;
static std::unordered_map<std::string, int> g_index;
static int64_t g_total = 0;

static void w_add(const std::string &name, int64_t numel) {
    WEntry e{name, numel, g_total};
    g_index[name] = (int)g_layout.size();
    g_layout.push_back(e);
    g_total += (numel + 31) / 32 * 32;
}
static int64_t gemm_numel(int k, int n) { return (int64_t)((k + 3) / 4) * n * 4; }
static void w_ln(const std::string &p) { w_add(p + ".g", 128); w_add(p + ".b", 128); }
static void w_attn(const std::string &p, bool has_pos) {
    w_ln(p + ".ln_src"); w_ln(p + ".ln_dst");
    w_add(p + ".w_qs", gemm_numel(128, 256)); w_add(p + ".b_qs", 256);
    w_add(p + ".w_kv", gemm_numel(128, 256)); w_add(p + ".b_kv", 256);
    if (has_pos) {
        w_add(p + ".w_kr", 128 * 128); w_ln(p + ".ln_r");
        w_add(p + ".w_vr", gemm_numel(128, 128)); w_add(p + ".b_vr", 128);
    }
    w_add(p + ".w_g", gemm_numel(256, 128)); w_add(p + ".b_g", 128);
    w_add(p + ".w_out", gemm_numel(128, 128)); w_add(p + ".b_out", 128);
    w_ln(p + ".ln_post"); w_ln(p + ".ln_ffpre");
    w_add(p + ".w_ff1", gemm_numel(128, 512)); w_add(p + ".b_ff1", 512);
    w_add(p + ".w_ff2", gemm_numel(512, 128)); w_add(p + ".b_ff2", 128);
    w_ln(p + ".ln_ffpost");
}
static void w_fourier(const std::string &p, int d) {
    w_add(p + ".freqs", d * 64);
    for (int i = 0; i < d; ++i) {
        std::string q = p + ".mlps." + std::to_string(i);
        w_add(q + ".w0", gemm_numel(129, 128)); w_add(q + ".b0", 128); w_ln(q + ".ln");
        w_add(q + ".w3", gemm_numel(128, 128)); w_add(q + ".b3", 128);
    }
    w_ln(p + ".out_ln"); w_add(p + ".w_out", gemm_numel(128, 128)); w_add(p + ".b_out", 128);
}
static void w_mlp_emb(const std::string &p, int kin) {
    w_add(p + ".w0", gemm_numel(kin, 128)); w_add(p + ".b0", 128); w_ln(p + ".ln1");
    w_add(p + ".w3", gemm_numel(128, 128)); w_add(p + ".b3", 128); w_ln(p + ".ln4");
    w_add(p + ".w6", gemm_numel(128, 128)); w_add(p + ".b6", 128);
}
static int pad128(int n) { return (n + 127) / 128 * 128; }
static void w_head(const std::string &p, int kin, int nout) {
    w_add(p + ".w0", gemm_numel(kin, 128)); w_add(p + ".b0", 128); w_ln(p + ".ln");
    w_add(p + ".w3", gemm_numel(128, pad128(nout))); w_add(p + ".b3", pad128(nout));
}
static const int GRID_SIZE = 1961, ANGLE_SIZE = 120, TOKEN_SIZE = 2048;
static const int MAP_TOKEN_SIZE = 1024, MAP_TOKEN_DIM = 22;       // map_decoder.py:59-63
// rows per scene: the scene's own agents plus everything the insertion stage may append (the reference never compacts:
// up to 10 rows per iteration, agent_decoder.py:1738, i.e. 3,000 over a 150 s rollout)
static const int MAX_ROW_CAPACITY = 8192;
static void build_layout() {
    if (!g_layout.empty()) return;
    w_add("type_a_emb", 4 * 128); w_add("state_a_emb", 4 * 128);
    w_add("no_token_emb", 128); w_add("bos_token_emb", 128); w_add("invalid_offset_token_emb", 128);
    w_mlp_emb("shape_emb", 3);
    w_fourier("x_a_emb", 2); w_fourier("r_t_emb", 4); w_fourier("r_pt2a_emb", 3); w_fourier("r_a2a_emb", 3);
    w_fourier("r_pt2sa_emb", 3); w_fourier("r_a2sa_emb", 3);
    w_mlp_emb("token_emb_veh", 8); w_mlp_emb("token_emb_ped", 8); w_mlp_emb("token_emb_cyc", 8);
    w_mlp_emb("token_emb_grid", 2); w_mlp_emb("fusion_emb", 512);
    const char *stacks6[] = {"t_attn_layers", "pt2a_attn_layers", "a2a_attn_layers"};
    for (auto s : stacks6)
        for (int i = 0; i < 6; ++i) w_attn(std::string(s) + "." + std::to_string(i), true);
    for (int i = 0; i < 3; ++i) w_attn("pt2sa_attn_layers." + std::to_string(i), true);
    for (int i = 0; i < 3; ++i) w_attn("a2sa_attn_layers." + std::to_string(i), true);
    for (int i = 0; i < 3; ++i) w_attn("occ2sa_attn_layers." + std::to_string(i), false);
    w_head("token_predict_head", 128, TOKEN_SIZE); w_head("state_predict_head", 128, 3);
    w_head("seed_state_predict_head", 128, 2); w_head("seed_type_predict_head", 128, 3);
    w_head("seed_shape_predict_head", 128, 3); w_head("seed_pos_rel_token_predict_head", 128, GRID_SIZE);
    w_head("seed_offset_xy_predict_head", 128, 2); w_head("seed_agent_occ_embed", GRID_SIZE, 128);
    w_head("seed_heading_rel_token_predict_head", 128, ANGLE_SIZE);
    w_head("grid_agent_occ_head", 128, GRID_SIZE); w_head("grid_pt_occ_head", 128, GRID_SIZE);
    // map encoder `InfGenMapDecoder` (map_decoder.py:46-64); names carry the prefix "map."
    w_add("map.type_pt_emb", 17 * 128); w_add("map.polygon_type_emb", 4 * 128); w_add("map.light_pl_emb", 4 * 128);
    w_fourier("map.r_pt2pt_emb", 3);
    for (int i = 0; i < 3; ++i) w_attn("map.pt2pt_layers." + std::to_string(i), true);
    w_head("map.token_predict_head", 128, MAP_TOKEN_SIZE);
    w_mlp_emb("map.token_emb", MAP_TOKEN_DIM);
}

// ---------------------------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct ProfRec { int cls; cudaEvent_t a, b; };

struct infgen_engine {
    infgen_config cfg;
    bool profile = false;
    std::vector<ProfRec> prof;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaStream_t side_stream = nullptr;                 // edges of the next column, concurrent with its embedding
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool ins_ride = false;                              // seed queries carry the last appended row (one scene per tile)
    cudaStream_t side_stream2 = nullptr;                // second concurrent branch (occupancy node of the next insertion pass)
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
    cudaStream_t side_stream3 = nullptr;                // third (heading-stack K|V rows of the row appended one pass earlier)
    cudaEvent_t ev_fork3 = nullptr, ev_join3 = nullptr;
    bool early_edges = false;                           // motion-only engines build the next column's edges early
    float *blob = nullptr;
    float *cs_blob = nullptr;                           // cluster-sliced AttentionLayer chunks (layer.cuh)
    std::unordered_map<std::string, FourierW> fourier_cache;
    std::vector<float *> wimgs;                         // FourierEmbedding tensor-core weight images (fourier_tc.cuh)
    std::unordered_map<std::string, const float *> npk;   // layer prefix -> node-packed copy
    float *np_blob = nullptr;                           // node-packed motion + map layers (node.cuh), [21][np::FLOATS]
    float *vrf_blob = nullptr;                          // their folded to_v_r tables, [21][vrf::FLOATS]
    std::unordered_map<std::string, const float *> vrfs;
    std::unordered_map<std::string, const float *> tcimgs;   // layer prefix -> tensor-core weight image (node_tc.cuh)
    int attn_ctas = 296;                                // persistent k_attn CTAs: 2 per SM
    bool node_mma = false;                              // k_node GEMMs on mma.sync 3xTF32 (INFGEN_NODE_GEMM=mma)
    bool node_tc = true;                                // row-tile path on k_node_tc (tcgen05); INFGEN_NODE_GEMM=ffma|mma: k_node
    void *prep_buf = nullptr; size_t prep_bytes = 0;    // arena of infgen_prepare_scene (row f2)
    float *tc_blob = nullptr;                           // tensor-core weight images of the node-packed layers, [30][ntc::IMG_FLOATS]
    int layer_path = 0;                                 // 0 auto, 1 cluster kernels only, 2 row-tile (k_attn + k_node) only
    bool fourier_tc = true;                             // INFGEN_FOURIER=ffma selects the FFMA row-tile kernel instead
    std::unordered_map<std::string, std::pair<const float *, const float *>> cs;   // layer -> (post, pre) chunks
    float *grid_cells = nullptr, *vocab = nullptr;
    float *tok_tab = nullptr, *grid_tab = nullptr;      // [3][V+2][128], [G+1][128]
    AttnW t[6], m[6], a[6];
    AttnW occ2sa[3], pt2sa[3], a2sa[3];                 // insertion stage (agent_decoder.py:235-247)
    // map encoder (map_decoder.py:70-130)
    AttnW mp[3];
    FourierW f_pp;                                      // r_pt2pt_emb
    MlpEmbW e_map_tok;
    MlpHeadW h_map_tok;
    float *map_tok_tab = nullptr;                       // [n_tokens][128] token_emb(traj_src)
    int map_n_tokens = 0, map_P = 0;                    // vocabulary size; tokens of the last infgen_map_encode
    FourierW f_ps, f_as;                                // r_pt2sa_emb, r_a2sa_emb
    MlpHeadW h_seed_state, h_seed_type, h_seed_shape, h_seed_pos, h_seed_heading, h_seed_offset, h_occ_embed,
        h_ag_occ, h_pt_occ;
    float *seed_feat = nullptr;                         // [128] input feature of the seed query row
    InsState ins;
    bool ins_ready = false;
    FourierW f_t, f_m, f_a, f_x;
    FourierW f_t3;                                      // r_t_emb restricted to its first three input dims (tensor-core path)
    float *t_dim_table = nullptr;                       // [window][128] per-dim MLP output of the 4th temporal input (-1 .. -window)
    MlpEmbW e_shape, e_fusion, e_tok[3], e_grid;
    MlpHeadW h_tok, h_state;
    const float *type_emb = nullptr, *state_emb = nullptr;
    // scene batch
    bool loaded = false;
    int n_scenes = 0, cap = 0, T = 0, S = 0, R = 0, P = 0, n_rows_sum = 0, max_rows = 0;
    std::vector<int> n_rows0;                           // rows of every scene at load time (appended rows follow them)
    int iters_done = 0, prefilled = 0;
    std::unordered_map<std::string, DevBuf> bufs;       // named device buffers (debug-readable)
    DecState st;
    int row_tile = 4;
    int *d_err = nullptr;
    // forcing
    bool forcing = false;
    bool forcing_no_insert = false;
    // graphs: [0] one motion iteration, [1] insertion stage (WHILE / IF conditional nodes) + motion iteration
    cudaGraph_t graph[2] = {nullptr, nullptr};
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    size_t graph_nodes[2] = {0, 0};
    size_t pass_nodes = 0, heading_nodes = 0;           // kernel nodes of the WHILE / IF bodies of graph [1]
    int64_t launches = 0;
    int64_t ins_replays = 0;                            // replays of graph [1] since the counters were last folded in
    bool capturing = false;
};

static const float *W(infgen_engine *e, const std::string &name) {
    auto it = g_index.find(name);
    if (it == g_index.end()) {
        fprintf(stderr, "infgen_b200: unknown packed weight '%s'\n", name.c_str());
        abort();
    }
    return e->blob + g_layout[it->second].offset;
}
static AttnW make_attn(infgen_engine *e, const std::string &p, bool has_pos) {
    AttnW w;
    memset(&w, 0, sizeof(w));
    w.ln_src_g = W(e, p + ".ln_src.g"); w.ln_src_b = W(e, p + ".ln_src.b");
    w.ln_dst_g = W(e, p + ".ln_dst.g"); w.ln_dst_b = W(e, p + ".ln_dst.b");
    w.w_qs = W(e, p + ".w_qs"); w.b_qs = W(e, p + ".b_qs"); w.w_kv = W(e, p + ".w_kv"); w.b_kv = W(e, p + ".b_kv");
    if (has_pos) {
        w.w_kr = W(e, p + ".w_kr"); w.ln_r_g = W(e, p + ".ln_r.g"); w.ln_r_b = W(e, p + ".ln_r.b");
        w.w_vr = W(e, p + ".w_vr"); w.b_vr = W(e, p + ".b_vr");
    }
    w.w_g = W(e, p + ".w_g"); w.b_g = W(e, p + ".b_g"); w.w_out = W(e, p + ".w_out"); w.b_out = W(e, p + ".b_out");
    w.ln_post_g = W(e, p + ".ln_post.g"); w.ln_post_b = W(e, p + ".ln_post.b");
    w.ln_ffpre_g = W(e, p + ".ln_ffpre.g"); w.ln_ffpre_b = W(e, p + ".ln_ffpre.b");
    w.w_ff1 = W(e, p + ".w_ff1"); w.b_ff1 = W(e, p + ".b_ff1"); w.w_ff2 = W(e, p + ".w_ff2"); w.b_ff2 = W(e, p + ".b_ff2");
    w.ln_ffpost_g = W(e, p + ".ln_ffpost.g"); w.ln_ffpost_b = W(e, p + ".ln_ffpost.b");
    w.has_pos = has_pos ? 1 : 0;
    auto it = e->cs.find(p);
    if (it != e->cs.end()) { w.cs_post = it->second.first; w.cs_pre = it->second.second; }
    auto nt = e->npk.find(p);
    if (nt != e->npk.end()) w.npk = nt->second;
    auto vt = e->vrfs.find(p);
    if (vt != e->vrfs.end()) w.vrf = vt->second;
    auto tt = e->tcimgs.find(p);
    if (tt != e->tcimgs.end()) w.tcimg = tt->second;
    return w;
}

// Cluster-sliced copies of every AttentionLayer (layer.cuh): for CTA c of a cluster the column slices of all Linears
// of the layer, contiguous, so one bulk copy brings them into shared memory.  Built on the host from the packed blob.
static const char *ATTN_STACKS[] = {"t_attn_layers", "pt2a_attn_layers", "a2a_attn_layers", "pt2sa_attn_layers",
                                    "a2sa_attn_layers", "occ2sa_attn_layers"};
static const int ATTN_STACK_LAYERS[] = {6, 6, 6, 3, 3, 3};
static void slice_cols(float *dst, int nl, int n_off, const float *src, int k4n, int ldn, int n0, int cnt) {
    for (int k4 = 0; k4 < k4n; ++k4)
        for (int n = 0; n < cnt; ++n)
            memcpy(dst + ((size_t)k4 * nl + n_off + n) * 4, src + ((size_t)k4 * ldn + n0 + n) * 4, 4 * sizeof(float));
}
static int build_cluster_weights(infgen_engine *e, const float *hw) {
    auto H = [&](const std::string &name) -> const float * {
        auto it = g_index.find(name);
        return it == g_index.end() ? nullptr : hw + g_layout[it->second].offset;
    };
    int n_layers = 0;
    for (int s = 0; s < 6; ++s) n_layers += ATTN_STACK_LAYERS[s];
    const size_t per_layer = (size_t)CL * (cs_post::FLOATS + cs_pre::FLOATS);
    std::vector<float> host(per_layer * n_layers, 0.f);
    CK(cudaMalloc(&e->cs_blob, host.size() * sizeof(float)));
    size_t li = 0;
    for (int s = 0; s < 6; ++s)
        for (int i = 0; i < ATTN_STACK_LAYERS[s]; ++i, ++li) {
            const std::string p = std::string(ATTN_STACKS[s]) + "." + std::to_string(i);
            float *post = host.data() + li * per_layer, *pre = post + (size_t)CL * cs_post::FLOATS;
            const bool has_pos = H(p + ".w_kr") != nullptr;
            for (int c = 0; c < CL; ++c) {
                float *d = post + (size_t)c * cs_post::FLOATS;
                if (has_pos) {
                    slice_cols(d + cs_post::WVR, 16, 0, H(p + ".w_vr"), 32, 128, 16 * c, 16);
                    memcpy(d + cs_post::BVR, H(p + ".b_vr") + 16 * c, 16 * sizeof(float));
                    memcpy(d + cs_post::LN_R_G, H(p + ".ln_r.g"), 128 * sizeof(float));
                    memcpy(d + cs_post::LN_R_B, H(p + ".ln_r.b"), 128 * sizeof(float));
                }
                slice_cols(d + cs_post::WG, 16, 0, H(p + ".w_g"), 64, 128, 16 * c, 16);
                slice_cols(d + cs_post::WO, 16, 0, H(p + ".w_out"), 32, 128, 16 * c, 16);
                slice_cols(d + cs_post::W1, 64, 0, H(p + ".w_ff1"), 32, 512, 64 * c, 64);
                slice_cols(d + cs_post::W2, 16, 0, H(p + ".w_ff2"), 128, 128, 16 * c, 16);
                memcpy(d + cs_post::BG, H(p + ".b_g") + 16 * c, 16 * sizeof(float));
                memcpy(d + cs_post::BO, H(p + ".b_out") + 16 * c, 16 * sizeof(float));
                memcpy(d + cs_post::B1, H(p + ".b_ff1") + 64 * c, 64 * sizeof(float));
                memcpy(d + cs_post::B2, H(p + ".b_ff2") + 16 * c, 16 * sizeof(float));
                memcpy(d + cs_post::LN_DST_G, H(p + ".ln_dst.g"), 128 * sizeof(float));
                memcpy(d + cs_post::LN_DST_B, H(p + ".ln_dst.b"), 128 * sizeof(float));
                memcpy(d + cs_post::LN_POST_G, H(p + ".ln_post.g"), 128 * sizeof(float));
                memcpy(d + cs_post::LN_POST_B, H(p + ".ln_post.b"), 128 * sizeof(float));
                memcpy(d + cs_post::LN_FFPRE_G, H(p + ".ln_ffpre.g"), 128 * sizeof(float));
                memcpy(d + cs_post::LN_FFPRE_B, H(p + ".ln_ffpre.b"), 128 * sizeof(float));
                memcpy(d + cs_post::LN_FFPOST_G, H(p + ".ln_ffpost.g"), 128 * sizeof(float));
                memcpy(d + cs_post::LN_FFPOST_B, H(p + ".ln_ffpost.b"), 128 * sizeof(float));
                float *q = pre + (size_t)c * cs_pre::FLOATS;
                slice_cols(q + cs_pre::WQS, 32, 0, H(p + ".w_qs"), 32, 256, 16 * c, 16);
                slice_cols(q + cs_pre::WQS, 32, 16, H(p + ".w_qs"), 32, 256, 128 + 16 * c, 16);
                slice_cols(q + cs_pre::WKV, 32, 0, H(p + ".w_kv"), 32, 256, 16 * c, 16);
                slice_cols(q + cs_pre::WKV, 32, 16, H(p + ".w_kv"), 32, 256, 128 + 16 * c, 16);
                memcpy(q + cs_pre::BQS, H(p + ".b_qs") + 16 * c, 16 * sizeof(float));
                memcpy(q + cs_pre::BQS + 16, H(p + ".b_qs") + 128 + 16 * c, 16 * sizeof(float));
                memcpy(q + cs_pre::BKV, H(p + ".b_kv") + 16 * c, 16 * sizeof(float));
                memcpy(q + cs_pre::BKV + 16, H(p + ".b_kv") + 128 + 16 * c, 16 * sizeof(float));
                if (has_pos) {
                    memcpy(q + cs_pre::WKR, H(p + ".w_kr") + (size_t)16 * c * 128, 16 * 128 * sizeof(float));
                    memcpy(q + cs_pre::LN_R_G, H(p + ".ln_r.g"), 128 * sizeof(float));
                }
                memcpy(q + cs_pre::LN_DST_G, H(p + ".ln_dst.g"), 128 * sizeof(float));
                memcpy(q + cs_pre::LN_DST_B, H(p + ".ln_dst.b"), 128 * sizeof(float));
            }
            e->cs[p] = {e->cs_blob + li * per_layer, e->cs_blob + li * per_layer + (size_t)CL * cs_post::FLOATS};
        }
    CK(cudaMemcpy(e->cs_blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}
static FourierW make_fourier(infgen_engine *e, const std::string &p, int d) {
    const std::string key = p + "#" + std::to_string(d);
    auto hit = e->fourier_cache.find(key);
    if (hit != e->fourier_cache.end()) return hit->second;
    FourierW w;
    memset(&w, 0, sizeof(w));
    w.freqs = W(e, p + ".freqs");
    for (int i = 0; i < d; ++i) {
        std::string q = p + ".mlps." + std::to_string(i);
        w.w0[i] = W(e, q + ".w0"); w.b0[i] = W(e, q + ".b0"); w.ln_g[i] = W(e, q + ".ln.g"); w.ln_b[i] = W(e, q + ".ln.b");
        w.w3[i] = W(e, q + ".w3"); w.b3[i] = W(e, q + ".b3");
    }
    w.out_ln_g = W(e, p + ".out_ln.g"); w.out_ln_b = W(e, p + ".out_ln.b");
    w.w_out = W(e, p + ".w_out"); w.b_out = W(e, p + ".b_out");
    // tensor-core image: the (2 d + 1) packed matrices split into hi / lo TF32 chunks, in the GEMM order of k_fourier_tc
    float *img = nullptr;
    if (cudaMalloc(&img, ftc::wimg_floats(d) * sizeof(float)) != cudaSuccess) {
        fprintf(stderr, "infgen_b200: cudaMalloc of a FourierEmbedding weight image failed\n");
        abort();
    }
    e->wimgs.push_back(img);
    int jt[9], jd[9];
    const int nj = ftc::job_list(d, jt, jd);
    for (int j = 0; j < nj; ++j) {
        const float *src = jt[j] == 0 ? w.w0[jd[j]] : (jt[j] == 1 ? w.w3[jd[j]] : w.w_out);
        k_wimg_split<<<64, 256, 0, e->stream>>>(src, img + (size_t)j * 4 * ftc::CHUNK, jt[j] == 0 ? 1 : 0);
    }
    for (int i = 0; i < d; ++i) k_wimg_xrow<<<1, 128, 0, e->stream>>>(w.w0[i], img + (size_t)nj * 4 * ftc::CHUNK + (size_t)i * 128);
    w.wimg = img;
    e->fourier_cache[key] = w;
    return w;
}
static MlpEmbW make_mlp_emb(infgen_engine *e, const std::string &p) {
    MlpEmbW w;
    w.w0 = W(e, p + ".w0"); w.b0 = W(e, p + ".b0"); w.ln1_g = W(e, p + ".ln1.g"); w.ln1_b = W(e, p + ".ln1.b");
    w.w3 = W(e, p + ".w3"); w.b3 = W(e, p + ".b3"); w.ln4_g = W(e, p + ".ln4.g"); w.ln4_b = W(e, p + ".ln4.b");
    w.w6 = W(e, p + ".w6"); w.b6 = W(e, p + ".b6");
    return w;
}
static MlpHeadW make_head(infgen_engine *e, const std::string &p, int kin, int nout) {
    MlpHeadW w;
    w.w0 = W(e, p + ".w0"); w.b0 = W(e, p + ".b0"); w.ln_g = W(e, p + ".ln.g"); w.ln_b = W(e, p + ".ln.b");
    w.w3 = W(e, p + ".w3"); w.b3 = W(e, p + ".b3");
    w.k4_in = (kin + 3) / 4; w.n_out = nout; w.n_pad = pad128(nout);
    return w;
}

static void drop_graph(infgen_engine *e) {
    for (int i = 0; i < 2; ++i) {
        if (e->graph_exec[i]) { cudaGraphExecDestroy(e->graph_exec[i]); e->graph_exec[i] = nullptr; }
        if (e->graph[i]) { cudaGraphDestroy(e->graph[i]); e->graph[i] = nullptr; }
    }
}
static int ensure(infgen_engine *e, const char *name, size_t bytes, void **out) {
    DevBuf &b = e->bufs[name];
    if (b.bytes < bytes) {
        if (b.p) CK(cudaFree(b.p));
        b.p = nullptr; b.bytes = 0;
        size_t alloc = (bytes + 255) / 256 * 256;
        CK(cudaMalloc(&b.p, alloc));
        CK(cudaMemsetAsync(b.p, 0, alloc, e->stream));
        b.bytes = alloc;
        drop_graph(e);
    }
    *out = b.p;
    return 0;
}
template <typename Tp>
static int ensure_t(infgen_engine *e, const char *name, size_t count, Tp **out) {
    void *p = nullptr;
    RET(ensure(e, name, count * sizeof(Tp), &p));
    *out = (Tp *)p;
    return 0;
}
static float *fbuf(infgen_engine *e, const char *name) { return (float *)e->bufs[name].p; }

static inline void count_launch(infgen_engine *e) { if (!e->capturing) e->launches++; }

// per-kernel-class device timing (bench.py roofline leg): event pairs around every launch, plain launches only
enum KClass { KC_EDGE_BUILD, KC_FOURIER, KC_EMBED, KC_LAYER_STACK, KC_LAYER_TM, KC_LAYER_A, KC_HEADS, KC_ADVANCE, KC_INSERT,
              KC_MISC, KC_ATTN, KC_NODE, KC_INS_LAYER_AGENTS, KC_INS_LAYER_QUERY, KC_INS_LAYER_NEW, KC_INS_FOURIER,
              KC_INS_HEADS, KC_COUNT };
static const char *KCLASS_NAME[KC_COUNT] = {"k_edge_build", "k_fourier:edges", "k_embed_column", "k_layer:stack18",
                                            "k_layer:temporal+map", "k_layer:agent", "k_heads", "k_advance",
                                            "insertion:small kernels", "misc", "k_attn", "k_node",
                                            "insertion:k_layer agents (edge-less K|V)", "insertion:k_layer seed query",
                                            "insertion:k_layer new rows", "insertion:k_fourier", "insertion:k_mlp_layer heads"};
struct ProfScope {
    infgen_engine *e;
    bool on;
    ProfScope(infgen_engine *e_, int cls) : e(e_), on(e_->profile && !e_->capturing) {
        if (!on) return;
        ProfRec r;
        r.cls = cls;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, e->stream);
        e->prof.push_back(r);
    }
    ~ProfScope() { if (on) cudaEventRecord(e->prof.back().b, e->stream); }
};

// ---------------------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------------------
// B200: at most 15 clusters of 8 CTAs with ~211 KB of shared memory are co-resident (cudaOccupancyMaxActiveClusters,
// tools/probe/cluster_occ.cu); a 16th cluster costs a whole second wave
static const int MAX_CLUSTERS = 15;
// tiles of M rows a kernel launches for a row space (one per scene for per-scene row spaces)
static int row_tiles(const RowSpace &r, int M) { return r.list ? (r.list_cap + M - 1) / M : (r.n_total + M - 1) / M; }
static int launch_layer(infgen_engine *e, const LayerArgs &a_in, int cls) {
    LayerArgs a = a_in;
    if (e->bufs.count("tstamp") && e->bufs["tstamp"].p) {
        // debug stamps: INFGEN_TSTAMP_CLS=<class index> records that class alone (default: the motion stack launches)
        static const int want = getenv("INFGEN_TSTAMP_CLS") ? atoi(getenv("INFGEN_TSTAMP_CLS")) : -1;
        if (want < 0 ? (cls == KC_LAYER_STACK || cls == KC_LAYER_TM || cls == KC_LAYER_A) : cls == want)
            a.tstamp = (long long *)e->bufs["tstamp"].p + (cls == KC_LAYER_A ? 256 : 0);
    }
    ProfScope ps(e, cls);
    const int M = e->row_tile;
    const int clusters = row_tiles(a.rows, M);
    if (clusters == 0) return 0;
    if (M == 8) k_layer<8><<<clusters * CL, NT, LayerSmem<8>::BYTES, e->stream>>>(a);
    else k_layer<4><<<clusters * CL, NT, LayerSmem<4>::BYTES, e->stream>>>(a);
    CKL();
    count_launch(e);
    return 0;
}
static PreArgs make_pre(const AttnW &w, bool pre_kv, float *kv_out, bool kv_ring, int col_add, bool to_global) {
    PreArgs p;
    memset(&p, 0, sizeof(p));
    p.w = w.cs_pre; p.pre_kv = pre_kv ? 1 : 0; p.kv_out = kv_out; p.kv_ring = kv_ring ? 1 : 0; p.col_add = col_add;
    p.to_global = to_global ? 1 : 0;
    return p;
}
// watchdog record of k_fourier_tc (fourier_tc.cuh:ftc_wait): non-zero = an mbarrier wait timed out
static int check_ftc_watchdog() {
    int h[8] = {0};
    CK(cudaMemcpyFromSymbol(h, g_ftc_hang, sizeof(h)));
    if (h[0]) {
        int z[8] = {0};
        cudaMemcpyToSymbol(g_ftc_hang, z, sizeof(z));
        return fail(INFGEN_ERR_CUDA, "k_fourier_tc: %d mbarrier waits timed out (first code*1000+thread per class: weights %d, "
                    "mma/A %d, mma/B %d, A stage %d, accumulator %d)", h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    CK(cudaMemcpyFromSymbol(h, g_ntc_hang, sizeof(h)));
    if (h[0]) {
        int z[8] = {0};
        cudaMemcpyToSymbol(g_ntc_hang, z, sizeof(z));
        return fail(INFGEN_ERR_CUDA, "k_node_tc: %d mbarrier waits timed out (first code*1000+thread per class: weights %d, "
                    "mma/A %d, mma/B %d, A stage %d, accumulator %d, accumulator free %d)", h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
    return 0;
}
// up to three FourierEmbeddings in one launch.  Embeddings without a categorical seed run on the tensor cores
// (k_fourier_tc, tiles of 128 slots); the FFMA row-tile kernel (tiles of 16) serves the rest and INFGEN_FOURIER=ffma.
static int launch_fourier(infgen_engine *e, const FourierArgs *jobs, int n_jobs, int cls = KC_MISC) {
    FourierBatch fb;
    memset(&fb, 0, sizeof(fb));
    bool tc = e->fourier_tc;
    for (int j = 0; j < n_jobs; ++j) {
        if (jobs[j].dim < 1 || jobs[j].dim > 4) return fail(INFGEN_ERR_INVALID_ARG, "FourierEmbedding input_dim %d unsupported", jobs[j].dim);
        if (jobs[j].cat_tab || !jobs[j].w.wimg || jobs[j].ffma) tc = false;
    }
    const int tm = tc ? ftc::TM : FM;
    int tiles = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const int t = (jobs[j].n_slots + tm - 1) / tm;
        if (t == 0) continue;
        fb.job[fb.n_jobs] = jobs[j];
        fb.tile0[fb.n_jobs] = tiles;
        tiles += t;
        fb.n_jobs++;
    }
    fb.tile0[fb.n_jobs] = tiles;
    if (tiles == 0) return 0;
    ProfScope ps(e, cls);
    if (tc) k_fourier_tc<<<tiles, ftc::THREADS, ftc::SMEM, e->stream>>>(fb);
    else k_fourier<<<tiles, NT_S, FOURIER_SMEM, e->stream>>>(fb);
    CKL();
    count_launch(e);
    return 0;
}
static int launch_mlp_embed(infgen_engine *e, const MlpEmbArgs &a, int cls = KC_MISC) {
    int grid = row_tiles(a.rows, EM);
    if (grid == 0) return 0;
    ProfScope ps(e, cls);
    k_mlp_embed<<<grid, NT_S, mlp_embed_smem(a.k4), e->stream>>>(a);
    CKL();
    count_launch(e);
    return 0;
}

static RowSpace scene_rows(infgen_engine *e) {
    RowSpace r;
    memset(&r, 0, sizeof(r));
    r.n_total = e->R; r.cap = e->cap; r.n_rows = e->st.n_rows;
    return r;
}
static RowSpace flat_rows(int n) {
    RowSpace r;
    memset(&r, 0, sizeof(r));
    r.n_total = n;
    return r;
}
// the rows the last insertion pass appended (at most one per scene), as a compact row list written by k_seed_decide
static RowSpace new_rows(infgen_engine *e) {
    RowSpace r = scene_rows(e);
    r.list = e->ins.new_list; r.n_list = e->ins.n_new_list; r.list_cap = e->n_scenes;
    return r;
}
// the rows the pass before the last one appended (their heading-stack K|V rows are due when the next row arrives)
static RowSpace prev_rows(infgen_engine *e) {
    RowSpace r = scene_rows(e);
    r.list = e->ins.prev_list; r.n_list = e->ins.n_prev_list; r.list_cap = e->n_scenes;
    return r;
}

// column embedding (agent_decoder.py:2265-2287) of column col+col_add -> x (temporal layer 0 projects it in k_layer)
static int enqueue_embed_column(infgen_engine *e, int col_add) {
    DecState &s = e->st;
    const int R = e->R;
    ColEmbArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.rows = scene_rows(e); ca.fx = e->f_x; ca.fusion = e->e_fusion;
    ca.s = s; ca.col_add = col_add; ca.cat_tab = fbuf(e, "cat_tab");
    ca.tok_tab = e->tok_tab; ca.state_tab = e->state_emb; ca.grid_tab = e->grid_tab; ca.out = fbuf(e, "x");
    {
        ProfScope ps(e, KC_EMBED);
        k_embed_column<<<(R + EM - 1) / EM, NT_S, COLEMB_SMEM, e->stream>>>(ca);
    }
    CKL(); count_launch(e);
    return 0;
}

static int enqueue_embed_rows(infgen_engine *e, const int *row_lo, float *out2 = nullptr, float *out3 = nullptr) {
    DecState &s = e->st;
    ColEmbArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.rows = new_rows(e); ca.fx = e->f_x; ca.fusion = e->e_fusion;
    ca.s = s; ca.col_add = 0; ca.cat_tab = fbuf(e, "cat_tab");
    ca.tok_tab = e->tok_tab; ca.state_tab = e->state_emb; ca.grid_tab = e->grid_tab; ca.out = fbuf(e, "x");
    ca.out2 = out2; ca.out3 = out3;
    {
        ProfScope ps(e, KC_INSERT);
        k_embed_column<<<row_tiles(ca.rows, EM), NT_S, COLEMB_SMEM, e->stream>>>(ca);
    }
    CKL(); count_launch(e);
    return 0;
}

// the 18-layer stack for the current column.  When every cluster of the launch is co-resident (<= MAX_CLUSTERS, i.e. up
// to 120 rows) ONE launch runs all 18 layers with a grid barrier before each agent<->agent attention (its K/V rows come
// from every row of the scene); otherwise one launch per {temporal + map} and per {agent} layer, whose K/V exchange is
// the launch boundary.  with_edges=false: history columns that receive no edges (prefill).
// node-packed copies of the 18 motion layers (node.cuh): every Linear cut into contiguous 128-column blocks
static int build_node_weights(infgen_engine *e) {
    CK(cudaMalloc(&e->np_blob, (size_t)30 * np::FLOATS * sizeof(float)));
    CK(cudaMalloc(&e->vrf_blob, (size_t)21 * vrf::FLOATS * sizeof(float)));
    CK(cudaMalloc(&e->tc_blob, (size_t)30 * ntc::IMG_FLOATS * sizeof(float)));
    float *kr_tmp = nullptr;                             // to_k_r.weight re-packed [32 k4][128][4]
    CK(cudaMalloc(&kr_tmp, 16384 * sizeof(float)));
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e->attn_ctas = 2 * sms;
    }
    // motion stacks and the map encoder (k_attn + k_node), then the seed stacks of the insertion stage (k_node only: the
    // agents' edge-less pass of large batches)
    AttnW *stacks[7] = {e->t, e->m, e->a, e->mp, e->occ2sa, e->pt2sa, e->a2sa};
    for (int s = 0; s < 7; ++s)
        for (int i = 0; i < (s < 3 ? 6 : 3); ++i) {
            AttnW &w = stacks[s][i];
            float *d = e->np_blob + (size_t)(s < 4 ? s * 6 + i : 21 + (s - 4) * 3 + i) * np::FLOATS;
            auto pack = [&](const float *src, int off, int K4, int N) {
                k_node_pack<<<(K4 * N + 255) / 256, 256, 0, e->stream>>>(src, d + off, K4, N);
            };
            if (w.has_pos) pack(w.w_vr, np::VR, 32, 128);
            pack(w.w_g, np::G, 64, 128); pack(w.w_out, np::OUT, 32, 128);
            pack(w.w_ff1, np::FF1, 32, 512); pack(w.w_ff2, np::FF2, 128, 128);
            pack(w.w_qs, np::QS, 32, 256); pack(w.w_kv, np::KV, 32, 256);
            CKL();
            w.npk = d;
            {
                // tensor-core image (node_tc.cuh): the sixteen 128 x 128 blocks of the layer as hi / lo TF32 chunks
                float *img = e->tc_blob + (size_t)(s < 4 ? s * 6 + i : 21 + (s - 4) * 3 + i) * ntc::IMG_FLOATS;
                auto split = [&](const float *blk, int b) {
                    k_wimg_split<<<64, 256, 0, e->stream>>>(blk, img + (size_t)b * 4 * ntc::CHUNK, 0);
                };
                split(d + np::G, ntc::B_G0); split(d + np::G + 16384, ntc::B_G1); split(d + np::OUT, ntc::B_OUT);
                for (int j = 0; j < 4; ++j) {
                    split(d + np::FF1 + j * 16384, ntc::B_UP0 + 2 * j);
                    split(d + np::FF2 + j * 16384, ntc::B_DN0 + 2 * j);
                }
                split(d + np::QS, ntc::B_Q); split(d + np::QS + 16384, ntc::B_S);
                split(d + np::KV, ntc::B_K); split(d + np::KV + 16384, ntc::B_V);
                if (w.has_pos) {
                    k_pack_kn<<<64, 256, 0, e->stream>>>(w.w_kr, kr_tmp);
                    split(kr_tmp, ntc::B_KR);
                }
                CKL();
                w.tcimg = img;
            }
            static const char *names[7] = {"t_attn_layers.", "pt2a_attn_layers.", "a2a_attn_layers.", "map.pt2pt_layers.",
                                           "occ2sa_attn_layers.", "pt2sa_attn_layers.", "a2sa_attn_layers."};
            e->npk[std::string(names[s]) + std::to_string(i)] = d;
            e->tcimgs[std::string(names[s]) + std::to_string(i)] = w.tcimg;
            if (s < 4) {
                float *vf = e->vrf_blob + (size_t)(s * 6 + i) * vrf::FLOATS;
                k_vr_fold_pack<<<1, 128, 0, e->stream>>>(w.w_vr, w.ln_r_g, w.ln_r_b, w.b_vr, vf);
                CKL();
                w.vrf = vf;
                e->vrfs[std::string(names[s]) + std::to_string(i)] = vf;
            }
        }
    CK(cudaStreamSynchronize(e->stream));
    cudaFree(kr_tmp);
    return 0;
}
struct NodeBufs {              // hand-over buffers of the row-tile path
    float *x, *q, *s, *qr, *agg;
};
static NodeBufs scene_node_bufs(infgen_engine *e) {
    return NodeBufs{fbuf(e, "x"), fbuf(e, "q"), fbuf(e, "s"), fbuf(e, "qr"), fbuf(e, "agg")};
}
// edge attention of layer `lw` (its folded to_v_r table) over `rows`
static int launch_attn(infgen_engine *e, const RowSpace &rows, const SubArgs &sub, const AttnW &lw, const NodeBufs *nb = nullptr) {
    const NodeBufs b = nb ? *nb : scene_node_bufs(e);
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = rows; a.sub = sub; a.q = b.q; a.qr = b.qr;
    a.agg = b.agg; a.vrf = sub.has_pos ? lw.vrf : nullptr;
    if (sub.has_pos && !lw.vrf) return fail(INFGEN_ERR_STATE, "layer has no folded to_v_r table for the row-tile path");
    ProfScope ps(e, KC_ATTN);
    // (the kernel deals the ACTIVE rows over its warps; rows.n_total only bounds the useful grid)
    const int upper = rows.cap ? std::min(rows.n_total, e->n_rows_sum + 10 * e->n_scenes * std::max(e->S, 1)) : rows.n_total;
    const int ctas = std::max(1, std::min(e->attn_ctas, (upper + ATTN_WARPS - 1) / ATTN_WARPS));
    k_attn<<<ctas, ATTN_WARPS * 32, ATTN_SMEM, e->stream>>>(a);
    CKL(); count_launch(e);
    return 0;
}
// finish layer `lw` (NULL: nothing to finish) and project the inputs of layer `pw` (NULL: none)
static int launch_node(infgen_engine *e, const RowSpace &rows, const AttnW *lw, const AttnW *pw, bool pre_kv, float *kv_out,
                       bool kv_ring, float *trace_out, const NodeBufs *nb = nullptr, bool edgeless = false, int cls = KC_NODE) {
    const NodeBufs b = nb ? *nb : scene_node_bufs(e);
    NodeArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = rows; a.x = b.x; a.agg = b.agg;
    a.q = b.q; a.s = b.s; a.qr = b.qr;
    if (lw) { a.w_post = lw->npk; a.lw = *lw; }
    if (pw) { a.w_pre = pw->npk; a.pw = *pw; }
    a.pre_kv = pre_kv ? 1 : 0; a.kv_out = kv_out; a.kv_ring = kv_ring ? 1 : 0; a.col_add = 0; a.ring = RING;
    a.col_ptr = e->st.col; a.trace_out = trace_out; a.edgeless = edgeless ? 1 : 0;
    ProfScope ps(e, cls);
    if (e->node_tc && (!lw || lw->tcimg) && (!pw || pw->tcimg)) {
        // tensor-core kernel: one CTA per 128 ACTIVE rows (the kernel compacts a capacity row space on the fly)
        if (lw) a.tc_post = lw->tcimg;
        if (pw) a.tc_pre = pw->tcimg;
        const int upper = rows.cap ? std::min(rows.n_total, e->n_rows_sum + 10 * e->n_scenes * std::max(e->S, 1)) : rows.n_total;
        k_node_tc<<<std::max(1, (upper + ntc::TM - 1) / ntc::TM), ntc::THREADS, ntc::SMEM, e->stream>>>(a);
    } else if (e->node_mma) k_node<true><<<(rows.n_total + NM - 1) / NM, NT_S, NodeSmem::BYTES, e->stream>>>(a);
    else k_node<false><<<(rows.n_total + NM - 1) / NM, NT_S, NodeSmem::BYTES, e->stream>>>(a);
    CKL(); count_launch(e);
    return 0;
}

static int enqueue_layers(infgen_engine *e, bool with_edges, int trace_iter) {
    DecState &s = e->st;
    const int R = e->R;
    float *kv_t = fbuf(e, "kv_t"), *kv_m = fbuf(e, "kv_m"), *kv_a = fbuf(e, "kv_a");
    const size_t kv_t_layer = (size_t)R * RING * 256, kv_m_layer = (size_t)e->P * 256, kv_a_buf = (size_t)R * 256;
    LayerArgs base;
    memset(&base, 0, sizeof(base));
    base.rows = scene_rows(e); base.x = fbuf(e, "x"); base.q = fbuf(e, "q"); base.s = fbuf(e, "s"); base.qr = fbuf(e, "qr");
    base.col_ptr = s.col; base.ring = RING; base.grid_bar = (unsigned *)e->bufs["grid_bar"].p;
    const bool fused = (R + e->row_tile - 1) / e->row_tile <= MAX_CLUSTERS;
    // batches beyond one wave of clusters: the row-tile throughput path (node.cuh), two grid-wide kernels per layer
    if (e->layer_path == 2 || (e->layer_path == 0 && !fused)) {
        const RowSpace rows = scene_rows(e);
        RET(launch_node(e, rows, nullptr, &e->t[0], true, kv_t, true, nullptr));
        for (int i = 0; i < 6; ++i) {
            float *kva = kv_a + (size_t)(i & 1) * kv_a_buf;
            SubArgs t, m, g;
            memset(&t, 0, sizeof(t)); memset(&m, 0, sizeof(m)); memset(&g, 0, sizeof(g));
            t.has_attn = with_edges; t.has_pos = 1; t.kv = kv_t + i * kv_t_layer; t.cnt = s.t_cnt; t.stride = s.W;
            t.src = s.t_src; t.rhat = fbuf(e, "rhat_t");
            m.has_attn = with_edges; m.has_pos = 1; m.kv = kv_m + i * kv_m_layer; m.cnt = s.m_cnt; m.stride = s.max_m;
            m.src = s.m_src; m.rhat = fbuf(e, "rhat_m");
            g.has_attn = with_edges; g.has_pos = 1; g.kv = kva; g.cnt = s.a_cnt; g.start = s.a_start; g.src = s.a_src;
            g.rhat = fbuf(e, "rhat_a");
            float *trace = (trace_iter >= 0 && e->cfg.trace) ? fbuf(e, "trace_layer_out") + ((size_t)trace_iter * 6 + i) * R * 128
                                                             : nullptr;
            RET(launch_attn(e, rows, t, e->t[i]));
            RET(launch_node(e, rows, &e->t[i], &e->m[i], false, nullptr, false, nullptr));
            RET(launch_attn(e, rows, m, e->m[i]));
            RET(launch_node(e, rows, &e->m[i], &e->a[i], true, kva, false, nullptr));
            RET(launch_attn(e, rows, g, e->a[i]));
            RET(launch_node(e, rows, &e->a[i], i < 5 ? &e->t[i + 1] : nullptr, true, kv_t + (size_t)(i + 1) * kv_t_layer, true, trace));
        }
        return 0;
    }
    auto fill = [&](int i, SubArgs &t, SubArgs &m, SubArgs &g) {
        float *kva = kv_a + (size_t)(i & 1) * kv_a_buf;
        t.w = e->t[i].cs_post; t.has_attn = with_edges; t.has_pos = 1; t.elist = 0;
        t.kv = kv_t + i * kv_t_layer; t.cnt = s.t_cnt; t.start = nullptr; t.stride = s.W; t.src = s.t_src;
        t.rhat = fbuf(e, "rhat_t");
        t.pre = make_pre(e->m[i], false, nullptr, false, 0, false);
        m.w = e->m[i].cs_post; m.has_attn = with_edges; m.has_pos = 1; m.elist = 1;
        m.kv = kv_m + i * kv_m_layer; m.cnt = s.m_cnt; m.start = nullptr; m.stride = s.max_m; m.src = s.m_src;
        m.rhat = fbuf(e, "rhat_m");
        m.pre = make_pre(e->a[i], true, kva, false, 0, !fused);
        g.w = e->a[i].cs_post; g.has_attn = with_edges; g.has_pos = 1; g.elist = 2;
        g.kv = kva; g.cnt = s.a_cnt; g.start = s.a_start; g.stride = 0; g.src = s.a_src; g.rhat = fbuf(e, "rhat_a");
        g.grid_sync = fused ? 1 : 0;
        if (i < 5) g.pre = make_pre(e->t[i + 1], true, kv_t + (i + 1) * kv_t_layer, true, 0, !fused);
        if (trace_iter >= 0 && e->cfg.trace)
            g.trace_out = fbuf(e, "trace_layer_out") + ((size_t)trace_iter * 6 + i) * R * 128;
    };
    if (fused) {
        LayerArgs la = base;
        la.pre0 = make_pre(e->t[0], true, kv_t, true, 0, false);
        la.n_sub = 18;
        for (int i = 0; i < 6; ++i) fill(i, la.sub[3 * i], la.sub[3 * i + 1], la.sub[3 * i + 2]);
        CK(cudaMemsetAsync(base.grid_bar, 0, sizeof(unsigned), e->stream));
        RET(launch_layer(e, la, KC_LAYER_STACK));
        return 0;
    }
    for (int i = 0; i < 6; ++i) {
        LayerArgs tm = base, ag = base;
        if (i == 0) tm.pre0 = make_pre(e->t[0], true, kv_t, true, 0, false);
        tm.n_sub = 2; ag.n_sub = 1;
        fill(i, tm.sub[0], tm.sub[1], ag.sub[0]);
        ag.sub[0].elist = 0;
        RET(launch_layer(e, tm, KC_LAYER_TM));
        RET(launch_layer(e, ag, KC_LAYER_A));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// insertion stage (agent_decoder.py:1744-2114).  The number of passes per iteration is data dependent (up to 10, each
// of which may append a row and then runs the heading stage for it).  Three enqueue functions hold the launches:
//     enqueue_insertion_begin   once per iteration
//     enqueue_insertion_pass    the seed query + decision (+ row append) of one pass
//     enqueue_heading_stage     heading / offset of the rows appended by the last pass
// and two drivers sequence them: the host-driven loop of run_insertion (plain launches: traced / profiled runs; one
// 8-byte flag read per pass) and the captured iteration graph, where the loop is a WHILE node and the heading stage an
// IF node inside its body whose condition values k_ins_begin / k_seed_decide set on the device (no host round trip).
// ---------------------------------------------------------------------------------------------------------------
// every active row >= row_lo through a stack of layers WITHOUT edges, keeping the K|V rows of the non-bipartite ones
// subset: 0 = every active row >= row_lo, 1 = the rows appended by the last pass, 2 = those of the pass before
static int enqueue_edgeless(infgen_engine *e, const int *row_lo, float *x, bool seed_stack, int subset = 0) {
    const bool new_only = subset != 0;
    if (!new_only && (e->R + 7) / 8 > MAX_CLUSTERS && e->layer_path != 1) {
        // more than one wave of clusters: the row-tile kernel, one launch per layer (finish layer i, project layer i + 1)
        RowSpace rows = scene_rows(e);
        rows.row_lo = row_lo;
        // (the two stacks run concurrently on two streams: each has its own skip-projection hand-over buffer)
        NodeBufs nb{x, fbuf(e, "q"), fbuf(e, seed_stack ? "s" : "s_ha"), fbuf(e, "qr"), fbuf(e, "zero")};
        const size_t kvl = (size_t)e->R * 256;
        std::vector<const AttnW *> chain;
        std::vector<float *> kv_of;                      // K|V destination of chain[i]'s projections (NULL: none)
        if (seed_stack) {
            for (int i = 0; i < 3; ++i) {
                chain.push_back(&e->occ2sa[i]); kv_of.push_back(nullptr);
                chain.push_back(&e->pt2sa[i]); kv_of.push_back(nullptr);
                if (i == 2) { chain.push_back(&e->a2sa[i]); kv_of.push_back(fbuf(e, "kv_sa") + i * kvl); break; }
                chain.push_back(&e->a2sa[i]); kv_of.push_back(fbuf(e, "kv_sa") + i * kvl);
            }
        } else {
            for (int i = 0; i < 3; ++i) {
                chain.push_back(&e->m[i]); kv_of.push_back(nullptr);
                chain.push_back(&e->a[i]); kv_of.push_back(fbuf(e, "kv_ha") + i * kvl);
            }
        }
        // the last layer of either chain only contributes its K|V projection (its output is never used)
        const int n = (int)chain.size();
        RET(launch_node(e, rows, nullptr, chain[0], kv_of[0] != nullptr, kv_of[0], false, nullptr, &nb, true, KC_INS_LAYER_AGENTS));
        for (int i = 0; i + 1 < n; ++i)
            RET(launch_node(e, rows, chain[i], chain[i + 1], kv_of[i + 1] != nullptr, kv_of[i + 1], false, nullptr, &nb, true,
                            KC_INS_LAYER_AGENTS));
        return 0;
    }
    LayerArgs la;
    memset(&la, 0, sizeof(la));
    la.rows = subset == 1 ? new_rows(e) : (subset == 2 ? prev_rows(e) : scene_rows(e));
    if (!new_only) la.rows.row_lo = row_lo;
    la.x = x; la.ring = RING; la.col_ptr = e->st.col;
    la.no_store = 1;                                    // the chain's output rows are never used: x may be a shared input
    const size_t kvl = (size_t)e->R * 256;
    int n = 0;
    if (seed_stack) {                                   // 3 x {occ2sa, pt2sa, a2sa}: K|V of the a2sa layers
        float *kv = fbuf(e, "kv_sa");
        la.pre0 = make_pre(e->occ2sa[0], false, nullptr, false, 0, false);
        for (int i = 0; i < 3; ++i) {
            SubArgs &o = la.sub[n++]; o.w = e->occ2sa[i].cs_post; o.has_pos = 0;
            o.pre = make_pre(e->pt2sa[i], false, nullptr, false, 0, false);
            SubArgs &p = la.sub[n++]; p.w = e->pt2sa[i].cs_post; p.has_pos = 1;
            p.pre = make_pre(e->a2sa[i], true, kv + i * kvl, false, 0, false);
            if (i == 2) break;                          // the agents' a2sa.2 output is never used
            SubArgs &g = la.sub[n++]; g.w = e->a2sa[i].cs_post; g.has_pos = 1;
            g.pre = make_pre(e->occ2sa[i + 1], false, nullptr, false, 0, false);
        }
    } else {                                            // 3 x {pt2a, a2a} (motion layers 0..2): K|V of a2a
        float *kv = fbuf(e, "kv_ha");
        la.pre0 = make_pre(e->m[0], false, nullptr, false, 0, false);
        for (int i = 0; i < 3; ++i) {
            SubArgs &p = la.sub[n++]; p.w = e->m[i].cs_post; p.has_pos = 1;
            p.pre = make_pre(e->a[i], true, kv + i * kvl, false, 0, false);
            if (i == 2) break;
            SubArgs &g = la.sub[n++]; g.w = e->a[i].cs_post; g.has_pos = 1;
            g.pre = make_pre(e->m[i + 1], false, nullptr, false, 0, false);
        }
    }
    la.n_sub = n;
    const int saved = e->row_tile;
    if (new_only && e->n_scenes <= 4 * MAX_CLUSTERS) e->row_tile = 4;     // see the heading stage
    const int rc = launch_layer(e, la, KC_INS_LAYER_AGENTS);
    e->row_tile = saved;
    return rc;
}

// Run `fn` on the side stream, concurrently with what the caller enqueues on the engine stream until side_join().  Works
// the same inside a stream capture (parallel branches of the graph) and outside; profiled runs stay serial (the per-class
// events are recorded on the engine stream).
template <typename Fn>
static int side_fork(infgen_engine *e, Fn fn) {
    if (e->profile) return fn();
    cudaStream_t main = e->stream;
    CK(cudaEventRecord(e->ev_fork, main));
    CK(cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0));
    e->stream = e->side_stream;
    const int rc = fn();
    e->stream = main;
    RET(rc);
    CK(cudaEventRecord(e->ev_join, e->side_stream));
    return 0;
}
static int side_join(infgen_engine *e) {
    if (e->profile) return 0;
    CK(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    return 0;
}
// the same on the second side stream (a branch that stays open across several fork / join pairs of the first)
template <typename Fn>
static int side2_fork(infgen_engine *e, Fn fn) {
    if (e->profile) return fn();
    cudaStream_t main = e->stream;
    CK(cudaEventRecord(e->ev_fork2, main));
    CK(cudaStreamWaitEvent(e->side_stream2, e->ev_fork2, 0));
    e->stream = e->side_stream2;
    const int rc = fn();
    e->stream = main;
    RET(rc);
    CK(cudaEventRecord(e->ev_join2, e->side_stream2));
    return 0;
}
static int side2_join(infgen_engine *e) {
    if (e->profile) return 0;
    CK(cudaStreamWaitEvent(e->stream, e->ev_join2, 0));
    return 0;
}
template <typename Fn>
static int side3_fork(infgen_engine *e, Fn fn) {
    if (e->profile) return fn();
    cudaStream_t main = e->stream;
    CK(cudaEventRecord(e->ev_fork3, main));
    CK(cudaStreamWaitEvent(e->side_stream3, e->ev_fork3, 0));
    e->stream = e->side_stream3;
    const int rc = fn();
    e->stream = main;
    RET(rc);
    CK(cudaEventRecord(e->ev_join3, e->side_stream3));
    return 0;
}
static int side3_join(infgen_engine *e) {
    if (e->profile) return 0;
    CK(cudaStreamWaitEvent(e->stream, e->ev_join3, 0));
    return 0;
}

// Inputs of the next seed query that do not come out of attention layers: the occupancy node (grid occupancy ->
// seed_agent_occ_embed -> K|V of the three occ2sa layers, :1850-1859) and the query row's input feature.  They only depend on
// the grid cells of the rows at column cur, which are final the moment a row is appended (k_seed_decide), so the kernel runs
// beside the edge-less stacks at the start of the stage and beside the heading stage of an appended row - never in the chain
// query -> heads -> decision of a pass.
static int enqueue_seed_prepare(infgen_engine *e) {
    SeedPrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.s = e->st; pa.q = e->ins; pa.occ_embed = e->h_occ_embed;
    for (int i = 0; i < 3; ++i) pa.occ2sa[i] = e->occ2sa[i];
    {
        ProfScope ps(e, KC_INSERT);
        k_seed_prepare<<<e->n_scenes, NT, 0, e->stream>>>(pa);
    }
    CKL(); count_launch(e);
    return 0;
}

static int enqueue_insertion_begin(infgen_engine *e) {
    DecState &s = e->st;
    InsState &q = e->ins;
    const int ns = e->n_scenes, R = e->R;
    cudaStream_t st = e->stream;
    {
        ProfScope ps(e, KC_INSERT);
        k_ins_begin<<<ns, NT, 0, st>>>(s, q);
    }
    CKL(); count_launch(e);
    // Every row through the two edge-less stacks of the stage -> the K|V rows later passes attend to: the seed stack
    // (3 x {occ2sa, pt2sa, a2sa}: K|V of the a2sa layers) on the engine stream; on side streams the relative embeddings of
    // the map -> seed and agent -> seed edges (the seed pose is the ego pose for every pass of the iteration; rows appended
    // by a pass add their own edge), the heading stack (motion layers 0..2 without edges: K|V of a2a for the heading stage
    // of appended rows) and the occupancy node of the first pass
    // The cluster kernel keeps the residual rows in shared memory and (no_store) never writes them back, so both chains
    // read x itself; the row-tile kernels of large batches update their residual stream in global memory and get copies.
    const bool chain_copies = (R + 7) / 8 > MAX_CLUSTERS && e->layer_path != 1;
    float *x_in_sa = chain_copies ? fbuf(e, "x_sa") : fbuf(e, "x"), *x_in_ha = chain_copies ? fbuf(e, "x_ha") : fbuf(e, "x");
    if (chain_copies) {
        ProfScope ps(e, KC_INSERT);
        k_copy_new_rows2<<<ns, 128, 0, st>>>(s, nullptr, fbuf(e, "x"), fbuf(e, "x_sa"), fbuf(e, "x_ha"));
        CKL(); count_launch(e);
    }
    RET(side2_fork(e, [&]() -> int { return enqueue_seed_prepare(e); }));
    RET(side_fork(e, [&]() -> int {
        FourierArgs fj[2];
        memset(fj, 0, sizeof(fj));
        fj[0].normalize = 1; fj[0].dim = 3; fj[0].n_slots = ns * SEED_MAP_MAX; fj[0].cnt = q.ps_cnt; fj[0].stride = SEED_MAP_MAX;
        fj[0].raw = q.ps_raw; fj[0].w = e->f_ps; fj[0].out = fbuf(e, "rhat_ps");
        fj[1].normalize = 1; fj[1].dim = 3; fj[1].n_slots = ns * q.as_stride; fj[1].cnt = q.as_cnt; fj[1].stride = q.as_stride;
        fj[1].raw = q.as_raw; fj[1].w = e->f_as; fj[1].out = fbuf(e, "rhat_as");
        return launch_fourier(e, fj, 2, KC_INS_FOURIER);
    }));
    RET(side3_fork(e, [&]() -> int { return enqueue_edgeless(e, nullptr, x_in_ha, false); }));
    RET(enqueue_edgeless(e, nullptr, x_in_sa, true));
    RET(side_join(e));
    RET(side3_join(e));
    return side2_join(e);
}

static int enqueue_insertion_pass(infgen_engine *e) {
    DecState &s = e->st;
    InsState &q = e->ins;
    const int ns = e->n_scenes, R = e->R;
    cudaStream_t st = e->stream;
    {   // the query rows: 3 x {occ2sa, pt2sa, a2sa} with their edges
        LayerArgs la;
        memset(&la, 0, sizeof(la));
        la.rows.n_total = ns * q.seed_stride; la.rows.cap = q.seed_stride; la.rows.n_rows = q.active;
        la.rows.row_lo = nullptr;
        la.x = q.x_seed; la.ring = RING;
        la.pre0 = make_pre(e->occ2sa[0], false, nullptr, false, 0, false);
        const size_t kvl = (size_t)R * 256, kvm = (size_t)e->P * 256;
        const int qshift = q.seed_stride == SEED_ROW_STRIDE ? 2 : 0, qwide = q.seed_stride == SEED_ROW_STRIDE ? 1 : 0;
        // one scene per tile: the row appended by the previous pass rides along (its a2sa K|V rows come out of this launch)
        const bool ride = e->ins_ride;
        if (ride) { la.x2 = fbuf(e, "x_sa"); la.ride_row = q.new_row; la.ride_cap = s.cap; }
        int n = 0;
        for (int i = 0; i < 3; ++i) {
            SubArgs &o = la.sub[n++];
            o.w = e->occ2sa[i].cs_post; o.has_attn = 1; o.has_pos = 0; o.elist = 0; o.row_shift = qshift; o.wide = qwide;
            o.kv = q.kv_occ + (size_t)i * ns * 256; o.cnt = q.one_cnt; o.start = nullptr; o.stride = 1; o.src = q.occ_src;
            o.pre = make_pre(e->pt2sa[i], false, nullptr, false, 0, false);
            SubArgs &p = la.sub[n++];
            p.w = e->pt2sa[i].cs_post; p.has_attn = 1; p.has_pos = 1; p.elist = 1; p.row_shift = qshift; p.wide = qwide;
            p.kv = fbuf(e, "kv_ms") + i * kvm; p.cnt = q.ps_cnt; p.start = nullptr; p.stride = SEED_MAP_MAX;
            p.src = q.ps_src; p.rhat = fbuf(e, "rhat_ps");
            p.pre = make_pre(e->a2sa[i], ride, ride ? fbuf(e, "kv_sa") + i * kvl : nullptr, false, 0, false);
            SubArgs &g = la.sub[n++];
            g.w = e->a2sa[i].cs_post; g.has_attn = 1; g.has_pos = 1; g.elist = 2; g.row_shift = qshift; g.wide = qwide;
            g.kv = fbuf(e, "kv_sa") + i * kvl; g.cnt = q.as_cnt; g.start = nullptr; g.stride = q.as_stride;
            g.src = q.as_src; g.rhat = fbuf(e, "rhat_as");
            if (i < 2) g.pre = make_pre(e->occ2sa[i + 1], false, nullptr, false, 0, false);
        }
        la.n_sub = n;
        const int saved = e->row_tile;
        e->row_tile = 4;
        int rc = launch_layer(e, la, KC_INS_LAYER_QUERY);
        e->row_tile = saved;
        RET(rc);
    }
    {   // the three grid-sized heads of the query rows in one launch
        MlpLayerArgs la;
        memset(&la, 0, sizeof(la));
        la.n = ns * q.seed_stride; la.x = q.x_seed;
        la.w = e->h_seed_pos; la.out = q.pos_logits;
        la.w2 = e->h_ag_occ; la.out2 = q.ag_occ_logits;
        la.w3 = e->h_pt_occ; la.out3 = q.pt_occ_logits;
        // and the three small ones (state, type, shape): [rows][2], [rows][3], [rows][3]
        la.w4 = e->h_seed_state; la.out4 = q.small_logits;
        la.w5 = e->h_seed_type; la.out5 = q.small_logits + (size_t)la.n * 2;
        la.w6 = e->h_seed_shape; la.out6 = q.small_logits + (size_t)la.n * 5;
        la.single_stride = q.seed_stride == SEED_ROW_STRIDE ? SEED_ROW_STRIDE : 0;
        const int npad = std::max(la.w.n_pad, std::max(la.w2.n_pad, la.w3.n_pad));
        ProfScope ps(e, KC_INS_HEADS);
        k_mlp_layer<<<dim3((la.n + HM - 1) / HM, npad / 128, 6), NT_S, MLP_LAYER_SMEM, st>>>(la);
        CKL(); count_launch(e);
    }
    SeedDecideArgs da;
    memset(&da, 0, sizeof(da));
    da.s = s; da.q = q;
    if (e->bufs.count("tstamp") && e->bufs["tstamp"].p) da.tstamp = (long long *)e->bufs["tstamp"].p + 384;
    {
        ProfScope ps(e, KC_INSERT);
        k_seed_decide<<<ns, NT, 0, st>>>(da);
    }
    CKL(); count_launch(e);
    return 0;
}

// heading stage of the rows appended by the last pass (:2003-2074)
static int enqueue_heading_stage(infgen_engine *e) {
    DecState &s = e->st;
    InsState &q = e->ins;
    const int ns = e->n_scenes, R = e->R;
    cudaStream_t st = e->stream;
    float *x = fbuf(e, "x"), *x_sa = fbuf(e, "x_sa"), *x_ha = fbuf(e, "x_ha");
    // the occupancy node of the next pass: the appended row's cell is known
    // and, behind it, the heading-stack K|V rows of the row the pass before appended (this row's a2a layers may attend to it)
    // (one scene per tile only; packed batches project the new rows at the end of this stage, beside their seed-stack chain)
    RET(side2_fork(e, [&]() -> int {
        {   // (first: the grid-sized records of the insertion read the occupancy the query saw)
            ProfScope ps(e, KC_INSERT);
            k_seed_records<<<ns, NT, 0, e->stream>>>(s, q);
        }
        CKL(); count_launch(e);
        return enqueue_seed_prepare(e);
    }));
    if (e->ins_ride) RET(side3_fork(e, [&]() -> int { return enqueue_edgeless(e, nullptr, x_ha, false, 2); }));
    // the new row's edges and their relative embeddings only need its pose: on the side stream, concurrently with its
    // categorical / column embedding (two chains of ~50 us each per inserted agent)
    RET(side_fork(e, [&]() -> int {
        {
            ProfScope ps(e, KC_INSERT);
            k_new_edges<<<ns, NEW_EDGE_NT, 0, e->stream>>>(s, q);
        }
        CKL(); count_launch(e);
        FourierArgs hj[2];
        memset(hj, 0, sizeof(hj));
        hj[0].normalize = 1; hj[0].dim = 3; hj[0].n_slots = ns * NEW_MAP_MAX; hj[0].cnt = q.hp_cnt_s; hj[0].stride = NEW_MAP_MAX;
        hj[0].raw = q.hp_raw; hj[0].w = e->f_m; hj[0].out = fbuf(e, "rhat_hp");
        hj[1].normalize = 1; hj[1].dim = 3; hj[1].n_slots = ns * NEW_AGENT_MAX; hj[1].cnt = q.ha_cnt_s; hj[1].stride = NEW_AGENT_MAX;
        hj[1].raw = q.ha_raw; hj[1].w = e->f_a; hj[1].out = fbuf(e, "rhat_ha");
        return launch_fourier(e, hj, 2, KC_INS_FOURIER);
    }));
    // categorical embedding row of the new agent: type_a_emb[type] + shape_emb(shape)
    MlpEmbArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.rows = new_rows(e); ma.w = e->e_shape; ma.kin = 3; ma.k4 = 1;
    ma.x = q.shape_rows; ma.x_ld = 3; ma.out = fbuf(e, "cat_tab"); ma.out_ld = 128;
    RET(launch_mlp_embed(e, ma, KC_INSERT));
    {
        ProfScope ps(e, KC_INSERT);
        k_add_type_emb_rows<<<ns, 128, 0, st>>>(s, q.row_lo, fbuf(e, "cat_tab"), e->type_emb);
    }
    CKL(); count_launch(e);
    RET(enqueue_embed_rows(e, q.row_lo));       // feature of the new row with the dummy heading
    RET(side_join(e));
    if (e->ins_ride) RET(side3_join(e));
    {   // the new rows through 3 x {pt2a, a2a} with their 10 m neighbourhoods
        LayerArgs la;
        memset(&la, 0, sizeof(la));
        la.rows = new_rows(e); la.x = x; la.ring = RING; la.col_ptr = s.col;
        la.pre0 = make_pre(e->m[0], false, nullptr, false, 0, false);
        const size_t kvl = (size_t)R * 256, kvm = (size_t)e->P * 256;
        int n = 0;
        for (int i = 0; i < 3; ++i) {
            SubArgs &p = la.sub[n++];
            p.w = e->m[i].cs_post; p.has_attn = 1; p.has_pos = 1; p.elist = 0;
            p.kv = fbuf(e, "kv_m") + i * kvm; p.cnt = q.hp_cnt; p.start = q.hp_start; p.src = q.hp_src;
            p.rhat = fbuf(e, "rhat_hp");
            p.pre = make_pre(e->a[i], false, nullptr, false, 0, false);
            SubArgs &g = la.sub[n++];
            g.w = e->a[i].cs_post; g.has_attn = 1; g.has_pos = 1; g.elist = 1;
            g.kv = fbuf(e, "kv_ha") + i * kvl; g.cnt = q.ha_cnt; g.start = q.ha_start; g.src = q.ha_src;
            g.rhat = fbuf(e, "rhat_ha");
            if (i < 2) g.pre = make_pre(e->m[i + 1], false, nullptr, false, 0, false);
        }
        la.n_sub = n;
        // a handful of rows: tiles of 4 (every projection phase of k_layer<4> is shorter than its k_layer<8> counterpart)
        const int saved = e->row_tile;
        if (e->n_scenes <= 4 * MAX_CLUSTERS) e->row_tile = 4;
        const int rc = launch_layer(e, la, KC_INS_LAYER_NEW);
        e->row_tile = saved;
        RET(rc);
    }
    HeadFinalArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.s = s; ha.q = q; ha.x = x; ha.h_heading = e->h_seed_heading; ha.h_offset = e->h_seed_offset;
    {
        ProfScope ps(e, KC_INSERT);
        k_head_finalize<<<ns, NT, 0, st>>>(ha);
    }
    CKL(); count_launch(e);
    // Final feature of the new row (:2086-2097), also written to the inputs of the two edge-less stacks (the new rows become
    // sources of later passes), and the relative embedding of its edge towards the query row.
    auto seed_edge_embedding = [&]() -> int {
        FourierArgs fj;
        memset(&fj, 0, sizeof(fj));
        fj.normalize = 1; fj.dim = 3; fj.n_slots = ns; fj.slot_list = q.as_new_list; fj.n_list = q.as_new_n;
        fj.raw = q.as_raw; fj.w = e->f_as; fj.out = fbuf(e, "rhat_as");
        fj.ffma = ns <= FM;                  // one tile of a few slots: the FFMA kernel (GEMV path for a single slot)
        return launch_fourier(e, &fj, 1, KC_INS_FOURIER);
    };
    if (e->ins_ride) {
        // One scene per tile: the edge embedding only needs the pose and runs beside the final embedding.  The row's
        // seed-stack K|V rows come out of the next seed query (ride-along row), its heading-stack K|V rows are left to the
        // next heading stage of the iteration (prev_rows above; the next iteration projects every row again).
        RET(side_fork(e, seed_edge_embedding));
        RET(enqueue_embed_rows(e, q.row_lo, x_sa, x_ha));
        RET(side_join(e));
        return side2_join(e);                    // (records + occupancy node of the next pass)
    }
    // packed query rows: both edge-less chains of the new rows here, on two streams
    RET(enqueue_embed_rows(e, q.row_lo, x_sa, x_ha));
    RET(side_fork(e, [&]() -> int {
        RET(seed_edge_embedding());
        return enqueue_edgeless(e, q.row_lo, x_ha, false, 1);
    }));
    RET(enqueue_edgeless(e, q.row_lo, x_sa, true, 1));
    RET(side_join(e));
    return side2_join(e);
}

// host-driven loop (plain launches)
static int run_insertion(infgen_engine *e) {
    InsState &q = e->ins;
    cudaStream_t st = e->stream;
    q.use_cond = 0;
    RET(enqueue_insertion_begin(e));
    // debug tools: INFGEN_DEBUG_STOP_PASS="<iteration>:<pass>" stops the rollout right after that pass's decision, so
    // that the buffers of the seed query can be compared with the oracle (tools/debug_long_insertion.py)
    int stop_iter = -1, stop_pass = -1;
    if (const char *sp = getenv("INFGEN_DEBUG_STOP_PASS")) sscanf(sp, "%d:%d", &stop_iter, &stop_pass);
    for (int pass = 0; pass < INSERT_LIMIT; ++pass) {
        RET(enqueue_insertion_pass(e));
        if (e->iters_done == stop_iter && pass == stop_pass) {
            CK(cudaStreamSynchronize(st));
            return fail(INFGEN_ERR_STATE, "debug stop after pass %d of iteration %d", pass, stop_iter);
        }
        int flags[2] = {0, 0};
        CK(cudaMemcpyAsync(flags, q.flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (flags[1]) RET(enqueue_heading_stage(e));
        if (!flags[0]) break;
    }
    return 0;
}

// ---- conditional graph nodes populated by stream capture (tools/probe/cond_graph.cu is the stand-alone probe) ---------
// While the engine stream is capturing, add_cond_node appends a WHILE / IF node behind the work captured so far and
// queues its body; drain_cond_bodies captures the queued bodies (which may queue nested ones) once the enclosing
// capture has ended.  Condition handles belong to the top-level graph.
struct CondBody { cudaGraph_t graph; int (*fn)(infgen_engine *); };
static thread_local std::vector<CondBody> g_cond_bodies;
static int add_cond_node(infgen_engine *e, cudaGraphConditionalHandle h, cudaGraphConditionalNodeType type,
                         int (*body)(infgen_engine *)) {
    cudaStreamCaptureStatus status;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t n_deps = 0;
    CK(cudaStreamGetCaptureInfo(e->stream, &status, nullptr, &g, &deps, &n_deps));
    if (status != cudaStreamCaptureStatusActive) return fail(INFGEN_ERR_STATE, "conditional node outside a stream capture");
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = type;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, g, deps, n_deps, &p));
    CK(cudaStreamUpdateCaptureDependencies(e->stream, &node, 1, cudaStreamSetCaptureDependencies));
    g_cond_bodies.push_back(CondBody{p.conditional.phGraph_out[0], body});
    return 0;
}
// node_counts: nodes of every body in the order they were captured
static int drain_cond_bodies(infgen_engine *e, std::vector<size_t> &node_counts) {
    while (!g_cond_bodies.empty()) {
        const CondBody b = g_cond_bodies.back();
        g_cond_bodies.pop_back();
        CK(cudaStreamBeginCaptureToGraph(e->stream, b.graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        const int rc = b.fn(e);
        const cudaError_t ce = cudaStreamEndCapture(e->stream, nullptr);
        if (rc != 0) { g_cond_bodies.clear(); return rc; }
        if (ce != cudaSuccess) {
            g_cond_bodies.clear();
            return fail(INFGEN_ERR_CUDA, "capture of a conditional body failed: %s", cudaGetErrorString(ce));
        }
        size_t n = 0;
        CK(cudaGraphGetNodes(b.graph, nullptr, &n));
        node_counts.push_back(n);
    }
    return 0;
}
static int cond_body_pass(infgen_engine *e) {
    RET(enqueue_insertion_pass(e));
    return add_cond_node(e, e->ins.h_new, cudaGraphCondTypeIf, enqueue_heading_stage);
}
// insertion stage inside the captured iteration graph
static int enqueue_insertion_graph(infgen_engine *e) {
    RET(enqueue_insertion_begin(e));
    return add_cond_node(e, e->ins.h_pass, cudaGraphCondTypeWhile, cond_body_pass);
}

// edges whose destination is column col + col_add and their relative embeddings
static int enqueue_edges(infgen_engine *e, int col_add) {
    DecState &s = e->st;
    const int R = e->R;
    {
        ProfScope ps(e, KC_EDGE_BUILD);
        k_edge_build<<<(R * 3 + NWARP - 1) / NWARP, NT, 0, e->stream>>>(s, col_add);
    }
    CKL(); count_launch(e);
    FourierArgs fj[3];
    memset(fj, 0, sizeof(fj));
    fj[0].normalize = 1; fj[0].dim = 3;       // the large one first: agent<->agent
    fj[0].n_slots = R * s.a_stride; fj[0].cnt = s.a_cnt; fj[0].stride = s.a_stride;
    fj[0].raw = s.a_raw; fj[0].w = e->f_a; fj[0].out = fbuf(e, "rhat_a");
    if (e->fourier_tc && (R * s.a_stride + ftc::TM - 1) / ftc::TM > 148) {
        // batches: walk a compact list of the valid agent<->agent slots (a single scene's tiles fit one wave anyway)
        int *list = (int *)e->bufs["a_slots"].p, *n_list = list + (size_t)R * s.a_stride;
        k_slot_compact<<<1, 1024, 0, e->stream>>>(s.a_cnt, R, s.a_stride, list, n_list, n_list + 4);
        CKL(); count_launch(e);
        fj[0].slot_list = list; fj[0].n_list = n_list;
    }
    fj[1].normalize = 1; fj[1].dim = 4;
    fj[1].n_slots = R * s.W; fj[1].cnt = s.t_cnt; fj[1].stride = s.W; fj[1].raw = s.t_raw; fj[1].w = e->f_t;
    if (e->fourier_tc && s.W <= 16 && !getenv("INFGEN_NO_DIM_TABLE")) {
        // three input dims through the tensor core, the column offset from the table
        fj[1].dim = 3; fj[1].raw_stride = 4; fj[1].w = e->f_t3; fj[1].dim_table = e->t_dim_table; fj[1].table_n = 16;
    }
    fj[1].out = fbuf(e, "rhat_t");
    fj[2].normalize = 1; fj[2].dim = 3;
    fj[2].n_slots = R * s.max_m; fj[2].cnt = s.m_cnt; fj[2].stride = s.max_m; fj[2].raw = s.m_raw; fj[2].w = e->f_m;
    fj[2].out = fbuf(e, "rhat_m");
    return launch_fourier(e, fj, 3, KC_FOURIER);
}

// One decode iteration.  With the insertion stage enabled the edges of column cur are built first (rows may have been
// appended since the last iteration).  Motion-only engines build the edges of the NEXT column right after the advance,
// on a side stream, concurrently with that column's embedding (8 CTAs): both only depend on the new poses.
static int enqueue_iteration(infgen_engine *e, int trace_iter) {
    DecState &s = e->st;
    const int R = e->R;
    if (!e->early_edges) RET(enqueue_edges(e, 0));
    RET(enqueue_layers(e, true, trace_iter));
    HeadArgs ha;
    memset(&ha, 0, sizeof(ha));
    ha.rows = scene_rows(e); ha.x = fbuf(e, "x"); ha.tok = e->h_tok; ha.st = e->h_state;
    ha.part_v = fbuf(e, "part_v"); ha.part_i = (int *)e->bufs["part_i"].p;
    ha.part_m = fbuf(e, "part_m"); ha.part_s = fbuf(e, "part_s"); ha.state_logits = fbuf(e, "state_logits");
    if (e->cfg.trace && trace_iter >= 0) {
        ha.trace_head_in = fbuf(e, "trace_head_in") + (size_t)trace_iter * R * 128;
        ha.trace_logits = fbuf(e, "trace_token_logits") + (size_t)trace_iter * R * e->cfg.token_size;
        ha.trace_state = fbuf(e, "trace_state_logits") + (size_t)trace_iter * R * 3;
    }
    {
        ProfScope ps(e, KC_HEADS);
        if (R > 512) k_heads<16><<<dim3((R + 15) / 16, NSLICE + 1), NT_S, heads_smem<16>(), e->stream>>>(ha);
        else k_heads<HM><<<dim3((R + HM - 1) / HM, NSLICE + 1), NT_S, HEADS_SMEM, e->stream>>>(ha);
    }
    CKL(); count_launch(e);
    {
        ProfScope ps(e, KC_ADVANCE);
        k_advance<<<(R + NWARP - 1) / NWARP, NT, 0, e->stream>>>(s);
    }
    CKL(); count_launch(e);
    if (e->early_edges && !e->profile) {
        cudaStream_t main = e->stream;
        CK(cudaEventRecord(e->ev_fork, main));
        CK(cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0));
        e->stream = e->side_stream;
        int rc = enqueue_edges(e, 1);
        e->stream = main;
        RET(rc);
        CK(cudaEventRecord(e->ev_join, e->side_stream));
        RET(enqueue_embed_column(e, 1));
        CK(cudaStreamWaitEvent(main, e->ev_join, 0));
    } else {
        RET(enqueue_embed_column(e, 1));
        if (e->early_edges) RET(enqueue_edges(e, 1));
    }
    {
        ProfScope ps(e, KC_MISC);
        k_next_iter<<<1, 1, 0, e->stream>>>(s.col, s.iter);
    }
    CKL(); count_launch(e);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int32_t infgen_abi_version(void) { return INFGEN_ABI_VERSION; }
const char *infgen_last_error(void) { return g_err; }

int32_t infgen_weight_count(void) { build_layout(); return (int32_t)g_layout.size(); }
const char *infgen_weight_name(int32_t i) {
    build_layout();
    return (i >= 0 && i < (int)g_layout.size()) ? g_layout[i].name.c_str() : nullptr;
}
int64_t infgen_weight_offset(const char *name) {
    build_layout();
    auto it = g_index.find(name);
    return it == g_index.end() ? -1 : g_layout[it->second].offset;
}
int64_t infgen_weight_numel(const char *name) {
    build_layout();
    auto it = g_index.find(name);
    return it == g_index.end() ? -1 : g_layout[it->second].numel;
}
int64_t infgen_weight_blob_floats(void) { build_layout(); return g_total; }

static int build_tables(infgen_engine *e) {
    const int V = e->cfg.token_size, G = e->cfg.grid_size;
    CK(cudaMalloc(&e->tok_tab, (size_t)3 * (V + 2) * 128 * sizeof(float)));
    CK(cudaMalloc(&e->grid_tab, (size_t)(G + 1) * 128 * sizeof(float)));
    for (int ty = 0; ty < 3; ++ty) {          // agent_decoder.py:347-362: MLPEmbedding of the last sub-step box, + BOS, + none
        MlpEmbArgs ma;
        memset(&ma, 0, sizeof(ma));
        ma.rows = flat_rows(V); ma.w = e->e_tok[ty]; ma.kin = 8; ma.k4 = 2;
        ma.x = e->vocab + (size_t)ty * V * 48 + 40; ma.x_ld = 48;
        ma.out = e->tok_tab + (size_t)ty * (V + 2) * 128; ma.out_ld = 128;
        RET(launch_mlp_embed(e, ma));
        CK(cudaMemcpyAsync(e->tok_tab + ((size_t)ty * (V + 2) + V) * 128, W(e, "bos_token_emb"), 128 * sizeof(float),
                           cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(e->tok_tab + ((size_t)ty * (V + 2) + V + 1) * 128, W(e, "no_token_emb"), 128 * sizeof(float),
                           cudaMemcpyDeviceToDevice, e->stream));
    }
    MlpEmbArgs ga;                            // agent_decoder.py:371-373
    memset(&ga, 0, sizeof(ga));
    ga.rows = flat_rows(G); ga.w = e->e_grid; ga.kin = 2; ga.k4 = 1; ga.x = e->grid_cells; ga.x_ld = 2;
    ga.out = e->grid_tab; ga.out_ld = 128;
    RET(launch_mlp_embed(e, ga));
    CK(cudaMemcpyAsync(e->grid_tab + (size_t)G * 128, W(e, "invalid_offset_token_emb"), 128 * sizeof(float),
                       cudaMemcpyDeviceToDevice, e->stream));
    return 0;
}

// input feature of the seed query row: `_build_agent_feature(num_step, device, None, None, state_index=invalid)`
// (agent_decoder.py:449-509 as called at :1817-1821) - no-token embedding, all-invalid motion vector, seed type / 0.1
// shape, invalid state, the CENTRE cell's grid embedding; the same for every scene, pass and column
__global__ void k_seed_raw(float *raw) {
    raw[0] = norm2(-2.f, -2.f);
    raw[1] = angle_between(cosf(0.f), sinf(0.f), -2.f, -2.f);
}
static int build_seed_feature(infgen_engine *e) {
    const int V = e->cfg.token_size, G = e->cfg.grid_size;
    cudaStream_t st = e->stream;
    float *tmp = nullptr;                                // [shape row 4 | cat 128 | raw 2(+2) | xa 128 | fused in 512]
    CK(cudaMalloc(&tmp, 1024 * sizeof(float)));
    CK(cudaMalloc(&e->seed_feat, 128 * sizeof(float)));
    float *shape_row = tmp, *cat = tmp + 4, *raw = tmp + 132, *xa = tmp + 136, *fin = tmp + 264;
    k_fill_shape_rows<<<1, 32, 0, st>>>(shape_row, nullptr, 0);          // one row of 0.1
    MlpEmbArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.rows = flat_rows(1); ma.w = e->e_shape; ma.kin = 3; ma.k4 = 1; ma.x = shape_row; ma.x_ld = 3; ma.out = cat; ma.out_ld = 128;
    RET(launch_mlp_embed(e, ma));
    k_add_type_emb<<<1, 128, 0, st>>>(cat, e->type_emb, nullptr, 0, 3);  // + type_a_emb['seed']
    k_seed_raw<<<1, 1, 0, st>>>(raw);
    FourierArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.n_slots = 1; fa.dim = 2; fa.raw = raw; fa.w = e->f_x; fa.cat_tab = cat; fa.out = xa; fa.normalize = 0;
    RET(launch_fourier(e, &fa, 1));
    CK(cudaMemcpyAsync(fin, e->tok_tab + (size_t)(V + 1) * 128, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));   // no-token
    CK(cudaMemcpyAsync(fin + 128, xa, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(fin + 256, e->state_emb, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));                   // invalid = 0
    CK(cudaMemcpyAsync(fin + 384, e->grid_tab + (size_t)(G / 2) * 128, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    memset(&ma, 0, sizeof(ma));
    ma.rows = flat_rows(1); ma.w = e->e_fusion; ma.kin = 512; ma.k4 = 128; ma.x = fin; ma.x_ld = 512; ma.out = e->seed_feat;
    ma.out_ld = 128;
    RET(launch_mlp_embed(e, ma));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(tmp));
    CKL();
    return 0;
}

int32_t infgen_create(const infgen_config *cfg, const float *weights, int64_t n_floats, const float *grid_cells,
                      const float *vocab, infgen_engine **out) {
    build_layout();
    if (!cfg || !weights || !grid_cells || !vocab || !out) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    if (cfg->abi_version != INFGEN_ABI_VERSION)
        return fail(INFGEN_ERR_INVALID_ARG, "ABI version %d != library %d", cfg->abi_version, INFGEN_ABI_VERSION);
    if (n_floats != g_total) return fail(INFGEN_ERR_INVALID_ARG, "weight blob has %lld floats, expected %lld",
                                          (long long)n_floats, (long long)g_total);
    if (!cfg->disable_insertion && (cfg->grid_size != GRID_SIZE || cfg->insert_beam_size < 1 ||
                                    cfg->insert_beam_size > INSERT_LIMIT || cfg->angle_interval <= 0.f))
        return fail(INFGEN_ERR_INVALID_ARG, "unsupported insertion configuration (grid=%d insert_beam=%d)",
                    cfg->grid_size, cfg->insert_beam_size);
    if (cfg->num_layers != 6 || cfg->token_size != TOKEN_SIZE || cfg->window + 1 > RING || cfg->window > 32 ||
        cfg->hist_cols < 1 || cfg->motion_beam_size < 1 || cfg->motion_beam_size > KTOP || cfg->max_pl2a_neighbors > 32)
        return fail(INFGEN_ERR_INVALID_ARG, "unsupported configuration (layers=%d tokens=%d window=%d beam=%d)",
                    cfg->num_layers, cfg->token_size, cfg->window, cfg->motion_beam_size);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(INFGEN_ERR_NO_DEVICE, "no CUDA device: the decode path has no CPU fallback");
    }
    CK(cudaSetDevice(cfg->device));
    infgen_engine *e = new infgen_engine();
    e->cfg = *cfg;
    CK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    CK(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&e->side_stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork2, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&e->side_stream3, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork3, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join3, cudaEventDisableTiming));
    e->early_edges = cfg->disable_insertion != 0 && !getenv("INFGEN_NO_EARLY_EDGES");   // (debug tools compare per-iteration edges)
    CK(cudaMalloc(&e->blob, (size_t)g_total * sizeof(float)));
    CK(cudaMemcpyAsync(e->blob, weights, (size_t)g_total * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMalloc(&e->grid_cells, (size_t)cfg->grid_size * 2 * sizeof(float)));
    CK(cudaMemcpyAsync(e->grid_cells, grid_cells, (size_t)cfg->grid_size * 2 * sizeof(float), cudaMemcpyHostToDevice,
                       e->stream));
    CK(cudaMalloc(&e->vocab, (size_t)3 * cfg->token_size * 48 * sizeof(float)));
    CK(cudaMemcpyAsync(e->vocab, vocab, (size_t)3 * cfg->token_size * 48 * sizeof(float), cudaMemcpyHostToDevice,
                       e->stream));
    CK(cudaMalloc(&e->d_err, sizeof(int)));
    CK(cudaMemsetAsync(e->d_err, 0, sizeof(int), e->stream));
    RET(build_cluster_weights(e, weights));
    for (int i = 0; i < 6; ++i) {
        e->t[i] = make_attn(e, "t_attn_layers." + std::to_string(i), true);
        e->m[i] = make_attn(e, "pt2a_attn_layers." + std::to_string(i), true);
        e->a[i] = make_attn(e, "a2a_attn_layers." + std::to_string(i), true);
    }
    e->f_t = make_fourier(e, "r_t_emb", 4); e->f_m = make_fourier(e, "r_pt2a_emb", 3);
    e->f_a = make_fourier(e, "r_a2a_emb", 3); e->f_x = make_fourier(e, "x_a_emb", 2);
    // the 4th temporal input is the column offset -1 .. -12 (agent_decoder.py:607): its per-dim MLP is a 12-row table
    e->f_t3 = make_fourier(e, "r_t_emb", 3);
    CK(cudaMalloc(&e->t_dim_table, (size_t)16 * 128 * sizeof(float)));
    k_fourier_dim_table<<<16, 128, 0, e->stream>>>(e->f_t, 3, e->t_dim_table);
    CKL();
    e->e_shape = make_mlp_emb(e, "shape_emb"); e->e_fusion = make_mlp_emb(e, "fusion_emb");
    e->e_tok[0] = make_mlp_emb(e, "token_emb_veh"); e->e_tok[1] = make_mlp_emb(e, "token_emb_ped");
    e->e_tok[2] = make_mlp_emb(e, "token_emb_cyc"); e->e_grid = make_mlp_emb(e, "token_emb_grid");
    e->h_tok = make_head(e, "token_predict_head", 128, cfg->token_size);
    e->h_state = make_head(e, "state_predict_head", 128, 3);
    e->type_emb = W(e, "type_a_emb"); e->state_emb = W(e, "state_a_emb");
    for (int i = 0; i < 3; ++i) {
        e->occ2sa[i] = make_attn(e, "occ2sa_attn_layers." + std::to_string(i), false);
        e->pt2sa[i] = make_attn(e, "pt2sa_attn_layers." + std::to_string(i), true);
        e->a2sa[i] = make_attn(e, "a2sa_attn_layers." + std::to_string(i), true);
    }
    e->f_ps = make_fourier(e, "r_pt2sa_emb", 3); e->f_as = make_fourier(e, "r_a2sa_emb", 3);
    e->h_seed_state = make_head(e, "seed_state_predict_head", 128, 2);
    e->h_seed_type = make_head(e, "seed_type_predict_head", 128, 3);
    e->h_seed_shape = make_head(e, "seed_shape_predict_head", 128, 3);
    e->h_seed_pos = make_head(e, "seed_pos_rel_token_predict_head", 128, cfg->grid_size);
    e->h_seed_heading = make_head(e, "seed_heading_rel_token_predict_head", 128, ANGLE_SIZE);
    e->h_seed_offset = make_head(e, "seed_offset_xy_predict_head", 128, 2);
    e->h_occ_embed = make_head(e, "seed_agent_occ_embed", cfg->grid_size, 128);
    e->h_ag_occ = make_head(e, "grid_agent_occ_head", 128, cfg->grid_size);
    e->h_pt_occ = make_head(e, "grid_pt_occ_head", 128, cfg->grid_size);
    for (int i = 0; i < 3; ++i) e->mp[i] = make_attn(e, "map.pt2pt_layers." + std::to_string(i), true);
    e->f_pp = make_fourier(e, "map.r_pt2pt_emb", 3);
    e->e_map_tok = make_mlp_emb(e, "map.token_emb");
    e->h_map_tok = make_head(e, "map.token_predict_head", 128, MAP_TOKEN_SIZE);
    // kernels that need more than 48 KB of dynamic shared memory
    CK(cudaFuncSetAttribute(k_layer<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LayerSmem<4>::BYTES));
    CK(cudaFuncSetAttribute(k_layer<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LayerSmem<8>::BYTES));
    CK(cudaFuncSetAttribute(k_fourier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FOURIER_SMEM));
    CK(cudaFuncSetAttribute(k_fourier_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ftc::SMEM));
    CK(cudaFuncSetAttribute(k_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ATTN_SMEM));
    CK(cudaFuncSetAttribute(k_node<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NodeSmem::BYTES));
    CK(cudaFuncSetAttribute(k_node<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NodeSmem::BYTES));
    CK(cudaFuncSetAttribute(k_node_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntc::SMEM));
    {
        const char *ng = getenv("INFGEN_NODE_GEMM");            // "mma": 3xTF32 mma.sync tiles in k_node instead of FFMA
        e->node_mma = ng && !strcmp(ng, "mma");
        e->node_tc = !(ng && (!strcmp(ng, "mma") || !strcmp(ng, "ffma")));   // default: k_node_tc (tcgen05, 128-row tiles)
        if (getenv("INFGEN_VERBOSE")) fprintf(stderr, "infgen_b200: node_mma=%d layer_path=%d fourier_tc=%d\n", (int)e->node_mma, e->layer_path, (int)e->fourier_tc);
    }
    RET(build_node_weights(e));
    {
        const char *lp = getenv("INFGEN_LAYER_PATH");          // "cluster" | "rows": force one of the two layer paths
        e->layer_path = lp ? (!strcmp(lp, "cluster") ? 1 : (!strcmp(lp, "rows") ? 2 : 0)) : 0;
    }
    {
        const char *fm = getenv("INFGEN_FOURIER");
        e->fourier_tc = !(fm && !strcmp(fm, "ffma"));
    }
    CK(cudaFuncSetAttribute(k_mlp_embed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlp_embed_smem(128)));
    CK(cudaFuncSetAttribute(k_embed_column, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLEMB_SMEM));
    CK(cudaFuncSetAttribute(k_heads<HM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEADS_SMEM));
    CK(cudaFuncSetAttribute(k_heads<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heads_smem<16>()));
    CK(cudaFuncSetAttribute(k_mlp_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_LAYER_SMEM));
    RET(build_tables(e));
    if (!cfg->disable_insertion) RET(build_seed_feature(e));
    CK(cudaStreamSynchronize(e->stream));
    *out = e;
    return 0;
}

int32_t infgen_destroy(infgen_engine *e) {
    if (!e) return 0;
    cudaStreamSynchronize(e->stream);
    drop_graph(e);
    for (auto &r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto &kv : e->bufs)
        if (kv.second.p) cudaFree(kv.second.p);
    cudaFree(e->blob); cudaFree(e->cs_blob); cudaFree(e->grid_cells); cudaFree(e->vocab); cudaFree(e->tok_tab); cudaFree(e->grid_tab);
    cudaFree(e->d_err); cudaFree(e->seed_feat);
    for (float *p : e->wimgs) cudaFree(p);
    cudaFree(e->np_blob); cudaFree(e->vrf_blob); cudaFree(e->tc_blob); cudaFree(e->prep_buf); cudaFree(e->t_dim_table); cudaFree(e->map_tok_tab);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->side_stream) cudaStreamDestroy(e->side_stream);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->side_stream2) cudaStreamDestroy(e->side_stream2);
    if (e->ev_fork2) cudaEventDestroy(e->ev_fork2);
    if (e->ev_join2) cudaEventDestroy(e->ev_join2);
    if (e->side_stream3) cudaStreamDestroy(e->side_stream3);
    if (e->ev_fork3) cudaEventDestroy(e->ev_fork3);
    if (e->ev_join3) cudaEventDestroy(e->ev_join3);
    delete e;
    return 0;
}

int32_t infgen_set_stream(infgen_engine *e, void *cuda_stream) {
    if (!e) return fail(INFGEN_ERR_INVALID_ARG, "null engine");
    CK(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    drop_graph(e);
    return 0;
}
int32_t infgen_set_sampler(infgen_engine *e, int32_t beam, uint32_t seed) {
    if (!e) return fail(INFGEN_ERR_INVALID_ARG, "null engine");
    if (beam < 1 || beam > KTOP) return fail(INFGEN_ERR_INVALID_ARG, "motion_beam_size %d not in [1,%d]", beam, KTOP);
    if (e->cfg.motion_beam_size != beam || e->cfg.seed != seed) drop_graph(e);
    e->cfg.motion_beam_size = beam; e->cfg.seed = seed;
    e->st.beam = beam; e->st.seed = seed;
    return 0;
}
// device-side error flag (1: an attention row exceeded its edge slots, 2: the insertion stage ran out of rows) and the
// mbarrier watchdog of k_fourier_tc; the stream must be idle
static int check_device_errors(infgen_engine *e) {
    int err = 0;
    CK(cudaMemcpy(&err, e->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) {
        CK(cudaMemset(e->d_err, 0, sizeof(int)));
        if (err == 2)
            return fail(INFGEN_ERR_CAPACITY, "insertion stage ran out of rows (row_capacity %d): reload the batch with a larger "
                        "row_capacity", e->cap);
        return fail(INFGEN_ERR_CAPACITY, "device error flag %d (an attention row exceeded its edge capacity)", err);
    }
    return check_ftc_watchdog();
}
int32_t infgen_synchronize(infgen_engine *e) {
    if (!e) return fail(INFGEN_ERR_INVALID_ARG, "null engine");
    CK(cudaStreamSynchronize(e->stream));
    return e->loaded ? check_device_errors(e) : 0;
}

static cudaMemcpyKind in_kind(int loc) { return loc == INFGEN_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }
static cudaMemcpyKind out_kind(int loc) { return loc == INFGEN_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; }

int32_t infgen_load_scenes(infgen_engine *e, const infgen_scene_batch *b, int32_t loc) {
    if (!e || !b) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    if (b->n_scenes < 1 || b->row_capacity < 1 || b->row_capacity % 4 != 0 || b->row_capacity > MAX_ROW_CAPACITY)
        return fail(INFGEN_ERR_CAPACITY, "row_capacity %d must be a multiple of 4 in [4,%d]", b->row_capacity, MAX_ROW_CAPACITY);
    if ((int64_t)b->n_scenes * b->row_capacity > (int64_t)1 << 22)
        return fail(INFGEN_ERR_CAPACITY, "%d scenes x %d rows exceed the row space (4 M rows)", b->n_scenes, b->row_capacity);
    const int HC = e->cfg.hist_cols;
    if (b->n_cols < HC + b->n_iters || b->n_iters < 0)
        return fail(INFGEN_ERR_INVALID_ARG, "n_cols %d < hist_cols %d + n_iters %d", b->n_cols, HC, b->n_iters);
    const int ns = b->n_scenes, cap = b->row_capacity, R = ns * cap, T = b->n_cols, S = b->n_iters;
    int sum = 0, mx = 0;
    for (int i = 0; i < ns; ++i) {
        if (b->n_rows[i] < 1 || b->n_rows[i] > cap) return fail(INFGEN_ERR_CAPACITY, "scene %d has %d rows, capacity %d", i, b->n_rows[i], cap);
        if (b->ego_row[i] < 0 || b->ego_row[i] >= b->n_rows[i]) return fail(INFGEN_ERR_INVALID_ARG, "scene %d: ego row %d out of range", i, b->ego_row[i]);
        sum += b->n_rows[i]; mx = std::max(mx, b->n_rows[i]);
    }
    const int P = b->pt_ptr[ns];
    if (b->pt_ptr[0] != 0 || P < 0) return fail(INFGEN_ERR_INVALID_ARG, "pt_ptr must start at 0");
    if (e->n_scenes != ns || e->cap != cap || e->T != T || e->S != S || e->P != P) drop_graph(e);
    e->n_scenes = ns; e->cap = cap; e->R = R; e->T = T; e->S = S; e->P = P; e->n_rows_sum = sum; e->max_rows = mx;
    e->n_rows0.assign(b->n_rows, b->n_rows + ns);
    e->row_tile = (R + 3) / 4 <= MAX_CLUSTERS ? 4 : 8;
    e->iters_done = 0; e->prefilled = 0; e->forcing = false;
    const int W = e->cfg.window, MM = e->cfg.max_pl2a_neighbors, V = e->cfg.token_size;
    DecState &s = e->st;
    DecState old = s;
    memset(&s, 0, sizeof(s));
    s.n_scenes = ns; s.cap = cap; s.T = T; s.S = S; s.HC = HC; s.W = W; s.q_rows = e->cfg.num_seed_feature;
    s.G = e->cfg.grid_size; s.V = V; s.max_m = MM;
    s.max_a = e->cfg.max_a2a_neighbors; s.a_stride = std::min(cap, s.max_a + 1);
    s.r_m2 = e->cfg.pl2a_radius * e->cfg.pl2a_radius; s.r_a2 = e->cfg.a2a_radius * e->cfg.a2a_radius;
    s.use_state_token = e->cfg.use_state_token; s.disable_insertion = e->cfg.disable_insertion;
    s.beam = e->cfg.motion_beam_size; s.seed = e->cfg.seed;
    s.teacher_forced = e->cfg.teacher_forced;
    s.grid_cells = e->grid_cells; s.vocab = e->vocab;
    // ---- inputs ----
    int *d_n_rows, *d_ego, *d_sid, *d_type, *d_pt_ptr, *d_state_h, *d_token_h, *d_grid_h;
    float *d_pos_h, *d_head_h, *d_shape, *d_pt_pos, *d_pt_ori, *d_x_pt;
    uint8_t *d_tsrc_h, *d_int_h;
    RET(ensure_t(e, "n_rows", ns, &d_n_rows)); RET(ensure_t(e, "ego_row", ns, &d_ego)); RET(ensure_t(e, "scene_id", ns, &d_sid));
    RET(ensure_t(e, "type", R, &d_type)); RET(ensure_t(e, "pt_ptr", ns + 1, &d_pt_ptr));
    RET(ensure_t(e, "state_hist", (size_t)R * HC, &d_state_h)); RET(ensure_t(e, "token_hist", (size_t)R * HC, &d_token_h));
    RET(ensure_t(e, "grid_hist", (size_t)R * HC, &d_grid_h));
    RET(ensure_t(e, "pos_hist", (size_t)R * HC * 2, &d_pos_h)); RET(ensure_t(e, "head_hist", (size_t)R * HC, &d_head_h));
    RET(ensure_t(e, "shape", (size_t)R * 3, &d_shape));
    RET(ensure_t(e, "pt_pos", (size_t)std::max(P, 1) * 2, &d_pt_pos)); RET(ensure_t(e, "pt_ori", (size_t)std::max(P, 1), &d_pt_ori));
    RET(ensure_t(e, "x_pt", (size_t)std::max(P, 1) * 128, &d_x_pt));
    RET(ensure_t(e, "tsrc_hist", (size_t)R * HC, &d_tsrc_h)); RET(ensure_t(e, "interact_hist", (size_t)R * HC, &d_int_h));
    const cudaMemcpyKind k = in_kind(loc);
    cudaStream_t st = e->stream;
    // the small per-scene descriptors are always host memory (they are validated above)
    CK(cudaMemcpyAsync(d_n_rows, b->n_rows, ns * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_ego, b->ego_row, ns * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_sid, b->scene_id, ns * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_type, b->type, (size_t)R * sizeof(int), k, st));
    CK(cudaMemcpyAsync(d_pt_ptr, b->pt_ptr, (ns + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_state_h, b->state_hist, (size_t)R * HC * sizeof(int), k, st));
    CK(cudaMemcpyAsync(d_token_h, b->token_hist, (size_t)R * HC * sizeof(int), k, st));
    CK(cudaMemcpyAsync(d_grid_h, b->grid_hist, (size_t)R * HC * sizeof(int), k, st));
    CK(cudaMemcpyAsync(d_pos_h, b->pos_hist, (size_t)R * HC * 2 * sizeof(float), k, st));
    CK(cudaMemcpyAsync(d_head_h, b->head_hist, (size_t)R * HC * sizeof(float), k, st));
    CK(cudaMemcpyAsync(d_shape, b->shape, (size_t)R * 3 * sizeof(float), k, st));
    CK(cudaMemcpyAsync(d_tsrc_h, b->tsrc_hist, (size_t)R * HC, k, st));
    CK(cudaMemcpyAsync(d_int_h, b->interact_hist, (size_t)R * HC, k, st));
    if (P > 0) {
        CK(cudaMemcpyAsync(d_pt_pos, b->pt_pos, (size_t)P * 2 * sizeof(float), k, st));
        CK(cudaMemcpyAsync(d_pt_ori, b->pt_ori, (size_t)P * sizeof(float), k, st));
        if (b->x_pt) {
            CK(cudaMemcpyAsync(d_x_pt, b->x_pt, (size_t)P * 128 * sizeof(float), k, st));
        } else {
            // x_pt == NULL: the output of the engine's own map encoder (infgen_map_encode on the same tokens) stays in HBM
            if (e->map_P != P || !e->bufs.count("map_x"))
                return fail(INFGEN_ERR_STATE, "x_pt is NULL but infgen_map_encode has not run on these %d map tokens", P);
            CK(cudaMemcpyAsync(d_x_pt, fbuf(e, "map_x"), (size_t)P * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
    }
    s.n_rows = d_n_rows; s.ego_row = d_ego; s.scene_id = d_sid; s.type = d_type; s.pt_ptr = d_pt_ptr;
    s.pt_pos = d_pt_pos; s.pt_ori = d_pt_ori;
    // ---- state ----
    RET(ensure_t(e, "col", 2, &s.col)); s.iter = s.col + 1;
    RET(ensure_t(e, "pos", (size_t)R * T * 2, &s.pos)); RET(ensure_t(e, "head", (size_t)R * T, &s.head));
    RET(ensure_t(e, "state", (size_t)R * T, &s.state)); RET(ensure_t(e, "token", (size_t)R * T, &s.token));
    RET(ensure_t(e, "grid", (size_t)R * T, &s.grid));
    RET(ensure_t(e, "interact", (size_t)R * T, &s.interact)); RET(ensure_t(e, "tsrc", (size_t)R * T, &s.tsrc));
    RET(ensure_t(e, "t_cnt", R, &s.t_cnt)); RET(ensure_t(e, "t_src", (size_t)R * W, &s.t_src));
    RET(ensure_t(e, "t_raw", (size_t)R * W * 4, &s.t_raw));
    RET(ensure_t(e, "m_cnt", R, &s.m_cnt)); RET(ensure_t(e, "m_src", (size_t)R * MM, &s.m_src));
    RET(ensure_t(e, "m_raw", (size_t)R * MM * 3, &s.m_raw));
    RET(ensure_t(e, "a_cnt", R, &s.a_cnt)); RET(ensure_t(e, "a_start", R, &s.a_start)); RET(ensure_t(e, "a_total", ns, &s.a_total));
    RET(ensure_t(e, "a_src", (size_t)R * s.a_stride, &s.a_src)); RET(ensure_t(e, "a_raw", (size_t)R * s.a_stride * 3, &s.a_raw));
    RET(ensure_t(e, "xa_raw", (size_t)R * 2, &s.xa_raw));
    RET(ensure_t(e, "tok_row", R, &s.tok_row)); RET(ensure_t(e, "state_idx", R, &s.state_idx));
    RET(ensure_t(e, "grid_row", R, &s.grid_row)); RET(ensure_t(e, "cat_idx", R, &s.cat_idx));
    float *part_v, *part_m, *part_s, *slog; int *part_i;
    RET(ensure_t(e, "part_v", (size_t)R * NSLICE * KTOP, &part_v)); RET(ensure_t(e, "part_i", (size_t)R * NSLICE * KTOP, &part_i));
    RET(ensure_t(e, "part_m", (size_t)R * NSLICE, &part_m)); RET(ensure_t(e, "part_s", (size_t)R * NSLICE, &part_s));
    RET(ensure_t(e, "state_logits", (size_t)R * 4, &slog));
    s.part_v = part_v; s.part_i = part_i; s.part_m = part_m; s.part_s = part_s; s.state_logits = slog;
    const int NR = std::max(5 * S, 1);
    RET(ensure_t(e, "pred_traj", (size_t)R * NR * 2, &s.pred_traj)); RET(ensure_t(e, "pred_head", (size_t)R * NR, &s.pred_head));
    RET(ensure_t(e, "pred_state", (size_t)R * NR, &s.pred_state));
    RET(ensure_t(e, "next_token", (size_t)R * T, &s.next_token)); RET(ensure_t(e, "next_state", (size_t)R * T, &s.next_state));
    // ---- scratch ----
    float *tmp;
    RET(ensure_t(e, "x", (size_t)R * 128, &tmp)); RET(ensure_t(e, "q", (size_t)R * 128, &tmp)); RET(ensure_t(e, "s", (size_t)R * 128, &tmp));
    RET(ensure_t(e, "qr", (size_t)R * 1024, &tmp)); RET(ensure_t(e, "agg", (size_t)R * 128, &tmp));
    RET(ensure_t(e, "zero", (size_t)R * 1024, &tmp)); RET(ensure_t(e, "xa", (size_t)R * 128, &tmp));
    RET(ensure_t(e, "kv_t", (size_t)6 * R * RING * 256, &tmp)); RET(ensure_t(e, "kv_a", (size_t)2 * R * 256, &tmp));
    { unsigned *gb; RET(ensure_t(e, "grid_bar", 4, &gb)); }
    RET(ensure_t(e, "kv_m", (size_t)6 * std::max(P, 1) * 256, &tmp));
    RET(ensure_t(e, "rhat_t", (size_t)R * W * 128, &tmp)); RET(ensure_t(e, "rhat_m", (size_t)R * MM * 128, &tmp));
    RET(ensure_t(e, "rhat_a", (size_t)R * s.a_stride * 128, &tmp));
    { int *itmp; RET(ensure_t(e, "a_slots", (size_t)R * s.a_stride + 4 + R, &itmp)); }   // slot list, count, row offsets
    RET(ensure_t(e, "cat_tab", (size_t)(R + 1) * 128, &tmp)); RET(ensure_t(e, "shape_rows", (size_t)(R + 1) * 4, &tmp));
    RET(ensure_t(e, "hist_traj", (size_t)R * HC * 5 * 2, &tmp)); RET(ensure_t(e, "hist_head", (size_t)R * HC * 5, &tmp));
    if (getenv("INFGEN_TSTAMP")) { long long *ts; RET(ensure_t(e, "tstamp", 512, &ts)); }
    if (e->cfg.trace && S > 0) {
        RET(ensure_t(e, "trace_head_in", (size_t)S * R * 128, &tmp));
        RET(ensure_t(e, "trace_token_logits", (size_t)S * R * V, &tmp));
        RET(ensure_t(e, "trace_state_logits", (size_t)S * R * 3, &tmp));
        RET(ensure_t(e, "trace_layer_out", (size_t)S * 6 * R * 128, &tmp));
    }
    // ---- insertion stage ----
    e->ins_ready = false;
    s.ins_col = nullptr; s.hv_src = nullptr;
    if (!e->cfg.disable_insertion) {
        InsState &q = e->ins;
        memset(&q, 0, sizeof(q));
        const int G = e->cfg.grid_size;
        q.beam = e->cfg.insert_beam_size; q.force_enter = e->cfg.debug_force_enter; q.seed = e->cfg.seed;
        q.r_seed2 = e->cfg.pl2seed_radius * e->cfg.pl2seed_radius;
        q.r_new_a2 = e->cfg.a2sa_radius * e->cfg.a2sa_radius; q.r_new_m2 = e->cfg.pl2sa_radius * e->cfg.pl2sa_radius;
        q.angle_interval = e->cfg.angle_interval;
        RET(ensure_t(e, "ins_active", ns, &q.active)); RET(ensure_t(e, "ins_n_new", ns, &q.n_new));
        RET(ensure_t(e, "ins_pass", ns, &q.pass)); RET(ensure_t(e, "ins_new_row", ns, &q.new_row));
        RET(ensure_t(e, "ins_row_lo", ns, &q.row_lo)); RET(ensure_t(e, "ins_flags", 4, &q.flags));
        RET(ensure_t(e, "ins_done_ctr", 4, &q.done_ctr)); RET(ensure_t(e, "ins_ha_lo", ns, &q.ha_lo));
        RET(ensure_t(e, "ins_stat", 4, &q.stat));
        RET(ensure_t(e, "as_seen", ns, &q.as_seen)); RET(ensure_t(e, "as_new_list", ns + 4, &q.as_new_list));
        q.as_new_n = q.as_new_list + ns;
        RET(ensure_t(e, "ins_new_list", ns + 4, &q.new_list)); q.n_new_list = q.new_list + ns;
        RET(ensure_t(e, "ins_prev_list", ns + 4, &q.prev_list)); q.n_prev_list = q.prev_list + ns;
        q.as_stride = std::min(cap, SEED_AGENT_MAX);
        RET(ensure_t(e, "ps_cnt", ns, &q.ps_cnt)); RET(ensure_t(e, "ps_src", (size_t)ns * SEED_MAP_MAX, &q.ps_src));
        RET(ensure_t(e, "ps_raw", (size_t)ns * SEED_MAP_MAX * 3, &q.ps_raw));
        RET(ensure_t(e, "as_cnt", ns, &q.as_cnt)); RET(ensure_t(e, "as_src", (size_t)ns * q.as_stride, &q.as_src));
        RET(ensure_t(e, "as_raw", (size_t)ns * q.as_stride * 3, &q.as_raw));
        RET(ensure_t(e, "one_cnt", ns, &q.one_cnt)); RET(ensure_t(e, "occ_src", ns, &q.occ_src));
        RET(ensure_t(e, "hp_cnt", R, &q.hp_cnt)); RET(ensure_t(e, "hp_start", R, &q.hp_start));
        RET(ensure_t(e, "hp_src", (size_t)ns * NEW_MAP_MAX, &q.hp_src)); RET(ensure_t(e, "hp_raw", (size_t)ns * NEW_MAP_MAX * 3, &q.hp_raw));
        RET(ensure_t(e, "ha_cnt", R, &q.ha_cnt)); RET(ensure_t(e, "ha_start", R, &q.ha_start));
        RET(ensure_t(e, "ha_src", (size_t)ns * NEW_AGENT_MAX, &q.ha_src)); RET(ensure_t(e, "ha_raw", (size_t)ns * NEW_AGENT_MAX * 3, &q.ha_raw));
        RET(ensure_t(e, "hp_cnt_s", ns, &q.hp_cnt_s)); RET(ensure_t(e, "ha_cnt_s", ns, &q.ha_cnt_s));
        RET(ensure_t(e, "occ", (size_t)ns * G, &q.occ)); RET(ensure_t(e, "occ_emb", (size_t)ns * 128, &q.occ_emb));
        RET(ensure_t(e, "kv_occ", (size_t)3 * ns * 256, &q.kv_occ)); RET(ensure_t(e, "x_seed", (size_t)ns * SEED_ROW_STRIDE * 128, &q.x_seed));
        RET(ensure_t(e, "pos_logits", (size_t)ns * SEED_ROW_STRIDE * G, &q.pos_logits));
        RET(ensure_t(e, "ag_occ_logits", (size_t)ns * SEED_ROW_STRIDE * G, &q.ag_occ_logits));
        RET(ensure_t(e, "pt_occ_logits", (size_t)ns * SEED_ROW_STRIDE * G, &q.pt_occ_logits));
        RET(ensure_t(e, "ins_col", R, &q.ins_col)); RET(ensure_t(e, "pred_type", R, &q.pred_type));
        RET(ensure_t(e, "pred_shape", (size_t)R * 3, &q.pred_shape));
        RET(ensure_t(e, "rec_meta", (size_t)R * 2, &q.rec_meta)); RET(ensure_t(e, "rec_state_prob", (size_t)R, &q.rec_state_prob));
        RET(ensure_t(e, "rec_pos_prob", (size_t)R * G, &q.rec_pos_prob)); RET(ensure_t(e, "rec_ag_occ", (size_t)R * G, &q.rec_ag_occ));
        RET(ensure_t(e, "rec_pt_occ", (size_t)R * G, &q.rec_pt_occ)); RET(ensure_t(e, "rec_occ_gt", (size_t)R * G, &q.rec_occ_gt));
        RET(ensure_t(e, "rec_softmax", (size_t)ns * 2, &q.rec_softmax));
        RET(ensure_t(e, "small_logits", (size_t)ns * SEED_ROW_STRIDE * 8, &q.small_logits));
        RET(ensure_t(e, "x_sa", (size_t)R * 128, &tmp)); RET(ensure_t(e, "x_ha", (size_t)R * 128, &tmp));
        RET(ensure_t(e, "s_ha", (size_t)R * 128, &tmp));
        RET(ensure_t(e, "kv_sa", (size_t)3 * R * 256, &tmp)); RET(ensure_t(e, "kv_ha", (size_t)3 * R * 256, &tmp));
        RET(ensure_t(e, "kv_ms", (size_t)3 * std::max(P, 1) * 256, &tmp));
        RET(ensure_t(e, "rhat_ps", (size_t)ns * SEED_MAP_MAX * 128, &tmp)); RET(ensure_t(e, "rhat_as", (size_t)ns * q.as_stride * 128, &tmp));
        RET(ensure_t(e, "rhat_hp", (size_t)ns * NEW_MAP_MAX * 128, &tmp)); RET(ensure_t(e, "rhat_ha", (size_t)ns * NEW_AGENT_MAX * 128, &tmp));
        q.shape_rows = fbuf(e, "shape_rows");
        q.seed_feat = e->seed_feat; q.err = e->d_err;
        // query rows: one per tile (all warps share its edges) while every scene's cluster fits one wave, packed otherwise
        // (INFGEN_SEED_WIDE_MAX: experiments with the switch-over point; default = one wave of clusters)
        const int wide_max = getenv("INFGEN_SEED_WIDE_MAX") ? atoi(getenv("INFGEN_SEED_WIDE_MAX")) : MAX_CLUSTERS;
        q.seed_stride = (ns <= wide_max || getenv("INFGEN_SEED_WIDE")) ? SEED_ROW_STRIDE : 1;
        e->ins_ride = q.seed_stride == SEED_ROW_STRIDE && !getenv("INFGEN_NO_RIDE");
        s.ins_col = q.ins_col;
        RET(ensure_t(e, "hv_src", R, &s.hv_src));
        CK(cudaMemsetAsync(s.hv_src, 0xff, (size_t)R * sizeof(int), e->stream));
        CK(cudaMemsetAsync(q.ins_col, 0xff, (size_t)R * sizeof(int), e->stream));
        CK(cudaMemsetAsync(q.pred_type, 0, (size_t)R * sizeof(int), e->stream));
        CK(cudaMemsetAsync(q.pred_shape, 0, (size_t)R * 3 * sizeof(float), e->stream));
        e->ins_ready = true;
    }
    if (memcmp(&old, &s, sizeof(s)) != 0) drop_graph(e);
    // ---- expand history into the state arrays, zero counters ----
    SetupArgs sa;
    sa.s = s; sa.pos_hist = d_pos_h; sa.head_hist = d_head_h; sa.state_hist = d_state_h; sa.token_hist = d_token_h;
    sa.grid_hist = d_grid_h; sa.tsrc_hist = d_tsrc_h; sa.interact_hist = d_int_h;
    k_setup_state<<<(R * T + 255) / 256, 256, 0, st>>>(sa);
    CKL(); count_launch(e);
    CK(cudaMemsetAsync(s.a_total, 0, ns * sizeof(int), st));
    CK(cudaMemsetAsync(e->d_err, 0, sizeof(int), st));
    CK(cudaMemsetAsync(s.col, 0, 2 * sizeof(int), st));
    // ---- categorical embedding rows (type + shape, agent_decoder.py:449-478) ----
    float *shape_rows = fbuf(e, "shape_rows"), *cat_tab = fbuf(e, "cat_tab");
    k_fill_shape_rows<<<((R + 1) * 3 + 255) / 256, 256, 0, st>>>(shape_rows, d_shape, R);
    CKL(); count_launch(e);
    MlpEmbArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.rows = flat_rows(R + 1); ma.w = e->e_shape; ma.kin = 3; ma.k4 = 1; ma.x = shape_rows; ma.x_ld = 3;
    ma.out = cat_tab; ma.out_ld = 128;
    RET(launch_mlp_embed(e, ma));
    k_add_type_emb<<<((R + 1) * 128 + 255) / 256, 256, 0, st>>>(cat_tab, e->type_emb, d_type, R, 3);
    CKL(); count_launch(e);
    // ---- map K/V cache of the six pt2a layers (x_pt is constant during the rollout) ----
    if (P > 0) {
        KvArgs ka;
        memset(&ka, 0, sizeof(ka));
        ka.n = P; ka.x = d_x_pt;
        for (int i = 0; i < 6; ++i) { ka.w[i] = e->m[i]; ka.out[i] = fbuf(e, "kv_m") + (size_t)i * P * 256; }
        k_kv_project<16><<<dim3((P + 15) / 16, 6), NT, 0, st>>>(ka);
        CKL(); count_launch(e);
        if (e->ins_ready) {                              // ... and of the three pt2sa layers of the insertion stage
            memset(&ka, 0, sizeof(ka));
            ka.n = P; ka.x = d_x_pt;
            for (int i = 0; i < 3; ++i) { ka.w[i] = e->pt2sa[i]; ka.out[i] = fbuf(e, "kv_ms") + (size_t)i * P * 256; }
            k_kv_project<16><<<dim3((P + 15) / 16, 3), NT, 0, st>>>(ka);
            CKL(); count_launch(e);
        }
    }
    e->loaded = true;
    return 0;
}

int32_t infgen_set_forcing(infgen_engine *e, const int32_t *tokens, const int32_t *states, int32_t loc) {
    if (!e || !e->loaded) return fail(INFGEN_ERR_STATE, "no scenes loaded");
    const size_t n = (size_t)e->R * std::max(e->S, 1);
    const int *old_t = e->st.forced_tok, *old_s = e->st.forced_state;
    e->st.forced_tok = nullptr; e->st.forced_state = nullptr;
    if (tokens) {
        int *d; RET(ensure_t(e, "forced_tok", n, &d));
        CK(cudaMemcpyAsync(d, tokens, n * sizeof(int), in_kind(loc), e->stream));
        e->st.forced_tok = d;
    }
    if (states) {
        int *d; RET(ensure_t(e, "forced_state", n, &d));
        CK(cudaMemcpyAsync(d, states, n * sizeof(int), in_kind(loc), e->stream));
        e->st.forced_state = d;
    }
    if (old_t != e->st.forced_tok || old_s != e->st.forced_state) drop_graph(e);
    return 0;
}

int32_t infgen_prefill(infgen_engine *e) {
    if (!e || !e->loaded) return fail(INFGEN_ERR_STATE, "no scenes loaded");
    if (e->prefilled) return fail(INFGEN_ERR_STATE, "prefill already done for this batch");
    DecState &s = e->st;
    CK(cudaMemsetAsync(s.col, 0, 2 * sizeof(int), e->stream));
    for (int c = 0; c + 1 < s.HC; ++c) {
        RET(enqueue_embed_column(e, 0));
        RET(enqueue_layers(e, false, -1));
        k_set_scalar<<<1, 1, 0, e->stream>>>(s.col, c + 1);
        CKL(); count_launch(e);
    }
    RET(enqueue_embed_column(e, 0));
    if (e->early_edges) RET(enqueue_edges(e, 0));
    e->prefilled = 1;
    return 0;
}

// capture one decode iteration (which = 1: preceded by the insertion stage) into e->graph[which]
static int capture_iteration(infgen_engine *e, int which) {
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    e->capturing = true;
    int rc = 0;
    if (which == 1) {
        cudaStreamCaptureStatus status;
        cudaGraph_t top = nullptr;
        cudaError_t ce = cudaStreamGetCaptureInfo(e->stream, &status, nullptr, &top, nullptr, nullptr);
        if (ce == cudaSuccess) ce = cudaGraphConditionalHandleCreate(&e->ins.h_pass, top, 0, cudaGraphCondAssignDefault);
        if (ce == cudaSuccess) ce = cudaGraphConditionalHandleCreate(&e->ins.h_new, top, 0, cudaGraphCondAssignDefault);
        if (ce != cudaSuccess) rc = fail(INFGEN_ERR_CUDA, "conditional handles: %s", cudaGetErrorString(ce));
        e->ins.use_cond = 1;
        if (rc == 0) rc = enqueue_insertion_graph(e);
    }
    if (rc == 0) rc = enqueue_iteration(e, -1);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
    if (rc == 0 && ce != cudaSuccess) rc = fail(INFGEN_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    if (rc == 0 && which == 1) {
        // bodies of the conditional nodes: the pass (WHILE) and, nested in it, the heading stage (IF)
        std::vector<size_t> counts;
        rc = drain_cond_bodies(e, counts);
        if (rc == 0 && counts.size() == 2) {
            e->pass_nodes = counts[0] - 1;               // minus the IF node
            e->heading_nodes = counts[1];
        }
    } else {
        g_cond_bodies.clear();
    }
    e->capturing = false;
    e->ins.use_cond = 0;
    if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
    e->graph[which] = g;
    CK(cudaGraphGetNodes(g, nullptr, &e->graph_nodes[which]));
    CK(cudaGraphInstantiate(&e->graph_exec[which], g, 0));
    return 0;
}

// Motion branch of the teacher-forced pass (agent_decoder.py:1104-1240): every column is embedded from the given token stream
// and run through the stack as the destination column, in order, so that the temporal K/V of the earlier columns are in the
// ring exactly as a closed-loop rollout would have left them (row a15 / f4).
int32_t infgen_forward(infgen_engine *e, float *x_a, float *token_logits, float *state_logits, int32_t loc) {
    if (!e || !e->loaded) return fail(INFGEN_ERR_STATE, "no scenes loaded");
    if (!e->cfg.teacher_forced || e->st.HC != e->T || e->cfg.num_seed_feature != 0)
        return fail(INFGEN_ERR_STATE, "infgen_forward needs an engine created with teacher_forced = 1, hist_cols == n_cols (%d vs %d) "
                    "and num_seed_feature = 0", e->st.HC, e->T);
    if (e->prefilled) return fail(INFGEN_ERR_STATE, "the batch has already been run");
    DecState &s = e->st;
    const int R = e->R, T = e->T, V = e->cfg.token_size;
    float *d_x = nullptr, *d_tok = nullptr, *d_st = nullptr;
    RET(ensure_t(e, "fwd_x", (size_t)T * R * 128, &d_x));
    RET(ensure_t(e, "fwd_logits", (size_t)T * R * V, &d_tok));
    RET(ensure_t(e, "fwd_state", (size_t)T * R * 3, &d_st));
    CK(cudaMemsetAsync(s.col, 0, 2 * sizeof(int), e->stream));
    for (int c = 0; c < T; ++c) {
        RET(enqueue_embed_column(e, 0));
        RET(enqueue_edges(e, 0));
        RET(enqueue_layers(e, true, -1));
        HeadArgs ha;
        memset(&ha, 0, sizeof(ha));
        ha.rows = scene_rows(e); ha.x = fbuf(e, "x"); ha.tok = e->h_tok; ha.st = e->h_state;
        ha.part_v = fbuf(e, "part_v"); ha.part_i = (int *)e->bufs["part_i"].p;
        ha.part_m = fbuf(e, "part_m"); ha.part_s = fbuf(e, "part_s"); ha.state_logits = fbuf(e, "state_logits");
        ha.trace_head_in = d_x + (size_t)c * R * 128;
        ha.trace_logits = d_tok + (size_t)c * R * V;
        ha.trace_state = d_st + (size_t)c * R * 3;
        {
            ProfScope ps(e, KC_HEADS);
            if (R > 512) k_heads<16><<<dim3((R + 15) / 16, NSLICE + 1), NT_S, heads_smem<16>(), e->stream>>>(ha);
            else k_heads<HM><<<dim3((R + HM - 1) / HM, NSLICE + 1), NT_S, HEADS_SMEM, e->stream>>>(ha);
        }
        CKL(); count_launch(e);
        k_set_scalar<<<1, 1, 0, e->stream>>>(s.col, c + 1);
        CKL(); count_launch(e);
    }
    e->prefilled = 1;
    const cudaMemcpyKind k = loc == INFGEN_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (x_a) CK(cudaMemcpyAsync(x_a, d_x, (size_t)T * R * 128 * sizeof(float), k, e->stream));
    if (token_logits) CK(cudaMemcpyAsync(token_logits, d_tok, (size_t)T * R * V * sizeof(float), k, e->stream));
    if (state_logits) CK(cudaMemcpyAsync(state_logits, d_st, (size_t)T * R * 3 * sizeof(float), k, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    RET(check_device_errors(e));
    return 0;
}

int32_t infgen_step(infgen_engine *e, int32_t n_iters) {
    if (!e || !e->loaded) return fail(INFGEN_ERR_STATE, "no scenes loaded");
    if (!e->prefilled) return fail(INFGEN_ERR_STATE, "infgen_prefill has not run");
    if (n_iters < 0 || e->iters_done + n_iters > e->S)
        return fail(INFGEN_ERR_INVALID_ARG, "%d iterations requested, %d of %d already done", n_iters, e->iters_done, e->S);
    const bool use_graph = e->cfg.use_cuda_graph && !e->cfg.trace && !e->profile;
    for (int i = 0; i < n_iters; ++i) {
        // the insertion stage runs from the second iteration on (agent_decoder.py:1773)
        const bool ins = e->ins_ready && e->iters_done > 0 && !e->forcing_no_insert;
        if (use_graph) {
            const int which = ins ? 1 : 0;
            if (!e->graph_exec[which]) RET(capture_iteration(e, which));
            CK(cudaGraphLaunch(e->graph_exec[which], e->stream));
            e->launches += (int64_t)e->graph_nodes[which] - (which == 1 ? 1 : 0);
            if (which == 1) e->ins_replays++;
        } else {
            if (ins) RET(run_insertion(e));
            RET(enqueue_iteration(e, e->iters_done));
        }
        e->iters_done++;
    }
    return 0;
}

int32_t infgen_rollout(infgen_engine *e) {
    if (!e || !e->loaded) return fail(INFGEN_ERR_STATE, "no scenes loaded");
    if (!e->prefilled) RET(infgen_prefill(e));
    return infgen_step(e, e->S - e->iters_done);
}

int32_t infgen_iterations_done(infgen_engine *e) { return e ? e->iters_done : -1; }
int64_t infgen_kernel_launches(infgen_engine *e) {
    if (!e) return -1;
    if (e->ins_replays > 0 && e->ins.stat) {
        // launches inside the conditional bodies of the replayed insertion stage: counted on the device (passes, heading
        // stages), folded in here
        int h[2] = {0, 0};
        if (cudaStreamSynchronize(e->stream) == cudaSuccess &&
            cudaMemcpy(h, e->ins.stat, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
            e->launches += (int64_t)h[0] * (int64_t)e->pass_nodes + (int64_t)h[1] * (int64_t)e->heading_nodes;
            cudaMemset(e->ins.stat, 0, sizeof(h));
        }
        e->ins_replays = 0;
    }
    return e->launches;
}

int32_t infgen_read(infgen_engine *e, const infgen_outputs *o, int32_t loc) {
    if (!e || !e->loaded || !o) return fail(INFGEN_ERR_STATE, "no scenes loaded");
    DecState &s = e->st;
    const size_t R = e->R, T = e->T, NR = (size_t)5 * e->S, HC = s.HC;
    const cudaMemcpyKind k = out_kind(loc);
    cudaStream_t st = e->stream;
    if (o->hist_traj || o->hist_head) {
        k_history_traj<<<((int)R + 127) / 128, 128, 0, st>>>(s, fbuf(e, "hist_traj"), fbuf(e, "hist_head"));
        CKL(); count_launch(e);
    }
    if (o->pos) CK(cudaMemcpyAsync(o->pos, s.pos, R * T * 2 * sizeof(float), k, st));
    if (o->head) CK(cudaMemcpyAsync(o->head, s.head, R * T * sizeof(float), k, st));
    if (o->pred_traj && NR) CK(cudaMemcpyAsync(o->pred_traj, s.pred_traj, R * NR * 2 * sizeof(float), k, st));
    if (o->pred_head && NR) CK(cudaMemcpyAsync(o->pred_head, s.pred_head, R * NR * sizeof(float), k, st));
    if (o->pred_state && NR) CK(cudaMemcpyAsync(o->pred_state, s.pred_state, R * NR * sizeof(float), k, st));
    if (o->next_token) CK(cudaMemcpyAsync(o->next_token, s.next_token, R * T * sizeof(int), k, st));
    if (o->next_state) CK(cudaMemcpyAsync(o->next_state, s.next_state, R * T * sizeof(int), k, st));
    if (o->hist_traj) CK(cudaMemcpyAsync(o->hist_traj, fbuf(e, "hist_traj"), R * HC * 5 * 2 * sizeof(float), k, st));
    if (o->hist_head) CK(cudaMemcpyAsync(o->hist_head, fbuf(e, "hist_head"), R * HC * 5 * sizeof(float), k, st));
    const size_t ns = e->n_scenes;
    if (o->n_rows_final) CK(cudaMemcpyAsync(o->n_rows_final, s.n_rows, ns * sizeof(int), k, st));
    if (e->ins_ready) {
        const InsState &q = e->ins;
        const size_t G = e->cfg.grid_size;
        if (o->pred_type) CK(cudaMemcpyAsync(o->pred_type, q.pred_type, R * sizeof(int), k, st));
        if (o->pred_shape) CK(cudaMemcpyAsync(o->pred_shape, q.pred_shape, R * 3 * sizeof(float), k, st));
        // insertion records: only the rows each scene appended travel (include/infgen_b200.h, "wire format")
        if (o->rec_meta || o->rec_state_prob || o->rec_pos_prob || o->rec_agent_occ || o->rec_pt_occ || o->rec_occ_gt) {
            std::vector<int> nf(ns);
            CK(cudaMemcpyAsync(nf.data(), s.n_rows, ns * sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (size_t b = 0; b < ns; ++b) {
                const size_t r0 = b * e->cap + e->n_rows0[b], n = (size_t)std::max(0, nf[b] - e->n_rows0[b]);
                if (!n) continue;
                if (o->rec_meta) CK(cudaMemcpyAsync(o->rec_meta + r0 * 2, q.rec_meta + r0 * 2, n * 2 * sizeof(int), k, st));
                if (o->rec_state_prob) CK(cudaMemcpyAsync(o->rec_state_prob + r0, q.rec_state_prob + r0, n * sizeof(float), k, st));
                if (o->rec_pos_prob) CK(cudaMemcpyAsync(o->rec_pos_prob + r0 * G, q.rec_pos_prob + r0 * G, n * G * sizeof(float), k, st));
                if (o->rec_agent_occ) CK(cudaMemcpyAsync(o->rec_agent_occ + r0 * G, q.rec_ag_occ + r0 * G, n * G * sizeof(float), k, st));
                if (o->rec_pt_occ) CK(cudaMemcpyAsync(o->rec_pt_occ + r0 * G, q.rec_pt_occ + r0 * G, n * G * sizeof(float), k, st));
                if (o->rec_occ_gt) CK(cudaMemcpyAsync(o->rec_occ_gt + r0 * G, q.rec_occ_gt + r0 * G, n * G * sizeof(float), k, st));
            }
        }
    }
    // results in host memory are complete when this returns, and so are the error checks; device-resident results are
    // asynchronous: errors of that rollout surface at infgen_synchronize
    if (loc == INFGEN_HOST) {
        CK(cudaStreamSynchronize(st));
        RET(check_device_errors(e));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// map encoder: InfGenMapDecoder.forward (map_decoder.py:70-130)
// ---------------------------------------------------------------------------------------------------------------
int32_t infgen_map_setup(infgen_engine *e, const float *traj_src, int32_t n_tokens) {
    if (!e || !traj_src || n_tokens < 1) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    cudaStream_t st = e->stream;
    CK(cudaStreamSynchronize(st));
    if (e->map_tok_tab) { CK(cudaFree(e->map_tok_tab)); e->map_tok_tab = nullptr; }
    float *d_src = nullptr;
    CK(cudaMalloc(&d_src, (size_t)n_tokens * MAP_TOKEN_DIM * sizeof(float)));
    CK(cudaMalloc(&e->map_tok_tab, (size_t)n_tokens * 128 * sizeof(float)));
    CK(cudaMemcpyAsync(d_src, traj_src, (size_t)n_tokens * MAP_TOKEN_DIM * sizeof(float), cudaMemcpyHostToDevice, st));
    MlpEmbArgs ma;                            // token_emb over the whole vocabulary (map_decoder.py:79-80)
    memset(&ma, 0, sizeof(ma));
    ma.rows = flat_rows(n_tokens); ma.w = e->e_map_tok; ma.kin = MAP_TOKEN_DIM; ma.k4 = (MAP_TOKEN_DIM + 3) / 4;
    ma.x = d_src; ma.x_ld = MAP_TOKEN_DIM; ma.out = e->map_tok_tab; ma.out_ld = 128;
    RET(launch_mlp_embed(e, ma));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_src));
    e->map_n_tokens = n_tokens;
    return 0;
}

int32_t infgen_map_encode(infgen_engine *e, const infgen_map_batch *b, int32_t loc, float *x_pt_out, float *logits_out) {
    if (!e || !b) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    if (!e->map_tok_tab) return fail(INFGEN_ERR_STATE, "infgen_map_setup has not run");
    const int ns = b->n_scenes;
    if (ns < 1 || b->pt_ptr[0] != 0) return fail(INFGEN_ERR_INVALID_ARG, "pt_ptr must start at 0");
    const int P = b->pt_ptr[ns];
    if (P < 1) return fail(INFGEN_ERR_INVALID_ARG, "no map tokens");
    cudaStream_t st = e->stream;
    const cudaMemcpyKind k = in_kind(loc);
    MapState m;
    memset(&m, 0, sizeof(m));
    m.n_scenes = ns; m.P = P; m.r2 = b->pl2pl_radius * b->pl2pl_radius;
    int *d_ptr, *d_type, *d_pl, *d_light, *d_tok, *d_scene;
    float *d_pos, *d_ori, *x, *tmp;
    RET(ensure_t(e, "map_pt_ptr", ns + 1, &d_ptr)); RET(ensure_t(e, "map_type", P, &d_type)); RET(ensure_t(e, "map_pl_type", P, &d_pl));
    RET(ensure_t(e, "map_light", P, &d_light)); RET(ensure_t(e, "map_token_idx", P, &d_tok)); RET(ensure_t(e, "map_scene_of", P, &d_scene));
    RET(ensure_t(e, "map_pos", (size_t)P * 2, &d_pos)); RET(ensure_t(e, "map_ori", P, &d_ori));
    RET(ensure_t(e, "map_x", (size_t)P * 128, &x));
    RET(ensure_t(e, "map_q", (size_t)P * 128, &tmp)); RET(ensure_t(e, "map_s", (size_t)P * 128, &tmp));
    RET(ensure_t(e, "map_qr", (size_t)P * 1024, &tmp)); RET(ensure_t(e, "map_agg", (size_t)P * 128, &tmp));
    RET(ensure_t(e, "map_kv", (size_t)P * 256, &tmp));
    RET(ensure_t(e, "map_cnt", P, &m.cnt)); RET(ensure_t(e, "map_start", P, &m.start));
    RET(ensure_t(e, "map_src", (size_t)P * MAP_STRIDE, &m.src)); RET(ensure_t(e, "map_raw", (size_t)P * MAP_STRIDE * 3, &m.raw));
    RET(ensure_t(e, "map_rhat", (size_t)P * MAP_STRIDE * 128, &tmp));
    { int *itmp; RET(ensure_t(e, "map_slots", (size_t)P * MAP_STRIDE + 4 + P, &itmp)); }
    for (int i = 0; i < ns; ++i)
        if (b->pt_ptr[i + 1] < b->pt_ptr[i]) return fail(INFGEN_ERR_INVALID_ARG, "pt_ptr must be non-decreasing");
    CK(cudaMemcpyAsync(d_ptr, b->pt_ptr, (ns + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_pos, b->pt_pos, (size_t)P * 2 * sizeof(float), k, st));
    CK(cudaMemcpyAsync(d_ori, b->pt_ori, (size_t)P * sizeof(float), k, st));
    CK(cudaMemcpyAsync(d_type, b->type, (size_t)P * sizeof(int), k, st));
    CK(cudaMemcpyAsync(d_pl, b->pl_type, (size_t)P * sizeof(int), k, st));
    CK(cudaMemcpyAsync(d_light, b->light_type, (size_t)P * sizeof(int), k, st));
    CK(cudaMemcpyAsync(d_tok, b->token_idx, (size_t)P * sizeof(int), k, st));
    m.pt_ptr = d_ptr; m.pos = d_pos; m.ori = d_ori; m.type = d_type; m.pl_type = d_pl; m.light_type = d_light;
    m.token_idx = d_tok; m.scene_of = d_scene;
    k_map_scene_of<<<ns, 256, 0, st>>>(d_ptr, ns, d_scene);
    CKL(); count_launch(e);
    k_map_embed<<<(P + NWARP - 1) / NWARP, NT, 0, st>>>(m, e->map_tok_tab, W(e, "map.type_pt_emb"), W(e, "map.polygon_type_emb"),
                                                       W(e, "map.light_pl_emb"), x);
    CKL(); count_launch(e);
    k_map_edges<<<(P + NWARP - 1) / NWARP, NT, 0, st>>>(m);
    CKL(); count_launch(e);
    // relative embedding of the edges (map_decoder.py:98-114), tensor-core path over a compact list of the valid slots
    FourierArgs fj;
    memset(&fj, 0, sizeof(fj));
    fj.normalize = 1; fj.dim = 3; fj.n_slots = P * MAP_STRIDE; fj.cnt = m.cnt; fj.stride = MAP_STRIDE;
    fj.raw = m.raw; fj.w = e->f_pp; fj.out = fbuf(e, "map_rhat");
    if (e->fourier_tc) {
        int *list = (int *)e->bufs["map_slots"].p, *n_list = list + (size_t)P * MAP_STRIDE;
        k_slot_compact<<<1, 1024, 0, st>>>(m.cnt, P, MAP_STRIDE, list, n_list, n_list + 4);
        CKL(); count_launch(e);
        fj.slot_list = list; fj.n_list = n_list;
    }
    RET(launch_fourier(e, &fj, 1, KC_MISC));
    // three pt2pt AttentionLayers (non-bipartite, :115-117) on the row-tile path
    const RowSpace rows = flat_rows(P);
    NodeBufs nb{x, fbuf(e, "map_q"), fbuf(e, "map_s"), fbuf(e, "map_qr"), fbuf(e, "map_agg")};
    float *kv = fbuf(e, "map_kv");
    RET(launch_node(e, rows, nullptr, &e->mp[0], true, kv, false, nullptr, &nb));
    for (int i = 0; i < 3; ++i) {
        SubArgs g;
        memset(&g, 0, sizeof(g));
        g.has_attn = 1; g.has_pos = 1; g.kv = kv; g.cnt = m.cnt; g.start = m.start; g.src = m.src; g.rhat = fbuf(e, "map_rhat");
        RET(launch_attn(e, rows, g, e->mp[i], &nb));
        RET(launch_node(e, rows, &e->mp[i], i < 2 ? &e->mp[i + 1] : nullptr, true, kv, false, nullptr, &nb));
    }
    e->map_P = P;
    if (x_pt_out) CK(cudaMemcpyAsync(x_pt_out, x, (size_t)P * 128 * sizeof(float), out_kind(loc), st));
    if (logits_out) {                                  // token_predict_head of every token (the caller selects pt_pred_mask)
        float *lg;
        RET(ensure_t(e, "map_logits", (size_t)P * MAP_TOKEN_SIZE, &lg));
        MlpLayerArgs la;
        memset(&la, 0, sizeof(la));
        la.n = P; la.x = x; la.w = e->h_map_tok; la.out = lg;
        k_mlp_layer<<<dim3((P + HM - 1) / HM, la.w.n_pad / 128, 1), NT_S, MLP_LAYER_SMEM, st>>>(la);
        CKL(); count_launch(e);
        CK(cudaMemcpyAsync(logits_out, lg, (size_t)P * MAP_TOKEN_SIZE * sizeof(float), out_kind(loc), st));
    }
    if (loc == INFGEN_HOST) {
        CK(cudaStreamSynchronize(st));
        RET(check_ftc_watchdog());
    }
    return 0;
}

static void prof_clear(infgen_engine *e) {
    for (auto &r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    e->prof.clear();
}
int32_t infgen_set_profile(infgen_engine *e, int32_t on) {
    if (!e) return fail(INFGEN_ERR_INVALID_ARG, "null engine");
    CK(cudaStreamSynchronize(e->stream));
    prof_clear(e);
    e->profile = on != 0;
    return 0;
}
int32_t infgen_profile_class_count(void) { return KC_COUNT; }
const char *infgen_profile_class_name(int32_t cls) { return (cls >= 0 && cls < KC_COUNT) ? KCLASS_NAME[cls] : nullptr; }
int32_t infgen_profile_read(infgen_engine *e, int32_t cls, double *total_ms, int64_t *count) {
    if (!e || !total_ms || !count) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    CK(cudaStreamSynchronize(e->stream));
    double tot = 0; int64_t n = 0;
    for (auto &r : e->prof) {
        if (r.cls != cls) continue;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, r.a, r.b));
        tot += ms; n++;
    }
    *total_ms = tot; *count = n;
    return 0;
}

int64_t infgen_debug_read(infgen_engine *e, const char *name, void *dst, int64_t max_bytes) {
    if (!e || !name || !dst) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    auto it = e->bufs.find(name);
    if (it == e->bufs.end() || !it->second.p) return fail(INFGEN_ERR_INVALID_ARG, "no buffer named '%s'", name);
    const int64_t n = std::min<int64_t>(max_bytes, (int64_t)it->second.bytes);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(dst, it->second.p, (size_t)n, cudaMemcpyDeviceToHost));
    return n;
}

#ifdef INFGEN_NODE_TRACE
int32_t infgen_debug_node_trace(long long *dst /* [32] */) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(dst, g_node_trace, sizeof(long long) * 32);
    return 0;
}
#endif
#ifdef INFGEN_NTC_TRACE
int32_t infgen_debug_ntc_trace(long long *dst /* [2][64] */) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(dst, g_ntc_trace, sizeof(long long) * 2 * 64);
    return 0;
}
#endif
#ifdef INFGEN_FTC_TRACE
int32_t infgen_debug_ftc_trace(long long *dst /* [2][3][64] */) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(dst, g_ftc_trace, sizeof(long long) * 2 * 3 * 64);
    return 0;
}
#endif
#ifdef INFGEN_WS_TRACE
int32_t infgen_debug_ws_trace(long long *dst /* [2][256] */, int32_t *n /* [2] */) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(dst, g_ws_trace, sizeof(long long) * 512);
    cudaMemcpyFromSymbol(n, g_ws_trace_n, sizeof(int) * 2);
    return 0;
}
#endif
// ---------------------------------------------------------------------------------------------------------------
// operator level
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"
struct TmpDev {
    std::vector<void *> ptrs;
    ~TmpDev() { for (void *p : ptrs) cudaFree(p); }
    template <typename Tp> Tp *alloc(size_t n) {
        void *p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(Tp)) != cudaSuccess) return nullptr;
        // the memset runs on the legacy default stream, asynchronously to the host, and the engine stream is non-blocking:
        // without the synchronisation it can land AFTER the kernels of the operator call have written the buffer (found as an
        // order-dependent failure of test_attention_layer: zeroed K/V rows)
        cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(Tp));
        cudaStreamSynchronize(cudaStreamLegacy);
        ptrs.push_back(p);
        return (Tp *)p;
    }
    template <typename Tp> Tp *upload(const Tp *h, size_t n) {
        Tp *d = alloc<Tp>(n);
        if (d && n) cudaMemcpy(d, h, n * sizeof(Tp), cudaMemcpyHostToDevice);
        return d;
    }
};
extern "C" {

// ---------------------------------------------------------------------------------------------------------------
// row f2: TokenProcessor._tokenize_agent + InfGen._fetch_enterings of one scene (prep.cuh)
// ---------------------------------------------------------------------------------------------------------------
int32_t infgen_prepare_scene(infgen_engine *e, const infgen_prep_in *in, const infgen_prep_out *out) {
    if (!e || !in || !out) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    const int A = in->n_agents, N = in->n_steps, P = in->n_pt, T = N / e->cfg.shift;
    if (A <= 0 || N < 2 * e->cfg.shift || N > PREP_MAX_STEPS || T > 128 || P < 0 || in->av_index < 0 || in->av_index >= A)
        return fail(INFGEN_ERR_INVALID_ARG, "prepare_scene: A=%d N=%d P=%d av=%d out of range", A, N, P, in->av_index);
    if (!in->valid_mask || !in->heading || !in->position || !in->velocity || !in->type || (P > 0 && !in->pt_position) ||
        !out->token_idx || !out->state_idx || !out->token_pos || !out->token_heading)
        return fail(INFGEN_ERR_INVALID_ARG, "prepare_scene: missing input / output array");
    // one arena for inputs, results and scratch, kept by the engine and grown on demand (no allocation per scene)
    const size_t AT = (size_t)A * T;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_valid = take((size_t)A * N), o_head = take((size_t)A * N * 4), o_pos = take((size_t)A * N * 12),
                 o_vel = take((size_t)A * N * 8), o_type = take(A), o_pt = take((size_t)std::max(P, 1) * 12);
    const size_t o_tok = take(AT * 8), o_state = take(AT * 8), o_con = take(AT * 32), o_tpos = take(AT * 8), o_thead = take(AT * 4),
                 o_rv = take(AT), o_av = take(AT), o_grid = take(AT * 8), o_goff = take(AT * 8), o_pxy = take(AT * 8),
                 o_htok = take(AT * 8), o_hth = take(AT * 4), o_sort = take(AT * 8), o_inr = take(AT), o_bos = take(AT),
                 o_ptg = take((size_t)T * std::max(P, 1) * 8), o_bear = take(AT * 4);
    if (off > e->prep_bytes) {
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(e->prep_buf);
        e->prep_buf = nullptr; e->prep_bytes = 0;
        CK(cudaMalloc(&e->prep_buf, off));
        e->prep_bytes = off;
    }
    char *base = (char *)e->prep_buf;
    auto up = [&](size_t o, const void *src, size_t bytes) {
        return cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, e->stream);
    };
    CK(up(o_valid, in->valid_mask, (size_t)A * N)); CK(up(o_head, in->heading, (size_t)A * N * 4));
    CK(up(o_pos, in->position, (size_t)A * N * 12)); CK(up(o_vel, in->velocity, (size_t)A * N * 8));
    CK(up(o_type, in->type, A));
    if (P > 0) CK(up(o_pt, in->pt_position, (size_t)P * 12));
    TokenizeArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.A = A; ta.N = N; ta.T = T; ta.V = e->cfg.token_size; ta.predict_state = 1;
    ta.valid = (const unsigned char *)(base + o_valid); ta.heading = (const float *)(base + o_head);
    ta.pos = (const float *)(base + o_pos); ta.vel = (const float *)(base + o_vel);
    ta.type = (const unsigned char *)(base + o_type); ta.vocab = e->vocab;
    ta.token_idx = (long long *)(base + o_tok); ta.state_idx = (long long *)(base + o_state);
    ta.contour = (float *)(base + o_con); ta.token_pos = (float *)(base + o_tpos); ta.token_heading = (float *)(base + o_thead);
    ta.raw_valid = (unsigned char *)(base + o_rv); ta.agent_valid = (unsigned char *)(base + o_av);
    EnterArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.A = A; ea.T = T; ea.P = P; ea.G = e->cfg.grid_size; ea.av = in->av_index;
    ea.radius = e->cfg.pl2seed_radius; ea.angle_interval = e->cfg.angle_interval;
    ea.token_pos = ta.token_pos; ea.token_heading = ta.token_heading; ea.state_idx = ta.state_idx;
    ea.pt_pos = (const float *)(base + o_pt); ea.cells = e->grid_cells;
    ea.grid_idx = (long long *)(base + o_grid); ea.grid_off = (float *)(base + o_goff); ea.pos_xy = (float *)(base + o_pxy);
    ea.head_tok = (long long *)(base + o_htok); ea.head_theta = (float *)(base + o_hth); ea.sort_idx = (long long *)(base + o_sort);
    ea.inrange = (unsigned char *)(base + o_inr); ea.bos = (unsigned char *)(base + o_bos);
    ea.pt_grid = (long long *)(base + o_ptg); ea.bearing = (float *)(base + o_bear);
    if (ea.angle_interval <= 0.f) ea.angle_interval = 3.0f;
    if (ea.radius <= 0.f) ea.radius = 75.0f;
    k_tokenize_agents<<<A, PREP_NT, 0, e->stream>>>(ta);
    CKL(); count_launch(e);
    k_fetch_enterings<<<dim3((A + P + 7) / 8, T), 256, 0, e->stream>>>(ea);
    CKL(); count_launch(e);
    k_sort_enterings<<<T, 128, 0, e->stream>>>(ea);
    CKL(); count_launch(e);
    auto down = [&](void *dst, const void *src, size_t bytes) {
        if (dst) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream);
    };
    down(out->token_idx, ta.token_idx, AT * 8); down(out->state_idx, ta.state_idx, AT * 8);
    down(out->token_contour, ta.contour, AT * 32); down(out->token_pos, ta.token_pos, AT * 8);
    down(out->token_heading, ta.token_heading, AT * 4); down(out->raw_agent_valid_mask, ta.raw_valid, AT);
    down(out->agent_valid_mask, ta.agent_valid, AT);
    down(out->grid_token_idx, ea.grid_idx, AT * 8); down(out->grid_offset_xy, ea.grid_off, AT * 8);
    down(out->heading_token_idx, ea.head_tok, AT * 8); down(out->pos_xy, ea.pos_xy, AT * 8);
    down(out->heading_theta, ea.head_theta, AT * 4); down(out->sort_indices, ea.sort_idx, AT * 8);
    down(out->inrange_mask, ea.inrange, AT); down(out->bos_mask, ea.bos, AT);
    if (P > 0) down(out->pt_grid_token_idx, ea.pt_grid, (size_t)T * P * 8);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return 0;
}

int32_t infgen_match_map_tokens(infgen_engine *e, const infgen_map_match_in *in, const infgen_map_match_out *out) {
    if (!e || !in || !out) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    const int P = in->n_tokens, V = in->n_vocab, NP = in->n_polygons;
    if (P <= 0 || V <= 0 || V > 8192 || NP <= 0)
        return fail(INFGEN_ERR_INVALID_ARG, "match_map_tokens: P=%d V=%d polygons=%d out of range", P, V, NP);
    if (!in->traj_pos || !in->traj_theta || !in->pl_rank || !in->side || !in->sample_pt || !out->token_idx ||
        !out->position || !out->orientation || !out->side_counts)
        return fail(INFGEN_ERR_INVALID_ARG, "match_map_tokens: missing input / output array");
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_pos = take((size_t)P * 24), o_th = take((size_t)P * 4), o_rank = take((size_t)P * 4), o_side = take(P),
                 o_voc = take((size_t)V * 24), o_tok = take((size_t)P * 8), o_out = take((size_t)P * 12), o_ori = take((size_t)P * 4),
                 o_best = take((size_t)P * 4), o_cnt = take((size_t)NP * 12);
    if (off > e->prep_bytes) {
        CK(cudaStreamSynchronize(e->stream));
        cudaFree(e->prep_buf);
        e->prep_buf = nullptr; e->prep_bytes = 0;
        CK(cudaMalloc(&e->prep_buf, off));
        e->prep_bytes = off;
    }
    char *base = (char *)e->prep_buf;
    auto up = [&](size_t o, const void *src, size_t bytes) {
        return cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, e->stream);
    };
    CK(up(o_pos, in->traj_pos, (size_t)P * 24)); CK(up(o_th, in->traj_theta, (size_t)P * 4));
    CK(up(o_rank, in->pl_rank, (size_t)P * 4)); CK(up(o_side, in->side, P)); CK(up(o_voc, in->sample_pt, (size_t)V * 24));
    CK(cudaMemsetAsync(base + o_cnt, 0, (size_t)NP * 12, e->stream));
    MapMatchArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.P = P; ma.V = V;
    ma.traj_pos = (const float *)(base + o_pos); ma.traj_theta = (const float *)(base + o_th);
    ma.pl_rank = (const int *)(base + o_rank); ma.side = (const unsigned char *)(base + o_side);
    ma.sample_pt = (const float *)(base + o_voc);
    ma.token_idx = (long long *)(base + o_tok); ma.position = (float *)(base + o_out); ma.orientation = (float *)(base + o_ori);
    ma.best = (float *)(base + o_best); ma.counts = (int *)(base + o_cnt);
    const size_t smem = (size_t)V * 24;
    if (smem > 48 * 1024)
        CK(cudaFuncSetAttribute(k_match_map_tokens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_match_map_tokens<<<(P + MATCH_NT / 32 - 1) / (MATCH_NT / 32), MATCH_NT, smem, e->stream>>>(ma);
    CKL(); count_launch(e);
    CK(cudaMemcpyAsync(out->token_idx, ma.token_idx, (size_t)P * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(out->position, ma.position, (size_t)P * 12, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(out->orientation, ma.orientation, (size_t)P * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(out->side_counts, ma.counts, (size_t)NP * 12, cudaMemcpyDeviceToHost, e->stream));
    if (out->best_distance) CK(cudaMemcpyAsync(out->best_distance, ma.best, (size_t)P * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaGetLastError());
    return 0;
}

int32_t infgen_op_attention_layer(infgen_engine *e, const char *layer, const float *x_src, int32_t n_src,
                                  const float *x_dst, int32_t n_dst, const float *r, const int32_t *edge_ptr,
                                  const int32_t *edge_src, float *out) {
    if (!e || !layer || !x_dst || !edge_ptr || !out) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    std::string p(layer);
    if (g_index.find(p + ".w_qs") == g_index.end()) return fail(INFGEN_ERR_INVALID_ARG, "unknown layer '%s'", layer);
    const bool has_pos = g_index.find(p + ".w_kr") != g_index.end() && r != nullptr;
    AttnW w = make_attn(e, p, g_index.find(p + ".w_kr") != g_index.end());
    w.has_pos = has_pos ? 1 : 0;
    const bool bip = x_src != nullptr;
    const int E = edge_ptr[n_dst];
    std::vector<int> cnt(n_dst), start(n_dst);
    for (int i = 0; i < n_dst; ++i) {
        cnt[i] = edge_ptr[i + 1] - edge_ptr[i]; start[i] = edge_ptr[i];
        if (cnt[i] < 0) return fail(INFGEN_ERR_INVALID_ARG, "edge_ptr must be non-decreasing");
    }
    CK(cudaStreamSynchronize(e->stream));
    TmpDev tmp;
    float *d_x = tmp.upload(x_dst, (size_t)n_dst * 128);
    float *d_q = tmp.alloc<float>((size_t)n_dst * 128), *d_s = tmp.alloc<float>((size_t)n_dst * 128);
    float *d_qr = tmp.alloc<float>((size_t)n_dst * 1024);
    float *d_out = tmp.alloc<float>((size_t)n_dst * 128);
    const int n_kv = bip ? n_src : n_dst;
    float *d_kv = tmp.alloc<float>((size_t)n_kv * 256);
    int *d_cnt = tmp.upload(cnt.data(), n_dst), *d_start = tmp.upload(start.data(), n_dst);
    int *d_src = tmp.upload(edge_src, (size_t)E);
    float *d_rhat = nullptr;
    if (has_pos) {
        float *d_r = tmp.upload(r, (size_t)E * 128);
        d_rhat = tmp.alloc<float>((size_t)E * 128);
        if (E > 0) {
            k_standardize<<<(E + NWARP - 1) / NWARP, NT, 0, e->stream>>>(d_r, d_rhat, E);
            count_launch(e);
        }
    }
    if (!d_x || !d_out || !d_kv || !d_src) return fail(INFGEN_ERR_CUDA, "temporary allocation failed");
    if (e->layer_path == 2 && w.npk) {
        // row-tile path (node.cuh): projections -> k_attn -> k_node, on a copy of x
        NodeBufs nb{d_out, d_q, d_s, d_qr, tmp.alloc<float>((size_t)n_dst * 128)};
        if (!nb.agg) return fail(INFGEN_ERR_CUDA, "temporary allocation failed");
        CK(cudaMemcpyAsync(d_out, d_x, (size_t)n_dst * 128 * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
        const RowSpace rows = flat_rows(n_dst);
        RET(launch_node(e, rows, nullptr, &w, !bip, d_kv, false, nullptr, &nb));
        if (bip) {
            float *d_xs = tmp.upload(x_src, (size_t)n_src * 128);
            KvArgs ka;
            memset(&ka, 0, sizeof(ka));
            ka.n = n_src; ka.x = d_xs; ka.w[0] = w; ka.out[0] = d_kv;
            k_kv_project<16><<<dim3((n_src + 15) / 16, 1), NT, 0, e->stream>>>(ka);
            CKL(); count_launch(e);
        }
        SubArgs g;
        memset(&g, 0, sizeof(g));
        g.has_attn = 1; g.has_pos = has_pos ? 1 : 0; g.kv = d_kv; g.cnt = d_cnt; g.start = d_start; g.src = d_src; g.rhat = d_rhat;
        RET(launch_attn(e, rows, g, w, &nb));
        RET(launch_node(e, rows, &w, nullptr, false, nullptr, false, nullptr, &nb));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaMemcpy(out, d_out, (size_t)n_dst * 128 * sizeof(float), cudaMemcpyDeviceToHost));
        return 0;
    }
    const int saved_tile = e->row_tile;
    e->row_tile = (n_dst + 3) / 4 <= MAX_CLUSTERS ? 4 : 8;
    // launch 1: projections of every row (K/V of all rows must exist before any row attends)
    LayerArgs la;
    memset(&la, 0, sizeof(la));
    la.rows = flat_rows(n_dst); la.x = d_x; la.q = d_q; la.s = d_s; la.qr = d_qr; la.ring = RING;
    la.pre0 = make_pre(w, !bip, d_kv, false, 0, true);
    la.n_sub = 0;
    int rc = launch_layer(e, la, KC_MISC);
    if (rc == 0 && bip) {
        float *d_xs = tmp.upload(x_src, (size_t)n_src * 128);
        KvArgs ka;
        memset(&ka, 0, sizeof(ka));
        ka.n = n_src; ka.x = d_xs; ka.w[0] = w; ka.out[0] = d_kv;
        k_kv_project<16><<<dim3((n_src + 15) / 16, 1), NT, 0, e->stream>>>(ka);
        count_launch(e);
    }
    if (rc == 0) {
        // launch 2: attention + update + FFN, in place on a copy of x
        if (cudaMemcpyAsync(d_out, d_x, (size_t)n_dst * 128 * sizeof(float), cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess)
            rc = fail(INFGEN_ERR_CUDA, "copy failed");
        memset(&la, 0, sizeof(la));
        la.rows = flat_rows(n_dst); la.x = d_out; la.q = d_q; la.s = d_s; la.qr = d_qr; la.ring = RING;
        la.n_sub = 1;
        SubArgs &g = la.sub[0];
        g.w = w.cs_post; g.has_attn = 1; g.has_pos = has_pos ? 1 : 0; g.kv = d_kv; g.cnt = d_cnt; g.start = d_start;
        g.src = d_src; g.rhat = d_rhat;
        if (rc == 0) rc = launch_layer(e, la, KC_MISC);
    }
    e->row_tile = saved_tile;
    RET(rc);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out, d_out, (size_t)n_dst * 128 * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int32_t infgen_op_fourier_embedding(infgen_engine *e, const char *name, const float *x, int32_t n, int32_t dim,
                                    const float *cat, float *out) {
    if (!e || !name || !x || !out) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    std::string p(name);
    if (g_index.find(p + ".freqs") == g_index.end()) return fail(INFGEN_ERR_INVALID_ARG, "unknown embedding '%s'", name);
    if (g_layout[g_index[p + ".freqs"]].numel != dim * 64) return fail(INFGEN_ERR_INVALID_ARG, "'%s' has a different input_dim", name);
    CK(cudaStreamSynchronize(e->stream));
    TmpDev tmp;
    FourierArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.n_slots = n; fa.dim = dim; fa.raw = tmp.upload(x, (size_t)n * dim); fa.w = make_fourier(e, p, dim);
    fa.cat_tab = cat ? tmp.upload(cat, (size_t)n * 128) : nullptr;
    float *d_out = tmp.alloc<float>((size_t)n * 128);
    fa.out = d_out; fa.normalize = 0;
    RET(launch_fourier(e, &fa, 1));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out, d_out, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost));
    RET(check_ftc_watchdog());
    return 0;
}

int32_t infgen_op_mlp_embedding(infgen_engine *e, const char *name, const float *x, int32_t n, int32_t dim,
                                float *out) {
    if (!e || !name || !x || !out) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    std::string p(name);
    if (g_index.find(p + ".w6") == g_index.end()) return fail(INFGEN_ERR_INVALID_ARG, "unknown embedding '%s'", name);
    if (g_layout[g_index[p + ".w0"]].numel != gemm_numel(dim, 128)) return fail(INFGEN_ERR_INVALID_ARG, "'%s' has a different input_dim", name);
    CK(cudaStreamSynchronize(e->stream));
    TmpDev tmp;
    MlpEmbArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.rows = flat_rows(n); ma.w = make_mlp_emb(e, p); ma.kin = dim; ma.k4 = (dim + 3) / 4;
    ma.x = tmp.upload(x, (size_t)n * dim); ma.x_ld = dim;
    float *d_out = tmp.alloc<float>((size_t)n * 128);
    ma.out = d_out; ma.out_ld = 128;
    if (ma.k4 > 128) return fail(INFGEN_ERR_INVALID_ARG, "input_dim %d too large", dim);
    RET(launch_mlp_embed(e, ma));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out, d_out, (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int32_t infgen_op_mlp_layer(infgen_engine *e, const char *name, const float *x, int32_t n, float *out) {
    if (!e || !name || !x || !out) return fail(INFGEN_ERR_INVALID_ARG, "null argument");
    std::string p(name);
    if (g_index.find(p + ".w3") == g_index.end() || g_index.find(p + ".ln.g") == g_index.end())
        return fail(INFGEN_ERR_INVALID_ARG, "unknown head '%s'", name);
    if (g_layout[g_index[p + ".w0"]].numel != gemm_numel(128, 128)) return fail(INFGEN_ERR_INVALID_ARG, "'%s' does not take 128 inputs", name);
    const int n_pad = (int)g_layout[g_index[p + ".b3"]].numel;
    int n_out = n_pad;
    if (p == "state_predict_head" || p == "seed_type_predict_head" || p == "seed_shape_predict_head") n_out = 3;
    else if (p == "seed_state_predict_head" || p == "seed_offset_xy_predict_head") n_out = 2;
    else if (p == "seed_heading_rel_token_predict_head") n_out = ANGLE_SIZE;
    else if (p == "token_predict_head") n_out = TOKEN_SIZE;
    else n_out = GRID_SIZE;
    CK(cudaStreamSynchronize(e->stream));
    TmpDev tmp;
    MlpLayerArgs la;
    memset(&la, 0, sizeof(la));
    la.n = n; la.x = tmp.upload(x, (size_t)n * 128); la.w = make_head(e, p, 128, n_out);
    float *d_out = tmp.alloc<float>((size_t)n * n_out);
    la.out = d_out;
    k_mlp_layer<<<dim3((n + HM - 1) / HM, la.w.n_pad / 128), NT_S, MLP_LAYER_SMEM, e->stream>>>(la);
    CKL(); count_launch(e);
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpy(out, d_out, (size_t)n * n_out * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
