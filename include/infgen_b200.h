/*
 * infgen_b200.h - C ABI of the B200-native (sm_100a) closed-loop decode engine.
 *
 * This library replaces ONE path of OrangeSodahub/InfGen: `InfGenAgentDecoder.inference`
 * (reference infgen/modules/agent_decoder.py:1605-2389) and the operators under it
 * (reference infgen/modules/layers.py:16-215).  The reference is pure Python/PyTorch and has no FFI of its
 * own; the binding a maintainer adds is the ctypes stub shown in INTEGRATION.md (the shipped one is
 * infgen_b200/_capi.py).  Every entry point below cites the reference interface it stands in for.
 *
 * Conventions
 *   - plain C, no torch types.  All pointers are caller-owned; `loc` says where they live
 *     (INFGEN_HOST: pageable or pinned host memory, INFGEN_DEVICE: device memory of the engine's GPU).
 *   - every function returns 0 on success and a negative `infgen_status` otherwise; the text of the last
 *     failure (per thread) is available from infgen_last_error().
 *   - an engine is not thread-safe; all device work is enqueued on the engine's stream
 *     (infgen_set_stream to share the caller's stream, e.g. torch.cuda.current_stream().cuda_stream).
 *   - rows: agents of all scenes of a batch live in one row space, scene b owning rows
 *     [b*row_capacity, b*row_capacity + n_rows[b]).  Columns are the reference's 2 Hz token columns.
 */
#ifndef INFGEN_B200_H
#define INFGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INFGEN_ABI_VERSION 8

typedef enum {
    INFGEN_OK = 0,
    INFGEN_ERR_INVALID_ARG = -1,
    INFGEN_ERR_CAPACITY = -2,
    INFGEN_ERR_CUDA = -3,
    INFGEN_ERR_STATE = -4,
    INFGEN_ERR_NO_DEVICE = -5
} infgen_status;

typedef enum { INFGEN_HOST = 0, INFGEN_DEVICE = 1 } infgen_loc;

typedef struct infgen_engine infgen_engine;

/* Run-time knobs of `InfGenAgentDecoder.__init__` (agent_decoder.py:100-314) that reach the decode loop. */
typedef struct {
    int32_t abi_version;          /* must be INFGEN_ABI_VERSION */
    int32_t device;               /* CUDA device ordinal */
    int32_t num_layers;           /* 6   num_agent_layers */
    int32_t hist_cols;            /* 2   (num_historical_steps-1)//shift, agent_decoder.py:1636 */
    int32_t window;               /* 12  time_span/shift, agent_decoder.py:586-587 */
    int32_t shift;                /* 5 */
    int32_t num_historical_steps; /* 11 */
    int32_t token_size;           /* 2048 */
    int32_t grid_size;            /* 1961 cells of Attr_Tokenizer, attr_tokenizer.py:24-43 */
    int32_t num_seed_feature;     /* 10: rows at the tail of a scene that get no temporal edges, :553-556 */
    int32_t max_pl2a_neighbors;   /* 5   agent_decoder.py:711 */
    int32_t max_a2a_neighbors;    /* 300 agent_decoder.py:633 */
    float pl2a_radius;            /* 30 */
    float a2a_radius;             /* 60 */
    int32_t use_state_token;      /* agent_decoder.py:2170-2171 */
    int32_t disable_insertion;    /* agent_decoder.py:2172-2173: every predicted state forced to 'valid' */
    int32_t motion_beam_size;     /* top-k of the motion-token sampler, 1 = greedy (agent_decoder.py:300, 2163) */
    uint32_t seed;                /* sampler seed (counter-based; see DESIGN.md "sampler") */
    int32_t use_cuda_graph;       /* replay one captured decode iteration per step */
    int32_t trace;                /* keep per-iteration head inputs / logits / layer outputs for parity tests */
    /* insertion stage (agent_decoder.py:1744-2114); only read when disable_insertion == 0 */
    int32_t insert_beam_size;     /* 10  top-k of the position sampler, 1 = greedy (agent_decoder.py:301, 1899) */
    int32_t debug_force_enter;    /* the reference's DEBUG=1 switch: the seed head always answers 'enter' (:1888-1889) */
    float pl2seed_radius;         /* 75  radius of the seed query (agent_decoder.py:1837, 1846) */
    float a2sa_radius;            /* 10  heading stage, agents (agent_decoder.py:2028) */
    float pl2sa_radius;           /* 10  heading stage, map tokens (agent_decoder.py:2034) */
    float angle_interval;         /* 3   degrees per heading token (attr_tokenizer.py:20) */
    /* teacher-forced pass (InfGenAgentDecoder.forward, agent_decoder.py:1104-1240): every column of the batch is given
     * (hist_cols == n_cols, n_iters == 0) and every column is a destination; run with infgen_forward instead of
     * prefill / step.  num_seed_feature must be 0 (the seed rows of _pad_feat are not part of the row space). */
    int32_t teacher_forced;
} infgen_config;

/* One batch of scenes, already filtered/padded as agent_decoder.py:1609-1657 does (host side: infgen_b200/host.py). */
typedef struct {
    int32_t n_scenes;
    int32_t row_capacity;         /* rows reserved per scene (>= max n_rows; multiple of 4) */
    int32_t n_cols;               /* T = num_infer_step, agent_decoder.py:1637 */
    int32_t n_iters;              /* S = num_recurrent_steps_val // shift, agent_decoder.py:1741 */
    const int32_t *n_rows;        /* [n_scenes] agents valid at the current column */
    const int32_t *ego_row;       /* [n_scenes] row (within the scene) of the ego agent, agent_decoder.py:1648-1650 */
    const int32_t *scene_id;      /* [n_scenes] sampler stream id of each scene */
    /* per row, [n_scenes*row_capacity, ...]; only rows < n_rows[scene] are read */
    const float *pos_hist;        /* [R][hist_cols][2]  token_pos   */
    const float *head_hist;       /* [R][hist_cols]     token_heading */
    const int32_t *state_hist;    /* [R][hist_cols]     state_idx   */
    const int32_t *token_hist;    /* [R][hist_cols]     token_idx (-1 none, -2 BOS) */
    const int32_t *grid_hist;     /* [R][hist_cols]     grid_token_idx (-1 invalid) */
    const uint8_t *tsrc_hist;     /* [R][hist_cols]     column usable as temporal source (temporal_mask & col>=bos) */
    const uint8_t *interact_hist; /* [R][hist_cols]     interact_mask, agent_decoder.py:1707-1719 */
    const int32_t *type;          /* [R] 0 veh 1 ped 2 cyc */
    const float *shape;           /* [R][3] shape at the current step */
    /* map tokens of all scenes, concatenated; scene b owns [pt_ptr[b], pt_ptr[b+1]) */
    const int32_t *pt_ptr;        /* [n_scenes+1] */
    const float *pt_pos;          /* [P][2] */
    const float *pt_ori;          /* [P]    */
    const float *x_pt;            /* [P][128] map encoder output, agent_decoder.py:2143; NULL = the output of this engine's
                                     infgen_map_encode on the same tokens, which then never leaves HBM */
} infgen_scene_batch;

/* Result buffers of one batch (any pointer may be NULL = not wanted). Layout [R = n_scenes*row_capacity][...]. */
typedef struct {
    float *pos;                   /* [R][T][2]   pos_a   */
    float *head;                  /* [R][T]      head_a  */
    float *pred_traj;             /* [R][5*S][2] generated part of pred_traj (agent_decoder.py:2201-2204) */
    float *pred_head;             /* [R][5*S]    */
    float *pred_state;            /* [R][5*S]    */
    int32_t *next_token;          /* [R][T]      history columns then one sampled token per iteration */
    int32_t *next_state;          /* [R][T]      */
    float *hist_traj;             /* [R][hist_cols*5][2] raw steps 1..10 rebuilt from history tokens (:2311-2335) */
    float *hist_head;             /* [R][hist_cols*5] */
    /* insertion stage (NULL when not wanted; zero-filled when the stage is disabled) */
    int32_t *n_rows_final;        /* [n_scenes] rows of each scene after the rollout (appended agents included) */
    int32_t *pred_type;           /* [R] predicted type of appended rows (:1955) */
    float *pred_shape;            /* [R][3] predicted shape of appended rows (:1956) */
    /* Wire format of the insertion records.  The reference returns five dense tensors
     * next_state_prob_seed [11][S], next_pos_rel_prob_seed / grid_agent_occ_seed / grid_pt_occ_seed /
     * grid_agent_occ_gt_seed [11][S][grid] (agent_decoder.py:2099-2113, 2376-2386) that are zero except at [slot][t] of
     * every insertion (slot = 1..10 within decode iteration t): 5.5 MB per 16-iteration scene, 104 MB per 150 s scene.
     * Here one record per APPENDED ROW travels instead, indexed by the row (rows [n_rows[b], n_rows_final[b]) of scene b;
     * nothing else is written or copied), and the host mirror scatters them into the dense tensors. */
    int32_t *rec_meta;            /* [R][2]    (decode iteration t, slot) of the insertion that appended the row */
    float *rec_state_prob;        /* [R]       next_state_prob_seed[slot][t] (:2105) */
    float *rec_pos_prob;          /* [R][grid] next_pos_rel_prob_seed[slot][t] (:2104) */
    float *rec_agent_occ;         /* [R][grid] grid_agent_occ_seed[slot][t] (:2102) */
    float *rec_pt_occ;            /* [R][grid] grid_pt_occ_seed[slot][t] (:2103) */
    float *rec_occ_gt;            /* [R][grid] grid_agent_occ_gt_seed[slot][t] (:2101) */
} infgen_outputs;

/* ---- library ------------------------------------------------------------------------------------------- */
int32_t infgen_abi_version(void);
const char *infgen_last_error(void);

/* ---- weights: replaces nn.Module.load_state_dict for the decoder (names: infgen_b200/weights.py) ---------- */
/* The library owns the packed layout; the host packs `state_dict()` tensors into one float blob by asking for
 * the offset/shape of each packed tensor by name. */
int32_t infgen_weight_count(void);
const char *infgen_weight_name(int32_t i);
int64_t infgen_weight_offset(const char *name);   /* in floats, -1 if unknown */
int64_t infgen_weight_numel(const char *name);
int64_t infgen_weight_blob_floats(void);

/* ---- engine life cycle ---------------------------------------------------------------------------------- */
/* InfGenAgentDecoder.__init__ + load_state_dict (agent_decoder.py:100-314).
 * weights: packed blob (host).  grid_cells: [grid_size][2] Attr_Tokenizer.grid (attr_tokenizer.py:24-43, host).
 * vocab: [3][token_size][6][4][2] motion-token box tracks veh/ped/cyc (preprocess.py:302-311, host). */
int32_t infgen_create(const infgen_config *cfg, const float *weights, int64_t n_floats, const float *grid_cells,
                      const float *vocab, infgen_engine **out);
int32_t infgen_destroy(infgen_engine *e);
int32_t infgen_set_stream(infgen_engine *e, void *cuda_stream);   /* NULL = engine-owned stream */
int32_t infgen_set_sampler(infgen_engine *e, int32_t motion_beam_size, uint32_t seed);
int32_t infgen_synchronize(infgen_engine *e);

/* ---- the decode path: InfGenAgentDecoder.inference (agent_decoder.py:1605-2389) ------------------------------ */
/* setup (:1609-1719): upload one batch, build per-scene caches (map K/V of the six pt2a layers, embeddings).
 * `loc` applies to the per-row / per-map-token arrays; n_rows, ego_row, scene_id and pt_ptr are always host. */
int32_t infgen_load_scenes(infgen_engine *e, const infgen_scene_batch *batch, int32_t loc);
/* teacher forcing for parity tests: [R][S] tokens / states applied instead of the sampled ones (NULL = off). */
int32_t infgen_set_forcing(infgen_engine *e, const int32_t *tokens, const int32_t *states, int32_t loc);
/* history columns through the layer stack (what iteration 0 of the reference computes for column 0, :2149-2150) */
int32_t infgen_prefill(infgen_engine *e);
/* n decode iterations of the loop at :1740-2301 (edges, 6x{temporal, map, agent} attention, heads, sampling,
 * token->pose advance, next-column embedding).  Asynchronous on the engine stream. */
int32_t infgen_step(infgen_engine *e, int32_t n_iters);
/* prefill + all remaining iterations */
int32_t infgen_rollout(infgen_engine *e);
/* outputs (:2303-2389). Synchronises the stream when loc == INFGEN_HOST, and whenever insertion records are requested
 * (their extent is read back first). */
int32_t infgen_read(infgen_engine *e, const infgen_outputs *out, int32_t loc);
int32_t infgen_iterations_done(infgen_engine *e);
/* Motion branch of the teacher-forced pass (agent_decoder.py:1104-1240) over the loaded batch of
 * an engine created with teacher_forced = 1: column by column, embedding -> edges whose destination is that column ->
 * 6 x {temporal, map, agent} attention (K/V of the earlier columns from the cache) -> heads.  Outputs are COLUMN-major,
 * [n_cols][R][...] with R = n_scenes * row_capacity (any may be NULL): x_a [T][R][128] last-layer features,
 * token_logits [T][R][token_size] (next_token_prob), state_logits [T][R][3] (next_state_prob).  Synchronises. */
int32_t infgen_forward(infgen_engine *e, float *x_a, float *token_logits, float *state_logits, int32_t loc);
/* number of kernels this library launched (or replayed through graphs) since the engine was created */
int64_t infgen_kernel_launches(infgen_engine *e);

/* ---- the map encoder: InfGenMapDecoder.forward (map_decoder.py:70-130), SURVEY.md 8f row f1 ---------------------- */
/* Map tokens of a batch of scenes, concatenated; scene b owns [pt_ptr[b], pt_ptr[b+1]).  `loc` applies to the per-token
 * arrays; pt_ptr is always host.  light_type is the owning polygon's light type gathered per token (map_decoder.py:85-86). */
typedef struct {
    int32_t n_scenes;
    const int32_t *pt_ptr;        /* [n_scenes+1] */
    const float *pt_pos;          /* [P][2] data['pt_token']['position'][:, :2] */
    const float *pt_ori;          /* [P]    data['pt_token']['orientation'] */
    const int32_t *type;          /* [P]    data['pt_token']['type']    (17 classes) */
    const int32_t *pl_type;       /* [P]    data['pt_token']['pl_type'] (4) */
    const int32_t *light_type;    /* [P]    data['map_polygon']['light_type'][token2pl[1]] (4) */
    const int32_t *token_idx;     /* [P]    data['pt_token']['token_idx'] (map vocabulary, 1024) */
    float pl2pl_radius;           /* 10     radius_graph radius (map_decoder.py:91), max_num_neighbors = 100 */
} infgen_map_batch;
/* token_emb over the map vocabulary (map_decoder.py:78-80): traj_src [n_tokens][22] = map_token['traj_src'].view(n, -1), host */
int32_t infgen_map_setup(infgen_engine *e, const float *traj_src, int32_t n_tokens);
/* x_pt [P][128] (and, when wanted, the token_predict_head logits [P][1024] of every token; the reference evaluates the head
 * on x_pt[pt_pred_mask], map_decoder.py:119).  Either output may be NULL; the result also stays in the engine for a
 * following infgen_load_scenes with x_pt == NULL.  Synchronises the stream when loc == INFGEN_HOST. */
int32_t infgen_map_encode(infgen_engine *e, const infgen_map_batch *batch, int32_t loc, float *x_pt_out, float *logits_out);

/* ---- per-scene preparation of the agent stream, SURVEY.md 8f row f2 ------------------------------------------------ */
/* Raw 10 Hz tracks of ONE scene (the inputs of TokenProcessor._tokenize_agent, infgen/datasets/preprocess.py:364-373) and
 * its map token positions.  All pointers host. */
typedef struct {
    int32_t n_agents, n_steps;    /* A, N raw steps (91); token steps T = n_steps / shift */
    int32_t av_index;             /* data['agent']['av_idx'] */
    int32_t n_pt;                 /* P map tokens */
    const uint8_t *valid_mask;    /* [A][N] */
    const float *heading;         /* [A][N] */
    const float *position;        /* [A][N][3] */
    const float *velocity;        /* [A][N][2] */
    const uint8_t *type;          /* [A] 0 veh 1 ped 2 cyc */
    const float *pt_position;     /* [P][3] data['pt_token']['position'] */
} infgen_prep_in;
/* Results (host, caller-allocated; any pointer of the second group may be NULL).  int64 where the reference has LongTensors. */
typedef struct {
    /* TokenProcessor._tokenize_agent (preprocess.py:533-546) */
    int64_t *token_idx;           /* [A][T] motion token, -1 invalid, -2 BOS */
    int64_t *state_idx;           /* [A][T] 0 invalid 1 valid 2 enter 3 exit */
    float *token_contour;         /* [A][T][4][2] matched box */
    float *token_pos;             /* [A][T][2] */
    float *token_heading;         /* [A][T] */
    uint8_t *raw_agent_valid_mask;/* [A][T] */
    uint8_t *agent_valid_mask;    /* [A][T] (all ones with predict_state) */
    /* InfGen._fetch_enterings (infgen/model/infgen.py:1008-1090) */
    int64_t *grid_token_idx;      /* [A][T] ego-centric cell, -1 invalid / out of range */
    float *grid_offset_xy;        /* [A][T][2] */
    int64_t *heading_token_idx;   /* [A][T] */
    float *pos_xy;                /* [A][T][2] */
    float *heading_theta;         /* [A][T] */
    int64_t *sort_indices;        /* [A][T] entering agents of a column by bearing, ego row elsewhere */
    uint8_t *inrange_mask;        /* [A][T] */
    uint8_t *bos_mask;            /* [A][T] */
    int64_t *pt_grid_token_idx;   /* [T][P] */
} infgen_prep_out;
/* tokenize + fetch_enterings of one scene on the device (vocabulary and grid cells are the engine's). */
int32_t infgen_prepare_scene(infgen_engine *e, const infgen_prep_in *in, const infgen_prep_out *out);

/* ---- row f2, map side: `InfGen.match_token_map` (infgen/model/infgen.py:918-984) on the device.  Inputs are the
 * `map_save` fields `TokenProcessor._tokenize_map` writes (preprocess.py:749-758), host arrays. --------------------- */
typedef struct {
    int32_t n_tokens;             /* P polylines of 5 m (three points each) */
    int32_t n_vocab;              /* entries of the map vocabulary (1024) */
    int32_t n_polygons;           /* distinct polygons of the scene */
    const float *traj_pos;        /* [P][3][2] data['map_save']['traj_pos'].float() */
    const float *traj_theta;      /* [P]       data['map_save']['traj_theta'].float() */
    const int32_t *pl_rank;       /* [P] row of data['map_save']['pl_idx_list'][i] among its sorted distinct values */
    const uint8_t *side;          /* [P] data['pt_token']['side'] */
    const float *sample_pt;       /* [n_vocab][3][2] map_token['sample_pt'] (infgen.py:208-209) */
} infgen_map_match_in;
typedef struct {
    int64_t *token_idx;           /* [P] data['pt_token']['token_idx'] */
    float *position;              /* [P][3] first point of the polyline, z = 0 */
    float *orientation;           /* [P] */
    int32_t *side_counts;         /* [n_polygons][3] polylines per polygon and side (rows of traj_mask, infgen.py:955-971) */
    float *best_distance;         /* [P] optional: summed squared distance of the match */
} infgen_map_match_out;
int32_t infgen_match_map_tokens(infgen_engine *e, const infgen_map_match_in *in, const infgen_map_match_out *out);

/* ---- per-kernel-class device timing for the roofline report (bench.py): CUDA events around every launch; turns
 * graph replay off while enabled ------------------------------------------------------------------------------ */
int32_t infgen_set_profile(infgen_engine *e, int32_t on);
int32_t infgen_profile_class_count(void);
const char *infgen_profile_class_name(int32_t cls);
int32_t infgen_profile_read(infgen_engine *e, int32_t cls, double *total_ms, int64_t *count);

/* ---- parity/debug taps (tests only): copy a named internal buffer to host; returns bytes written or <0 ------- */
int64_t infgen_debug_read(infgen_engine *e, const char *name, void *dst, int64_t max_bytes);

/* ---- operator level: the reference's layers.py modules, for unit parity ------------------------------------- */
/* AttentionLayer.forward((x_src, x_dst), r, edge_index) (layers.py:61-113).  `layer` is the state_dict prefix,
 * e.g. "a2a_attn_layers.3".  Edges are grouped by destination: edges of node i are
 * [edge_ptr[i], edge_ptr[i+1]) with source ids edge_src[] (PyG edge_index[0]); r is [E][128] (pre-LayerNorm).
 * x_src == NULL means the non-bipartite form (x_src = x_dst). All pointers host. */
int32_t infgen_op_attention_layer(infgen_engine *e, const char *layer, const float *x_src, int32_t n_src,
                                  const float *x_dst, int32_t n_dst, const float *r, const int32_t *edge_ptr,
                                  const int32_t *edge_src, float *out);
/* FourierEmbedding.forward(continuous_inputs, [categorical sum]) (layers.py:142-160); name e.g. "r_t_emb". */
int32_t infgen_op_fourier_embedding(infgen_engine *e, const char *name, const float *x, int32_t n, int32_t dim,
                                    const float *cat, float *out);
/* MLPEmbedding.forward (layers.py:189); name e.g. "fusion_emb". */
int32_t infgen_op_mlp_embedding(infgen_engine *e, const char *name, const float *x, int32_t n, int32_t dim,
                                float *out);
/* MLPLayer.forward (layers.py:213-215); name "token_predict_head" | "state_predict_head". */
int32_t infgen_op_mlp_layer(infgen_engine *e, const char *name, const float *x, int32_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif /* INFGEN_B200_H */
