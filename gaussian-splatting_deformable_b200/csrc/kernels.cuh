// Kernel argument blocks and host-side launchers shared between the .cu files
// and the C-ABI (api.cu).  All pointers are device pointers unless noted.
#pragma once
#include "common.cuh"

// Digit plan of a sort on bits [begin_bit, end_bit): ceil(bits/8) passes, bits split evenly.
// The temp buffer starts with the digit histograms: u32 hist[pass][256].
#define GSR_SORT_MAX_PASSES 8
#define GSR_SORT_RADIX 256
struct GsrSortPlan {
    int passes;
    int shift[GSR_SORT_MAX_PASSES];
    uint32_t mask[GSR_SORT_MAX_PASSES];
};
GsrSortPlan gsr_make_sort_plan(int begin_bit, int end_bit);

struct PreprocessArgs {
    int P;
    const float* means;          // [P,3] un-deformed
    const float* scales;         // [P,3] or null
    const float* rotations;      // [P,4] or null
    const float* opacities;      // [P]
    const float* shs;            // [P,M,3] or null
    const float* cov3D_precomp;  // [P,6] or null
    const float* colors_precomp; // [P,3] or null
    int deform_mode;
    const float* twist_S;        // [P,6] or [B,6]
    const float* twist_theta;    // [P] or [B]
    const int* body_id;          // [P] (rigid-body mode)
    float* means_out;            // [P,3] deformed means (required when deform_mode != 0)
    int* radii;                  // [P]
    float* depths;               // [P]
    uint32_t* tiles_touched;     // [P]
    float4* recs;                // [P,3] splat records
    uint8_t* clamped;            // [P] bit c set when colour channel c was clamped at 0
    float* cov3D_out;            // [P,6] or null (parity dumps)
    uint32_t* block_sums;        // [ceil(P/256)] sum of tiles_touched per 256-Gaussian block
    uint32_t* depth_keys;        // [P] float bits of the view depth (0xffffffff when the Gaussian emits nothing)
    uint32_t* depth_state;       // control words of the depth sort (GsrDepthSortCtrl, pre-zeroed): ~min / max of the emitting keys accumulated here
    uint32_t* total;             // running sum of tiles_touched = num_rendered (pre-zeroed; one atomic per block)
    int prefiltered;             // GaussianRasterizationSettings.prefiltered: a culled point is an error (auxiliary.h:156-160)
    uint32_t* flags;             // bit 0 raised when prefiltered is set and a point fails the near-plane test
    uint2* rects;                // [P] tile rect (rmin.x | rmin.y << 16, rmax.x | rmax.y << 16); (0,0) when nothing is emitted
};

int gsr_launch_preprocess_fwd(const PreprocessArgs& a, const GsrView& v, cudaStream_t stream);

// View-batched variant (preprocess.cu): up to GSR_BATCH_MAX_VIEWS views of the same parameters in one pass.
#define GSR_BATCH_MAX_VIEWS 8
struct FwdViewSlot {                  // per view, in device memory: camera constants + pointers into the view's workspaces
    GsrView v;
    int* radii; float* depths; uint32_t* tiles_touched; float4* recs; uint8_t* clamped;
    uint32_t* block_sums; uint32_t* depth_keys; uint32_t* depth_state; uint2* rects;
};
struct PreprocessBatchArgs {
    int P, n_views;
    const float* means; const float* scales; const float* rotations; const float* opacities; const float* shs;
    int deform_mode; const float* twist_S; const float* twist_theta; const int* body_id;
    float* means_out;                // [P,3] deformed means (one copy for the batch)
    float scale_modifier;
};
int gsr_launch_preprocess_fwd_batched(const PreprocessBatchArgs& a, const FwdViewSlot* d_slots, cudaStream_t stream);
int gsr_launch_mark_visible(int P, const float* means, const GsrView& v, uint8_t* present, cudaStream_t stream);

// ---- binning -------------------------------------------------------------
// Exclusive scan of the per-block sums in place; total -> *d_total.
int gsr_launch_scan_block_sums(uint32_t* block_sums, int num_blocks, uint32_t* d_total, cudaStream_t stream);
// Per-256 sums of tiles_touched taken in DEPTH-SORTED order (order[] from the depth sort).
int gsr_launch_sorted_block_sums(int P, const uint32_t* n_emit /* device: entries of order[] */, const uint32_t* order,
                                 const uint32_t* tiles_touched, uint32_t* block_sums, cudaStream_t stream);
// rasterizer_impl.cu:70-111 duplicateWithKeys, walking the Gaussians in depth order and
// emitting (tile id, Gaussian id) pairs (block-cooperative, balanced).
int gsr_launch_duplicate(int P, const uint32_t* n_emit /* device: entries of order[] */, const uint32_t* order, const int* radii, const uint32_t* tiles_touched,
                         const float4* recs, const uint32_t* block_offsets, uint32_t* tile_ids, uint32_t* vals,
                         int grid_x, int grid_y, GsrSortPlan tile_plan, uint32_t* tile_hist /* pre-zeroed */,
                         cudaStream_t stream);
// rasterizer_impl.cu:116-138 identifyTileRanges (+ the memset of :310) on sorted tile ids.
int gsr_launch_tile_ranges(uint32_t R, const uint32_t* sorted_tile_ids, uint2* ranges, int num_tiles,
                           cudaStream_t stream);
// The reference's 64-bit sorted keys (tile << 32 | depth bits), for parity checks.
int gsr_launch_materialize_keys(uint32_t R, const uint32_t* sorted_tile_ids, const uint32_t* point_list,
                                const float* depths, uint64_t* keys64, cudaStream_t stream);

// ---- depth order of the Gaussians: two-level bucket sort (depth_sort.cu) ----
// state = u32 [GSR_DS_CTRL_WORDS control words | 2^nb bucket counts -> cursors | segment starts]; zeroed by the caller.
#define GSR_DS_CTRL_WORDS 64
enum GsrDepthSortCtrl {
    GSR_DS_NOT_KMIN = 0,        // ~(smallest emitting key), by atomicMax (so that zero = "none yet")
    GSR_DS_KMAX = 1,            // largest emitting key
    GSR_DS_N_EMIT = 4,          // number of Gaussians that emit duplicates = length of the depth order
    GSR_DS_SHIFT = 5, GSR_DS_KMIN = 6,
    GSR_DS_SLOW_SEGMENTS = 7,   // segments that took the CTA-serial radix path (diagnostic)
    // words of the forward that live in the same zeroed block
    GSR_DS_OVERFLOW = 59, GSR_DS_PREFILTERED = 60, GSR_DS_SCAN_TICKET = 61, GSR_DS_NUM_RENDERED = 62, GSR_DS_ERROR = 63
};
size_t gsr_depth_sort_state_bytes(uint32_t P);
// keys[P] (0xffffffff = emits nothing) -> order[0 .. n_emit) = ids by ascending (key, id); srec[i] = (order[i], rects[order[i]]) when given.
// pairs_a / pairs_b: uint2[P] scratch.  minmax_ready: the control words already hold ~min / max (preprocess_fwd wrote them).
int gsr_launch_depth_sort(uint32_t P, const uint32_t* keys, uint32_t* state, uint2* pairs_a, uint2* pairs_b, const uint2* rects,
                          uint32_t* order, uint4* srec, bool minmax_ready, cudaStream_t stream);

// ---- tile binning as one stable counting sort over the depth-ordered Gaussians (tile_sweep.cu) ----
#define GSR_SWEEP_WARPS 8                 // warps (= tile-row stripes) per CTA
#define GSR_SWEEP_ROWS 4                  // tile rows per stripe (lane = row * 8 + column)
#define GSR_SWEEP_MAX_STRIPE_TILES 4096   // counters per warp (16 KB): images up to 16384 px wide
#define GSR_SWEEP_MAX_CHUNKS 1024         // rows of the chunk x tile matrix
struct GsrTileBinPlan {
    int feasible;          // 0: fall back to the radix path
    int chunks;            // rows of the matrix (<= GSR_SWEEP_MAX_CHUNKS); a chunk is ceil(n_emit / chunks) Gaussians
    int stripes, groups, stripe_tiles, num_tiles;
    int scatter_warps, scatter_groups;   // tile_scatter: stripes spread evenly over as few CTAs per chunk as shared memory allows
};
#define GSR_SCATTER_MAX_WARPS 24          // 768 threads
GsrTileBinPlan gsr_make_tile_bin_plan(int grid_x, int grid_y);
size_t gsr_tile_matrix_bytes(int grid_x, int grid_y);
// srec[i] = (id, rect lo, rect hi, 0) of the i-th Gaussian in depth order (written by the depth sort); matrix[chunks][tiles] scratch;
// totals/tile_base [tiles] scratch.  Writes ranges[tiles] and point_list[R].
// n_emit (device) = number of Gaussians that emit duplicates (the depth sort puts them first).
int gsr_launch_tile_binning(int P, const uint32_t* n_emit, const uint4* srec,
                            const GsrTileBinPlan& pl, int grid_x, int grid_y, uint32_t* matrix, uint32_t* totals,
                            uint32_t* tile_base, uint2* ranges, uint32_t* point_list, uint32_t* scan_ticket /* zeroed word */,
                            cudaStream_t stream, uint32_t capacity = 0 /* > 0: point_list holds this many entries; more -> *overflow = 1, empty ranges */,
                            uint32_t* overflow = nullptr /* zeroed device word */);
int gsr_launch_expand_tile_ids(int num_tiles, const uint2* ranges, uint32_t* tile_ids, cudaStream_t stream);

// ---- training-step kernels either side of the rasterizer: loss.cu, adam.cu ----
int gsr_launch_ssim_l1_fwd(const float* img, const float* gt, int H, int W, const float* window11, float lambda,
                           float* dmaps, float* sums8, cudaStream_t stream);
int gsr_launch_ssim_l1_bwd(const float* img, const float* gt, int H, int W, const float* window11, float lambda,
                           const float* dmaps, const float* upstream, float* dL_dimg, cudaStream_t stream);
#define GSR_ADAM_MAX_GROUPS 16
int gsr_launch_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int num_groups,
                    const unsigned long long* begin, const float* step_size, const float* bc2_sqrt,
                    const double* beta1, const double* beta2, const float* eps, cudaStream_t stream);

// ---- onesweep radix sort of (u64|u32 key, u32 value) pairs -----------------

size_t gsr_sort_temp_bytes(uint32_t n, int begin_bit, int end_bit);
// Sorts on bits [begin_bit, end_bit).  Ping-pongs between the a/b buffers and
// returns (through *result_in_b) where the sorted data ended up.
int gsr_launch_sort_pairs(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t n,
                          int begin_bit, int end_bit, void* temp, size_t temp_bytes, int* result_in_b,
                          cudaStream_t stream);

int gsr_launch_sort_pairs32(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t n,
                            int begin_bit, int end_bit, void* temp, size_t temp_bytes, int* result_in_b,
                            cudaStream_t stream, int site = 0 /* profiler label: 1 depth sort, 2 tile sort */,
                            bool hist_ready = false /* caller zeroed temp and already accumulated hist[][] */);

// ---- blending -------------------------------------------------------------
struct BlendFwdArgs {
    const uint2* ranges; const uint32_t* point_list; const float4* recs;
    int W, H, grid_x, grid_y;
    float bg[3];
    float* out_color;        // [3,H,W]
    float* final_T;          // [H*W]
    uint32_t* n_contrib;     // [H*W]
    uint32_t* region_masks;  // [4 (range.x / 32 + tile) ...]: survivors of the per-region cull, one bit per list entry, one word per
                             // (8x8 region, 32-entry batch); written by blend_fwd_v2, read by blend_bwd_v2 (null: not kept)
};
int gsr_launch_blend_fwd(const BlendFwdArgs& a, cudaStream_t stream);
int gsr_launch_blend_fwd_v2(const BlendFwdArgs& a, cudaStream_t stream);      // blend_v2.cu
int gsr_launch_exp_check(float x_max, unsigned long long* out2, cudaStream_t stream);
int gsr_launch_blend_stats(const BlendFwdArgs& a, unsigned long long* out8, cudaStream_t stream);
int gsr_launch_blend_group_stats(const BlendFwdArgs& a, unsigned long long* out32, cudaStream_t stream);   // blend_v2.cu

struct BlendBwdArgs {
    const uint2* ranges; const uint32_t* point_list; const float4* recs;
    int W, H, grid_x, grid_y;
    float bg[3];
    const float* final_T; const uint32_t* n_contrib;
    const float* dL_dpix;    // [3,H,W]
    float4* grad_recs;       // [P,3], zero-initialised by the caller
    const uint32_t* region_masks;   // see BlendFwdArgs (null: the backward repeats the cull)
};
// words of the region-mask array for R list entries over num_tiles tiles (a tile's batches start at range.x / 32 + tile)
static inline size_t gsr_region_mask_words(uint32_t R, int num_tiles) { return 4 * ((size_t)R / 32 + (size_t)num_tiles + 1); }
int gsr_launch_blend_bwd(const BlendBwdArgs& a, cudaStream_t stream);
int gsr_launch_blend_bwd_v2(const BlendBwdArgs& a, cudaStream_t stream);   // blend_v2.cu
int gsr_blend_region_masks_enabled();  // 1 when blend_fwd_v2 writes and blend_bwd_v2 reads the per-region survivor masks
int gsr_blend_bwd_writes_moments();   // 1 when the selected blend backward writes the moment form of the gradient record

// ---- fused per-Gaussian backward ------------------------------------------
struct PreprocessBwdArgs {
    int P;
    const float* means;          // [P,3] un-deformed input means
    const float* means_deformed; // [P,3] or null when deform_mode == NONE (then == means)
    const float* scales; const float* rotations; const float* shs;
    const float* cov3D_precomp; const float* colors_precomp;
    int deform_mode; const float* twist_S; const float* twist_theta; const int* body_id; int num_bodies;
    const int* radii; const uint8_t* clamped;
    const float4* grad_recs;     // [P,3] from blend backward
    int grad_moments;            // 1: records hold the moment form written by blend_v2.cu (see common.cuh)
    const float4* recs;          // [P,3] splat records (conic + opacity; read only when grad_moments)
    float half_W, half_H;        // d(pixel)/d(ndc) = 0.5 W, 0.5 H (applied here in the moment form)
    // outputs (every element written, zeros for culled Gaussians)
    float* dL_dmeans3D;          // [P,3]  w.r.t. the un-deformed means
    float* dL_dmeans2D;          // [P,3]
    float* dL_dopacity;          // [P]
    float* dL_dcolors;           // [P,3]  (meaningful when colors_precomp given)
    float* dL_dcov3D;            // [P,6]
    float* dL_dsh;               // [P,M,3] or null
    float* dL_dscales;           // [P,3] or null
    float* dL_drots;             // [P,4] or null
    float* dL_dtwist_S;          // [P,6] direct, or [B,6] accumulated (pre-zeroed), or null
    float* dL_dtwist_theta;      // [P] or [B] (pre-zeroed), or null
    int acc;                     // GSR_ACC_* bits: those outputs are accumulated into (`+=`) instead of written
};
#define GSR_ACC_MEANS3D 1
#define GSR_ACC_OPACITY 2
#define GSR_ACC_SH 4
#define GSR_ACC_SCALES 8
#define GSR_ACC_ROTS 16
#define GSR_ACC_TWIST 32
int gsr_launch_preprocess_bwd(const PreprocessBwdArgs& a, const GsrView& v, cudaStream_t stream);

// View-batched variant (backward.cu): one pass over the Gaussians for up to GSR_BATCH_MAX_VIEWS views of the same parameters.
struct BwdViewSlot {                  // per view, in device memory
    GsrView v;
    const int* radii; const uint8_t* clamped;
    const float4* grad_recs;         // [P,3] from the view's blend backward
    const float4* recs;              // [P,3] the view's splat records
    float* dL_dmeans2D;              // [P,3] written for every Gaussian (zeros where the view culled it)
};
struct PreprocessBwdBatchArgs {
    int P, n_views;                  // the launch covers Gaussians [first, P)
    int first;
    const float* means; const float* means_deformed; const float* scales; const float* rotations; const float* shs;
    int deform_mode; const float* twist_S; const float* twist_theta; const int* body_id; int num_bodies;
    float scale_modifier; int grad_moments;
    float* dL_dmeans3D; float* dL_dopacity; float* dL_dsh; float* dL_dscales; float* dL_drots;
    float* dL_dtwist_S; float* dL_dtwist_theta;
    int acc;
};
int gsr_launch_preprocess_bwd_batched(const PreprocessBwdBatchArgs& a, const BwdViewSlot* d_slots, cudaStream_t stream);

// ---- standalone SE3 (rigid_body drop-in) ------------------------------------
int gsr_launch_se3_matrices(int N, const float* S, const float* theta, float* T44, cudaStream_t stream);
int gsr_launch_se3_matrices_bwd(int N, const float* S, const float* theta, const float* dT44, float* dS,
                                float* dtheta, cudaStream_t stream);

// ---- kNN -------------------------------------------------------------------
size_t gsr_knn_temp_bytes(int P);
int gsr_launch_knn_dist2(int P, const float* points, float* mean_dist2, void* temp, size_t temp_bytes,
                         cudaStream_t stream);

// ---- deformation-network GEMMs on tcgen05 (mlp_gemm.cu) -----------------------------------
// C[M x N] = epilogue(A[M x K] . B[N x K]^T); every operand enters the tensor core as x = hi + lo (hi a TF32 value) and
// three kind::tf32 MMAs per product give an fp32-grade result.  An operand is given either as ONE fp32 plane (*_lo ==
// null: split in shared memory by the kernel) or as pre-split (hi, lo) planes.  A may be the concatenation of two K
// segments (the skip connection: [embedding | hidden]).  K-major operands only: A row stride ldA, B row stride ldB.
// Outputs: out_lo == null -> one fp32 plane in out_hi; else (hi, lo) planes.  Same for the transposed copy outT_*.
#define GSR_GEMM_RELU_SPLIT 0   // v = relu(acc + bias)
#define GSR_GEMM_SPLIT 1        // v = (acc + bias) [masked]
#define GSR_GEMM_PLAIN 2        // v = (acc + bias) [masked]   (same as 1; kept for callers that name the fp32 output)
#define GSR_GEMM_ATOMIC 3       // out_hi[M x ld_out] += acc   (split-K partial sums, weight gradients)
struct GsrGemmArgs {
    int M, N;
    const float* A0_hi; const float* A0_lo; int K0; long long ldA0;
    const float* A1_hi; const float* A1_lo; int K1; long long ldA1;       // optional second K segment (null: none)
    const float* B_hi; const float* B_lo; long long ldB;                  // [N x (K0pad + K1pad)], K padded to 32 per segment
    int mode;
    int k_splits;                 // > 1: tiles are (row tile, K range) pairs (use with GSR_GEMM_ATOMIC)
    const float* bias;            // [N] or null
    const float* mask_src; int ld_mask;    // v = mask_src[m][n] > 0 ? v : 0  (ReLU backward) or null
    float* out_hi; float* out_lo; int ld_out;
    float* outT_hi; float* outT_lo; long long ld_outT;   // transposed planes [N x ld_outT] or null
    float* colsum;                // [N] += column sums of the stored values (bias gradients) or null
    uint32_t* error_flag;         // device word, set if a pipeline wait timed out
    int mn_major;                 // 1: A0_hi is [K0 x M] and B_hi is [K0 x N] row-major single planes (C += A^T B, atomic mode)
    const char* prof_name;
};
int gsr_launch_mlp_gemm(const GsrGemmArgs& g, cudaStream_t stream);
int gsr_launch_mlp_split(const float* x, long long n, float* hi, float* lo, cudaStream_t stream);
int gsr_launch_mlp_split_transpose(const float* x, int rows, int cols, int ld_in, float* hi, float* lo, int ldT, cudaStream_t stream);
int gsr_launch_mlp_prepare(const float* x, int rows, int cols, long long ld_in, float* hi, float* lo, long long ld_out,
                           float* hiT, float* loT, long long ldT, float* colsum, cudaStream_t stream);
int gsr_launch_mlp_embed(const float* xyz, int P, float* e_hi, float* e_lo, float* eT_hi, float* eT_lo, long long ldT, cudaStream_t stream);
int gsr_launch_mlp_embed_bwd(const float* xyz, int P, const float* de, float* dxyz, int accumulate, cudaStream_t stream);

// ---- densification (densify.cu) -------------------------------------------------------------
int gsr_launch_densify_stats(int P, const float* grad2d, const int* radii, float* accum, float* accum3, float* denom, float* max_radii,
                             cudaStream_t stream);
int gsr_launch_densify_decide(int P, const float* accum, const float* denom, const float* scaling_raw, float grad_threshold,
                              float size_threshold, uint8_t* flags, cudaStream_t stream);
int gsr_launch_densify_split(int n, int N, const float* xyz, const float* scaling_raw, const float* rot_raw, const float* z, float* new_xyz,
                             float* new_scaling, cudaStream_t stream);
int gsr_launch_densify_prune(int P, const float* opacity_raw, const float* scaling_raw, const float* max_radii, float min_opacity,
                             float max_screen_size, float world_size_limit, int use_size, uint8_t* prune, cudaStream_t stream);
int gsr_launch_deform_glue_fwd(int P, const float* heads, const float* xyz, const float* scaling, const float* rotation, const float* f_dc,
                               const float* f_rest, float* means3D, float* scales, float* rotations, float* shs, cudaStream_t stream);
int gsr_launch_deform_glue_bwd(int P, const float* heads, const float* rotation, const float* scales, const float* g_means, const float* g_scales,
                               const float* g_rot, const float* g_shs, float* d_heads, float* d_xyz, float* d_scaling, float* d_rotation,
                               float* d_f_dc, float* d_f_rest, cudaStream_t stream);
