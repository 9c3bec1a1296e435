/* gsr_b200.h - C ABI of libgsr_b200.so: the sm_100a (B200) deformable
 * Gaussian-splatting hot path.
 *
 * This is the drop-in boundary.  Every entry point below takes plain pointers and
 * sizes (device pointers unless the name says host), a cudaStream_t passed as
 * void*, and returns 0 on success or a non-zero code (cudaError_t value, or a
 * negative library code) with a message available from gsr_last_error_string().
 * The library never allocates or frees device memory: the caller owns inputs,
 * outputs and the workspaces whose sizes the *_bytes() queries report.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   gsr_forward_preprocess + gsr_forward_render
 *       <- CudaRasterizer::Rasterizer::forward
 *          submodules/diff-gaussian-rasterization/cuda_rasterizer/rasterizer.h:33-56,
 *          rasterizer_impl.cu:198-336 (bound to Python by rasterize_points.cu:35-115,
 *          ext.cpp:16 "rasterize_gaussians")
 *       <- scene/rigid_body.py:86-93 exp_se3 + gaussian_renderer/__init__.py:92-95
 *          (the SE3 deformation is fused into the preprocess stage)
 *   gsr_backward
 *       <- CudaRasterizer::Rasterizer::backward  rasterizer.h:58-84,
 *          rasterizer_impl.cu:340-434 (rasterize_points.cu:117-196,
 *          ext.cpp:17 "rasterize_gaussians_backward") + autograd of rigid_body.py
 *   gsr_mark_visible
 *       <- CudaRasterizer::Rasterizer::markVisible rasterizer.h:26-31,
 *          rasterizer_impl.cu:141-153 (rasterize_points.cu:198-217, ext.cpp:18)
 *   gsr_knn_dist2
 *       <- distCUDA2  submodules/simple-knn/spatial.cu:15-26, simple_knn.cu:185-221
 *   gsr_exp_se3 / gsr_exp_se3_backward
 *       <- scene/rigid_body.py:86-93 exp_se3 (materialised [N,4,4] transforms)
 *   gsr_sort_pairs
 *       <- cub::DeviceRadixSort::SortPairs as called at rasterizer_impl.cu:303-308
 *          (exported so the sort can be tested and profiled on its own)
 *   gsr_ssim_l1_loss_forward / gsr_ssim_l1_loss_backward
 *       <- utils/loss_utils.py:17-64 (l1_loss, ssim) combined as in train.py:323,529
 *   gsr_adam_step
 *       <- torch.optim.Adam as configured at scene/gaussian_model.py:834-846
 *   gsr_densify_stats / gsr_densify_decide / gsr_densify_split / gsr_densify_prune
 *       <- GaussianModel.add_densification_stats, densify_and_clone, densify_and_split, densify_and_prune
 *          scene/gaussian_model.py:1129-1257 (torch indexing ops in the reference)
 *   gsr_mlp_embed / gsr_mlp_embed_backward
 *       <- Embedder.embed, scene/gaussian_model.py:33-81 (multires 10 positional embedding of the positions)
 *   gsr_mlp_gemm (+ gsr_mlp_split / gsr_mlp_split_transpose operand preparation)
 *       <- the nn.Linear layers of DirectTemporalNeRF.query_time, scene/gaussian_model.py:285-293 (forward) and
 *          their autograd (input and weight gradients): cuBLAS SGEMM in the reference, tcgen05 tensor cores here
 */
#ifndef GSR_B200_H
#define GSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Host-side description of one view.  Field meaning follows
 * GaussianRasterizationSettings (diff_gaussian_rasterization/__init__.py:157-169);
 * matrices use the reference's memory convention (16 floats, true row i =
 * m[i], m[i+4], m[i+8], m[i+12]).  All arrays here are HOST values. */
typedef struct gsr_view {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float bg[3];
    float scale_modifier;
    float viewmatrix[16];
    float projmatrix[16];
    int32_t sh_degree;     /* active degree D */
    float campos[3];
    int32_t prefiltered;
    int32_t debug;
} gsr_view;

/* Optional SE3 deformation fused into preprocess.  mode: 0 none,
 * 1 per-Gaussian twists (S[P,6], theta[P]), 2 rigid bodies (body_id[P] int32,
 * S[B,6], theta[B]).  S = (w, v) as in rigid_body.exp_se3. */
typedef struct gsr_deform {
    int32_t mode;
    int32_t num_bodies;
    const float* S;
    const float* theta;
    const int32_t* body_id;
} gsr_deform;

const char* gsr_last_error_string(void);
int gsr_version(void);

/* Launch accounting and per-kernel timing.  gsr_launch_count: kernels launched by
 * this library since the last reset.  With profiling enabled every kernel launch is
 * bracketed by CUDA events on its stream; gsr_profile_dump synchronises those events
 * and writes one "name launches total_ms" line per kernel into out (returns the
 * number of kernels).  Single-threaded use. */
void gsr_profile_enable(int on);
unsigned long long gsr_launch_count(int reset);
int gsr_profile_dump(char* out, size_t cap);

/* ---- workspace sizes ---------------------------------------------------- */
size_t gsr_geom_bytes(int P);                       /* per-Gaussian state of one view        */
size_t gsr_image_bytes(int width, int height);      /* final_T, n_contrib, tile ranges       */
size_t gsr_binning_bytes(uint32_t num_rendered, int width, int height); /* keys/values x2 + sort temp */
size_t gsr_grad_bytes(int P);                       /* blend-backward gradient records       */
/* Byte offsets of the geometry-workspace sub-arrays, for tests and tools:
 * out[0]=depths f32[P], out[1]=tiles_touched u32[P], out[2]=splat records float4[3P]
 * (x,y,conic.x,conic.y | conic.z,opacity,r,g | b,cut,0,0), out[3]=clamped u8[P],
 * out[4]=depth order u32[P] (Gaussian ids, front to back, after gsr_forward_render),
 * out[5]=cov3D f32[6P] (debug only) */
void gsr_geom_layout(int P, size_t out[6]);
/* out[0]=final_T f32[WH], out[1]=n_contrib u32[WH], out[2]=ranges uint2[tiles] */
void gsr_image_layout(int width, int height, size_t out[3]);
/* out[0]=reference-format sorted keys u64[R] (only when materialize_keys was set),
 * out[1]=point_list u32[R] (sorted Gaussian ids), out[2]=sorted tile ids u32[R], out[3]=scratch */
void gsr_binning_layout(uint32_t num_rendered, int width, int height, size_t out[4]);

/* ---- forward -------------------------------------------------------------
 * Stage 1: deform + project + covariance + SH (num_rendered is summed by the same kernel).  Writes radii[P]
 * (int32), optional means_out[P,3] (required when deform->mode != 0), fills the
 * geometry workspace, copies num_rendered to *host_num_rendered (pinned host
 * memory recommended) and synchronises the stream - the one host round trip of
 * the forward, as in the reference (rasterizer_impl.cu:281).
 * Optional inputs are NULL when absent (exactly one of shs/colors_precomp and one
 * of (scales,rotations)/cov3D_precomp must be given).  M = SH coefficients stored
 * per Gaussian. */
int gsr_forward_preprocess(const gsr_view* view, int P, int M,
                           const float* means3D, const float* scales, const float* rotations,
                           const float* opacities, const float* shs,
                           const float* cov3D_precomp, const float* colors_precomp,
                           const gsr_deform* deform, float* means_out,
                           int32_t* radii, void* geom_ws, size_t geom_bytes,
                           uint32_t* host_num_rendered, int debug_dump_cov3D, void* stream);

/* Stage 2: depth-order the Gaussians (two-level bucket sort, csrc/depth_sort.cu), then one stable counting sort of their tile duplicates into the
 * per-tile lists (= the reference's sorted point_list and ranges; radix fallback: duplicate + onesweep by tile id +
 * tile ranges), then blend.  out_color[3,H,W].  materialize_keys != 0 also
 * writes the reference-format sorted 64-bit keys (tile << 32 | depth bits) for parity checks. */
int gsr_forward_render(const gsr_view* view, int P, uint32_t num_rendered,
                       const int32_t* radii, void* geom_ws, void* binning_ws, size_t binning_bytes,
                       void* image_ws, float* out_color, int materialize_keys, void* stream);

/* Sync-free forward.  gsr_forward_preprocess + gsr_forward_render cost one device->host read of num_rendered and a stream
 * synchronisation between them (as does the reference, rasterizer_impl.cu:281), because the duplicate list is sized from
 * it.  These two calls size the list from a CAPACITY the caller chooses instead (e.g. the previous step's num_rendered
 * plus a margin): nothing is read back in between, the whole forward is enqueued at once and can be captured in a CUDA
 * graph.  If a view emits more duplicates than the capacity, the device raises a flag, writes no list entry, empties every
 * tile (the image is then the background) and host_status4 tells: 4 words of PINNED host memory filled asynchronously at
 * the end of the call: [0] overflow flag, [1] prefiltered violation, [2] unused, [3] num_rendered.  The binning workspace
 * must hold gsr_binning_bytes(R_capacity, W, H); gsr_backward takes R = R_capacity.  Images that need the radix binning
 * path (wider than 16 384 px) are refused. */
int gsr_forward_preprocess_async(const gsr_view* view, int P, int M, const float* means3D, const float* scales,
                                 const float* rotations, const float* opacities, const float* shs,
                                 const float* cov3D_precomp, const float* colors_precomp, const gsr_deform* deform,
                                 float* means_out, int32_t* radii, void* geom_ws, size_t geom_bytes, void* stream);
int gsr_forward_render_capacity(const gsr_view* view, int P, uint32_t R_capacity, const int32_t* radii, void* geom_ws,
                                void* binning_ws, size_t binning_bytes, void* image_ws, float* out_color,
                                uint32_t* host_status4, void* stream);

/* ---- backward -------------------------------------------------------------
 * dL_dout_color[3,H,W] -> gradients.  Every output element is written (no
 * pre-zeroing needed) except dL_dtwist_* in rigid-body mode, which are
 * ACCUMULATED into and must be zeroed by the caller.  Outputs may be NULL when
 * the corresponding input was absent (dL_dsh, dL_dscales, dL_drots, dL_dtwist_*; dL_dcolors and
 * dL_dcov3D are only needed with precomputed colours / covariances).
 * accumulate_mask: for the outputs whose bit is set (1 means3D, 2 opacity, 4 sh, 8 scales,
 * 16 rotations, 32 per-Gaussian twists) the buffer is the caller's running gradient and is
 * accumulated into (`+=`; Gaussians culled in this view are not touched) - this is how
 * view-batched training sums gradients over views without a separate accumulation pass. */
int gsr_backward(const gsr_view* view, int P, int M, uint32_t num_rendered,
                 const float* means3D, const float* means_deformed,
                 const float* scales, const float* rotations, const float* shs,
                 const float* cov3D_precomp, const float* colors_precomp,
                 const gsr_deform* deform, const int32_t* radii,
                 const void* geom_ws, const void* binning_ws, const void* image_ws,
                 void* grad_ws, const float* dL_dout_color,
                 float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dopacity, float* dL_dcolors,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drots,
                 float* dL_dtwist_S, float* dL_dtwist_theta, int accumulate_mask, void* stream);

/* Debug/measurement aid: replays the blend loop of a finished forward and writes 8
 * workload counters (device u64[8]): staged entries, tile-cull survivors,
 * (warp,entry) iterations, ... with a lane passing the power test, ... with a lane
 * blending, (pixel,entry) pairs evaluated, pairs blended, total list entries. */
int gsr_debug_blend_stats(const gsr_view* view, int P, uint32_t num_rendered, const void* geom_ws,
                          const void* binning_ws, const void* image_ws, unsigned long long* out8, void* stream);

/* ---- view-batched forward preprocess ----------------------------------------------------------------------------------
 * gsr_forward_preprocess for up to all views of a step at once (they share the parameters): every parameter record is read
 * once, SE3 and cov3D are evaluated once, and each view's records go to that view's geometry workspace - byte for byte
 * what gsr_forward_preprocess writes there.  Then gsr_forward_render / gsr_forward_render_capacity per view as usual.
 * Requires scales + rotations, SH colours with M = 16 and 32-byte aligned shs, one scale_modifier / sh_degree for the
 * batch, no prefiltered / debug views.  means_out ([P,3], deform modes) is one copy for the batch.  Slots as for the
 * batched backward: fill on the host (gsr_forward_batched_fill_slots), copy to the device in stream order.
 * gsr_read_num_rendered: asynchronous read-back of a view's num_rendered (the exact-size path needs it on the host). */
typedef struct gsr_view_fwd {
    const gsr_view* view;
    int32_t* radii;            /* [P] out */
    void* geom_ws;             /* gsr_geom_bytes(P) */
} gsr_view_fwd;
size_t gsr_forward_batched_slots_bytes(int n_views);
int gsr_forward_batched_fill_slots(int n_views, const gsr_view_fwd* views, int P, int M, void* slots_host, size_t slots_bytes);
int gsr_forward_preprocess_batched(int n_views, const gsr_view_fwd* views, const void* slots_device, int P, int M,
                                   const float* means3D, const float* scales, const float* rotations, const float* opacities,
                                   const float* shs, const gsr_deform* deform, float* means_out, size_t geom_bytes, void* stream);
int gsr_read_num_rendered(const void* geom_ws, int P, uint32_t* host_num_rendered, void* stream);

/* ---- view-batched backward (training steps that render several views of the SAME parameters) -----------------------
 * gsr_backward = renderCUDA backward + the per-Gaussian chain rule (computeCov2DCUDA / preprocessCUDA backward,
 * cuda_rasterizer/backward.cu:144-396, + the SE3 autograd of scene/rigid_body.py).  The second half reads every parameter
 * record and read-modify-writes the running gradient once per view.  Split in two, the first half runs per view
 * (gsr_backward_blend: fills the view's grad_ws) and the second ONCE for up to all views of the step
 * (gsr_backward_gaussians_batched): parameters read once, gradients updated once; cov3D and SE3 backward are linear in
 * their upstream gradients, which are summed over the views first.  Requires scales + rotations, SH colours with M = 16
 * and 32-byte aligned shs / dL_dsh, one scale_modifier for the batch; results equal the per-view path up to fp32
 * summation order.  dL_dmeans2D is per view.  A call covers the Gaussians [first, first + count) (all pointers are those of
 * the whole arrays), so that a caller can start the all-reduce of one range of the gradients while the next is computed.
 *   gsr_backward_batched_fill_slots   host-only: packs the views' constants and workspace pointers into slots_host
 *                                     (gsr_backward_batched_slots_bytes(n) bytes, e.g. pinned memory) - copy them to the
 *                                     device in stream order and pass the device copy as slots_device. */
typedef struct gsr_view_grads {
    const gsr_view* view;
    const int32_t* radii;      /* of the view's forward */
    const void* geom_ws;       /* the view's geometry workspace */
    const void* grad_ws;       /* filled by gsr_backward_blend for this view */
    float* dL_dmeans2D;        /* [P,3], written for every Gaussian */
} gsr_view_grads;
int gsr_backward_blend(const gsr_view* view, int P, uint32_t num_rendered, const void* geom_ws, const void* binning_ws,
                       const void* image_ws, void* grad_ws, const float* dL_dout_color, void* stream);
size_t gsr_backward_batched_slots_bytes(int n_views);
int gsr_backward_batched_fill_slots(int n_views, const gsr_view_grads* views, int P, int M, void* slots_host, size_t slots_bytes);
int gsr_backward_gaussians_batched(int n_views, const void* slots_device, float scale_modifier, int P, int first, int count, int M, const float* means3D,
                                   const float* means_deformed, const float* scales, const float* rotations, const float* shs,
                                   const gsr_deform* deform, float* dL_dmeans3D, float* dL_dopacity, float* dL_dsh,
                                   float* dL_dscales, float* dL_drots, float* dL_dtwist_S, float* dL_dtwist_theta,
                                   int accumulate_mask, void* stream);

/* The depth sort of the forward on its own (csrc/depth_sort.cu; replaces the depth half of the reference's
 * cub::DeviceRadixSort::SortPairs, rasterizer_impl.cu:303-308): keys = device u32[P], 0xffffffff marks an entry that is
 * left out; order[0 .. n) receives the indices of the other entries by ascending (key, index) - the order a stable sort
 * by key gives.  ws = device scratch of gsr_depth_order_ws_bytes(P); info2 (device u32[2], may be NULL) receives n and the
 * number of segments that needed the skew path. */
size_t gsr_depth_order_ws_bytes(int P);
int gsr_depth_order(const uint32_t* keys, int P, void* ws, size_t ws_bytes, uint32_t* order, uint32_t* info2, void* stream);

/* Debug/measurement aid: replays blend_fwd_v2's walk per 8x8 region and counts the loop iterations that would remain if
 * the warp's lanes were split into groups (1 x 8x8, 2 x 8x4, 4 x 4x4, 4 x 8x2, 2 x 4x8, 8 x 4x2 pixels) walking separately
 * culled lists; out32 = device u64[32] (layout in csrc/blend_v2.cu: blend_group_stats_kernel). */
int gsr_debug_blend_group_stats(const gsr_view* view, int P, uint32_t num_rendered, const void* geom_ws,
                                const void* binning_ws, const void* image_ws, unsigned long long* out32, void* stream);

/* Debug/parity aid: compares the library's restatements of CUDA's expf (scalar and packed
 * FP32x2, csrc/f32x2.cuh) with expf itself on EVERY float in [-x_max, -0]; writes the two
 * mismatch counts to out2 (device u64[2]).  Both must be 0 for the blend to be bit-exact. */
int gsr_debug_exp_check(float x_max, unsigned long long* out2, void* stream);

/* ---- the training step either side of the rasterizer -------------------------
 * Loss of train.py:323,529: (1 - lambda) * mean|image - gt| + lambda * (1 - ssim(image, gt)) with
 * utils/loss_utils.py:36-64's SSIM (11x11 Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2).
 * window11 (HOST, 11 floats) is loss_utils.gaussian(11, 1.5); image/gt are device float [3,H,W].
 * forward writes dmaps (device float [3][3][H][W], kept for backward) and out8 (device float[8]:
 * [3] = loss, [4] = L1 mean, [5] = SSIM mean; [0..2] scratch).  backward writes dL/dimage scaled by the
 * device scalar *upstream (NULL = 1). */
int gsr_ssim_l1_loss_forward(const float* image, const float* gt, int channels, int height, int width,
                             const float* window11, float lambda_dssim, float* dmaps, float* out8, void* stream);
int gsr_ssim_l1_loss_backward(const float* image, const float* gt, int channels, int height, int width,
                              const float* window11, float lambda_dssim, const float* dmaps, const float* upstream,
                              float* dL_dimage, void* stream);

/* torch.optim.Adam (no amsgrad / weight decay; scene/gaussian_model.py:846, eps = 1e-15) over ONE flat fp32
 * buffer: group k covers elements [group_begin[k], group_begin[k+1]) (HOST arrays, <= 16 groups) with
 * step_size[k] = lr_k / (1 - beta1^t) and bias_correction2_sqrt[k] = sqrt(1 - beta2^t) (computed in double by the
 * caller, as torch does); the betas are DOUBLES so that 1 - beta rounds to float the way torch rounds it.  params,
 * grads and both moments are device pointers, 16-byte aligned. */
int gsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int num_groups,
                  const unsigned long long* group_begin, const float* step_size, const float* bias_correction2_sqrt,
                  const double* beta1, const double* beta2, const float* eps, void* stream);

/* ---- the rest of the reference's operator surface -------------------------- */
int gsr_mark_visible(const gsr_view* view, int P, const float* means3D, uint8_t* present, void* stream);

/* Deformation-network linear layers on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
 * TMEM, operands staged by TMA), with an fp32-grade result: every operand enters the tensor core as x = hi + lo with hi
 * a TF32 value, and each product is evaluated as hi.hi + lo.hi + hi.lo.  An operand is given as ONE fp32 plane (its *_lo
 * pointer NULL: the kernel splits the tile in shared memory - for the P-sized activations, which then cross HBM once) or
 * as planes pre-split by gsr_mlp_split (the weights).  Outputs likewise: out_lo / outT_lo NULL = one fp32 plane.
 *   C[M x N] = epilogue(A[M x K] . B[N x K]^T),  N <= 256,  all operands K-major (row stride ldA / ldB floats,
 *   multiples of 4, 16-byte aligned planes).  A may be two K segments ([A0 | A1], the network's skip connection);
 *   each segment's K is padded to a multiple of 16 in B's column numbering (segment 1 starts at 16 ceil(K0 / 16)).
 * mode 0: relu(acc + bias);  1 (= 2): (acc + bias), optionally zeroed where mask_src <= 0;
 *      3: out_hi += acc with atomics (k_splits > 1: split over K, weight gradients).
 * outT_*: the same values transposed ([N x ld_outT]); colsum[n] += column sums (bias gradients). */
typedef struct gsr_gemm {
    int32_t M, N;
    const float* A0_hi; const float* A0_lo; int32_t K0; int64_t ldA0;
    const float* A1_hi; const float* A1_lo; int32_t K1; int64_t ldA1;
    const float* B_hi; const float* B_lo; int64_t ldB;
    int32_t mode;
    int32_t k_splits;
    const float* bias;
    const float* mask_src; int32_t ld_mask;
    float* out_hi; float* out_lo; int32_t ld_out;
    float* outT_hi; float* outT_lo; int64_t ld_outT;
    float* colsum;
    uint32_t* error_flag;
    int32_t mn_major;      /* 1: C[M x N] += A^T B with A0_hi [K0 x M] and B_hi [K0 x N] ROW-major single planes (the reduction
                              runs over rows = points: weight gradients straight from the row-major activations); mode 3 only */
} gsr_gemm;
int gsr_mlp_gemm(const gsr_gemm* g, void* stream);
int gsr_mlp_split(const float* x, int64_t n, float* hi, float* lo, void* stream);
/* x [rows x cols] (row stride ld_in) -> planes of x^T [cols x ldT] */
int gsr_mlp_split_transpose(const float* x, int rows, int cols, int ld_in, float* hi, float* lo, int ldT, void* stream);
/* x [rows x cols] (row stride ld_in) -> any of: row-major planes [rows x ld_out], transposed planes [cols x ldT],
 * colsum[c] += column sums.  Prepares a gradient that arrives from autograd as a GEMM operand. */
int gsr_mlp_prepare(const float* x, int rows, int cols, int64_t ld_in, float* hi, float* lo, int64_t ld_out, float* hiT,
                    float* loT, int64_t ldT, float* colsum, void* stream);
/* positions [P,3] -> embedding [P x 64] (63 values, column 63 zero; one plane if e_lo is NULL) and, optionally, transposed [64 x ldT] */
int gsr_mlp_embed(const float* xyz, int P, float* e_hi, float* e_lo, float* eT_hi, float* eT_lo, int64_t ldT, void* stream);
int gsr_mlp_embed_backward(const float* xyz, int P, const float* d_embed /*[P x 64]*/, float* dxyz, int accumulate, void* stream);

/* Densification (scene/gaussian_model.py:1129-1257, train.py:610-648): per-point statistics and decisions.  All arrays are
 * the reference's own tensors (pre-activation log-scales / opacity logits / un-normalised quaternions).
 *   gsr_densify_stats   add_densification_stats + the max_radii2D update for the points a view saw (radii > 0)
 *   gsr_densify_decide  flags[i] bit 0 = clone, bit 1 = split  (grads = accum / denom, NaN -> 0; size_threshold = percent_dense * extent)
 *   gsr_densify_split   the n selected points x N samples: new positions R(q)(z exp(scaling)) + xyz and new log-scales
 *                       log(exp(scaling) / (0.8 N)); `normals` = [n N, 3] standard normals, row r belongs to point r % n
 *   gsr_densify_prune   prune[i] = sigmoid(opacity) < min_opacity, or (use_size) max_radii2D > max_screen_size or
 *                       max(exp(scaling)) > world_size_limit */
int gsr_densify_stats(int P, const float* viewspace_grad, const int32_t* radii, float* xyz_gradient_accum, float* xyz_gradient_accum_3vec,
                      float* denom, float* max_radii2D, void* stream);
int gsr_densify_decide(int P, const float* xyz_gradient_accum, const float* denom, const float* scaling_raw, float grad_threshold,
                       float size_threshold, uint8_t* flags, void* stream);
int gsr_densify_split(int n, int N, const float* xyz, const float* scaling_raw, const float* rotation_raw, const float* normals, float* new_xyz,
                      float* new_scaling, void* stream);
int gsr_densify_prune(int P, const float* opacity_raw, const float* scaling_raw, const float* max_radii2D, float min_opacity,
                      float max_screen_size, float world_size_limit, int use_size, uint8_t* prune, void* stream);

/* Activation glue between the deformation network and the rasterizer (gaussian_renderer/__init__.py:79,116,122,140), one pass:
 *   means3D = xyz + dx;  scales = exp(scaling + dscale);  rotations = normalize(rotation + drot);  shs = [f_dc | f_rest] + dshs
 * heads = the network's output [P x 64] (dx 0-2, dscale 3-5, drot 6-9, dshs 10-57).  The backward takes the gradients w.r.t.
 * the four outputs (any may be NULL = zero) and writes d_heads [P x 64] and, where non-NULL, the parameters' gradients. */
int gsr_deform_glue_forward(int P, const float* heads, const float* xyz, const float* scaling, const float* rotation, const float* f_dc,
                            const float* f_rest, float* means3D, float* scales, float* rotations, float* shs, void* stream);
int gsr_deform_glue_backward(int P, const float* heads, const float* rotation, const float* scales, const float* g_means, const float* g_scales,
                             const float* g_rotations, const float* g_shs, float* d_heads, float* d_xyz, float* d_scaling, float* d_rotation,
                             float* d_f_dc, float* d_f_rest, void* stream);

size_t gsr_knn_bytes(int P);
int gsr_knn_dist2(int P, const float* points, float* mean_dist2, void* temp, size_t temp_bytes, void* stream);

int gsr_exp_se3(int N, const float* S, const float* theta, float* T44, void* stream);
int gsr_exp_se3_backward(int N, const float* S, const float* theta, const float* dT44,
                         float* dS, float* dtheta, void* stream);

size_t gsr_sort_bytes(uint32_t n, int begin_bit, int end_bit);
int gsr_sort_pairs(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                   uint32_t n, int begin_bit, int end_bit, void* temp, size_t temp_bytes,
                   int* result_in_b, void* stream);
int gsr_sort_pairs32(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                     uint32_t n, int begin_bit, int end_bit, void* temp, size_t temp_bytes,
                     int* result_in_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H */
