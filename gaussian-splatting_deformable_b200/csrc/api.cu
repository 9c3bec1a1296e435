// extern "C" entry points of libgsr_b200.so (declared in include/gsr_b200.h).
// Host-side sequencing only: workspace carving and kernel launches on the
// caller's stream.  Mirrors CudaRasterizer::Rasterizer::{forward,backward,
// markVisible} (rasterizer_impl.cu:141-153,198-434) behind a C ABI.
#include "../../include/gsr_b200.h"
#include "kernels.cuh"
#include <atomic>
#include <mutex>
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

int gsr_set_error(cudaError_t e, const char* what, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return (int)e;
}
int gsr_set_error_msg(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "gsr error %d: %s", code, msg);
    return code;
}

// ---- launch accounting / profiler -------------------------------------------
namespace {
struct ProfEntry { const char* name; unsigned long long count; double ms; };
struct ProfPending { int slot; cudaEvent_t e0, e1; };
// launches may come from several host threads (one per stream): the counter is atomic, the profiler tables sit behind a mutex
std::atomic<bool> g_prof_on{false};
std::atomic<unsigned long long> g_launches{0};
std::mutex g_prof_mu;
ProfEntry g_prof[64];
int g_prof_n = 0;
ProfPending g_pending[4096];
int g_pending_n = 0;
int prof_slot(const char* name) {
    for (int i = 0; i < g_prof_n; i++) if (g_prof[i].name == name || !strcmp(g_prof[i].name, name)) return i;
    if (g_prof_n >= 64) return 63;
    g_prof[g_prof_n] = ProfEntry{name, 0ull, 0.0};
    return g_prof_n++;
}
void prof_collect() {
    for (int i = 0; i < g_pending_n; i++) {
        float ms = 0.f;
        cudaEventSynchronize(g_pending[i].e1);
        cudaEventElapsedTime(&ms, g_pending[i].e0, g_pending[i].e1);
        g_prof[g_pending[i].slot].ms += ms;
        cudaEventDestroy(g_pending[i].e0);
        cudaEventDestroy(g_pending[i].e1);
    }
    g_pending_n = 0;
}
}  // namespace

GsrProfScope::GsrProfScope(const char* name, cudaStream_t s) : slot(-1), stream(s) {
    g_launches++;
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (g_pending_n >= 4095) prof_collect();
    slot = prof_slot(name);
    g_prof[slot].count++;
    ProfPending& p = g_pending[g_pending_n];
    p.slot = slot;
    cudaEventCreate(&p.e0);
    cudaEventCreate(&p.e1);
    cudaEventRecord(p.e0, stream);
    pending = g_pending_n++;
}
GsrProfScope::~GsrProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (pending < g_pending_n && g_pending[pending].slot == slot) cudaEventRecord(g_pending[pending].e1, stream);
}

namespace {

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct GeomLayout {
    size_t in_b_order() const { return dorder; }
    size_t depths, tiles, recs, clamped, cov3d, block_sums, total, dkeys, dpairs_a, dpairs_b, dorder,
                    dstate, dstate_bytes, sblock_sums, rects, srec, bytes; };
GeomLayout geom_layout(int P) {
    GeomLayout L{};
    size_t o = 0;
    const size_t p = (size_t)(P > 0 ? P : 0);
    L.depths = o; o += al(4 * p);
    L.tiles = o; o += al(4 * p);
    L.recs = o; o += al(48 * p);
    L.clamped = o; o += al(p);
    L.cov3d = o; o += al(24 * p);
    L.block_sums = o; o += al(4 * ((p + 255) / 256 + 1));
    L.total = o; o += 256;
    // depth sort (depth_sort.cu): keys, two pair buffers (bucketed pairs; scratch of the skew path), the order, the state
    L.dkeys = o; o += al(4 * p);
    L.dpairs_a = o; o += al(8 * p);
    L.dpairs_b = o; o += al(8 * p);
    L.dorder = o; o += al(4 * p);
    L.dstate = o; L.dstate_bytes = gsr_depth_sort_state_bytes((uint32_t)p); o += al(L.dstate_bytes);
    L.sblock_sums = o; o += al(4 * ((p + 255) / 256 + 1));
    L.rects = o; o += al(8 * p);
    L.srec = o; o += al(16 * p);
    L.bytes = o;
    return L;
}
struct ImageLayout { size_t final_T, n_contrib, ranges, bytes; };
ImageLayout image_layout(int W, int H) {
    ImageLayout L{};
    const size_t n = (size_t)W * H;
    const size_t tiles = (size_t)((W + 15) / 16) * ((H + 15) / 16);
    size_t o = 0;
    L.final_T = o; o += al(4 * n);
    L.n_contrib = o; o += al(4 * n);
    L.ranges = o; o += al(8 * tiles);
    L.bytes = o;
    return L;
}
// Bits of the tile id, as the reference's getHigherMsb (rasterizer_impl.cu:35-50).
int tile_bits(uint32_t n) {
    int msb = 16, step = 16;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}
struct BinLayout { size_t tkeys_a, tkeys_b, vals_a, vals_b, sort_temp, sort_temp_bytes, keys64, matrix, totals, tile_base, masks, bytes;
                   int tile_bits; int passes; int sweep; };
BinLayout bin_layout(uint32_t R, int W, int H) {
    BinLayout L{};
    const uint32_t tiles = (uint32_t)((W + 15) / 16) * ((H + 15) / 16);
    L.tile_bits = tile_bits(tiles);
    L.passes = (L.tile_bits + 7) / 8;
    size_t o = 0;
    L.tkeys_a = o; o += al(4 * (size_t)R);
    L.tkeys_b = o; o += al(4 * (size_t)R);
    L.vals_a = o; o += al(4 * (size_t)R);
    L.vals_b = o; o += al(4 * (size_t)R);
    L.sort_temp = o; L.sort_temp_bytes = gsr_sort_temp_bytes(R, 0, L.tile_bits); o += al(L.sort_temp_bytes);
    L.keys64 = o; o += al(8 * (size_t)R);          // reference-format keys, filled on request only
    // counting-sort path (tile_sweep.cu): point_list = vals_a, sorted tile ids (debug) = tkeys_a
    const int gx = (W + 15) / 16, gy = (H + 15) / 16;
    L.sweep = gsr_make_tile_bin_plan(gx, gy).feasible;
    L.matrix = o; o += al(gsr_tile_matrix_bytes(gx, gy));
    L.totals = o; o += al(4 * (size_t)tiles);
    L.tile_base = o; o += al(4 * (size_t)tiles);
    L.masks = o; o += al(4 * gsr_region_mask_words(R, (int)tiles));     // survivors of the per-region cull (blend_fwd_v2 -> blend_bwd_v2)
    L.bytes = o + 256;
    return L;
}

int fill_view(const gsr_view* s, int M, GsrView& v) {
    if (!s) return gsr_set_error_msg(-1, "view is NULL");
    memcpy(v.view, s->viewmatrix, sizeof(v.view));
    memcpy(v.proj, s->projmatrix, sizeof(v.proj));
    memcpy(v.campos, s->campos, sizeof(v.campos));
    v.tan_fovx = s->tanfovx; v.tan_fovy = s->tanfovy;
    v.W = s->image_width; v.H = s->image_height;
    // rasterizer_impl.cu:222-223
    v.focal_y = v.H / (2.0f * v.tan_fovy);
    v.focal_x = v.W / (2.0f * v.tan_fovx);
    v.scale_modifier = s->scale_modifier;
    v.grid_x = (v.W + GSR_TILE - 1) / GSR_TILE;
    v.grid_y = (v.H + GSR_TILE - 1) / GSR_TILE;
    v.sh_degree = s->sh_degree;
    v.sh_coeffs = M;
    if (v.W <= 0 || v.H <= 0) return gsr_set_error_msg(-1, "image size must be positive");
    if (v.grid_x >= 65536 || v.grid_y >= 65536) return gsr_set_error_msg(-1, "image too large");
    return 0;
}

}  // namespace

extern "C" {

const char* gsr_last_error_string(void) { return g_err; }

void gsr_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    prof_collect();
    g_prof_on = on != 0;
    if (on) { g_prof_n = 0; }
}
unsigned long long gsr_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}
int gsr_profile_dump(char* out, size_t cap) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    prof_collect();
    size_t o = 0;
    if (cap) out[0] = 0;
    for (int i = 0; i < g_prof_n; i++) {
        const int w = snprintf(out + o, cap > o ? cap - o : 0, "%s %llu %.6f\n", g_prof[i].name, g_prof[i].count, g_prof[i].ms);
        if (w < 0 || (size_t)w >= (cap > o ? cap - o : 0)) break;
        o += (size_t)w;
    }
    return g_prof_n;
}
int gsr_version(void) { return 100; }

size_t gsr_geom_bytes(int P) { return geom_layout(P).bytes; }
size_t gsr_image_bytes(int W, int H) { return image_layout(W, H).bytes; }
size_t gsr_binning_bytes(uint32_t R, int W, int H) { return bin_layout(R, W, H).bytes; }
size_t gsr_grad_bytes(int P) { return al(48 * (size_t)(P > 0 ? P : 0)) + 256; }

void gsr_geom_layout(int P, size_t out[6]) {
    const GeomLayout L = geom_layout(P);
    out[0] = L.depths; out[1] = L.tiles; out[2] = L.recs; out[3] = L.clamped; out[4] = L.in_b_order(); out[5] = L.cov3d;
}
void gsr_image_layout(int W, int H, size_t out[3]) {
    const ImageLayout L = image_layout(W, H);
    out[0] = L.final_T; out[1] = L.n_contrib; out[2] = L.ranges;
}
void gsr_binning_layout(uint32_t R, int W, int H, size_t out[4]) {
    const BinLayout L = bin_layout(R, W, H);
    const bool in_b = !L.sweep && (L.passes & 1) != 0;
    out[0] = L.keys64;
    out[1] = in_b ? L.vals_b : L.vals_a;
    out[2] = in_b ? L.tkeys_b : L.tkeys_a;
    out[3] = in_b ? L.tkeys_a : L.tkeys_b;
}

static int forward_preprocess_impl(const gsr_view* view, int P, int M, const float* means3D, const float* scales,
                                   const float* rotations, const float* opacities, const float* shs,
                                   const float* cov3D_precomp, const float* colors_precomp, const gsr_deform* deform,
                                   float* means_out, int32_t* radii, void* geom_ws, size_t geom_bytes,
                                   uint32_t* host_num_rendered, int debug_dump_cov3D, void* stream_, bool sync) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (host_num_rendered && sync) *host_num_rendered = 0;
    if (P < 0) return gsr_set_error_msg(-1, "P must be >= 0");
    if (P == 0) return 0;
    GsrView v;
    if (int rc = fill_view(view, M, v)) return rc;
    if (!means3D || !opacities || !radii || !geom_ws || (sync && !host_num_rendered))
        return gsr_set_error_msg(-1, "forward_preprocess: required pointer is NULL");
    if ((shs == nullptr) == (colors_precomp == nullptr))
        return gsr_set_error_msg(-1, "provide exactly one of SHs or precomputed colors");
    if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))
        return gsr_set_error_msg(-1, "provide exactly one of scale/rotation pair or precomputed 3D covariance");
    if (shs && (M <= 0 || M > 16 || (v.sh_degree + 1) * (v.sh_degree + 1) > M || v.sh_degree > 3 || v.sh_degree < 0))
        return gsr_set_error_msg(-1, "SH layout: need 0 <= degree <= 3 and (degree+1)^2 <= M <= 16");
    const GeomLayout L = geom_layout(P);
    if (geom_bytes < L.bytes) return gsr_set_error_msg(-3, "geometry workspace too small");
    char* ws = reinterpret_cast<char*>(geom_ws);
    PreprocessArgs a{};
    a.P = P; a.means = means3D; a.scales = scales; a.rotations = rotations; a.opacities = opacities; a.shs = shs;
    a.cov3D_precomp = cov3D_precomp; a.colors_precomp = colors_precomp;
    a.deform_mode = deform ? deform->mode : GSR_DEFORM_NONE;
    if (a.deform_mode != GSR_DEFORM_NONE) {
        if (!deform->S || !deform->theta || !means_out) return gsr_set_error_msg(-1, "deform: S, theta and means_out required");
        if (a.deform_mode == GSR_DEFORM_RIGID_BODIES && !deform->body_id) return gsr_set_error_msg(-1, "deform: body_id required");
        a.twist_S = deform->S; a.twist_theta = deform->theta; a.body_id = deform->body_id;
    }
    a.means_out = means_out;
    a.radii = radii;
    a.depths = reinterpret_cast<float*>(ws + L.depths);
    a.tiles_touched = reinterpret_cast<uint32_t*>(ws + L.tiles);
    a.recs = reinterpret_cast<float4*>(ws + L.recs);
    a.clamped = reinterpret_cast<uint8_t*>(ws + L.clamped);
    a.cov3D_out = debug_dump_cov3D ? reinterpret_cast<float*>(ws + L.cov3d) : nullptr;
    a.block_sums = reinterpret_cast<uint32_t*>(ws + L.block_sums);
    a.depth_keys = reinterpret_cast<uint32_t*>(ws + L.dkeys);
    // the depth sort's state (control words, bucket counts, segment starts) is zeroed here because the preprocess
    // kernel accumulates the key range in its control words; num_rendered, the prefiltered-violation flag and the
    // tile-scan ticket live in spare control words of the same block (GsrDepthSortCtrl)
    a.depth_state = reinterpret_cast<uint32_t*>(ws + L.dstate);
    a.rects = reinterpret_cast<uint2*>(ws + L.rects);
    uint32_t* d_total = a.depth_state + GSR_DS_NUM_RENDERED;
    a.total = d_total;
    a.prefiltered = view->prefiltered ? 1 : 0;
    a.flags = a.depth_state + GSR_DS_PREFILTERED;
    GSR_CHECK(cudaMemsetAsync(ws + L.dstate, 0, L.dstate_bytes, stream));
    if (int rc = gsr_launch_preprocess_fwd(a, v, stream)) return rc;
    if (!sync) return 0;            // capacity mode: counters and flags are read back by gsr_forward_render_capacity
    if (a.prefiltered) {
        // rare path: one more 4-byte read-back before the sync
        static thread_local uint32_t h_flags;
        h_flags = 0;
        GSR_CHECK(cudaMemcpyAsync(&h_flags, a.flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        GSR_CHECK(cudaMemcpyAsync(host_num_rendered, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        GSR_CHECK(cudaStreamSynchronize(stream));
        if (h_flags & 1u)
            return gsr_set_error_msg(-4, "Point is filtered although prefiltered is set. This shouldn't happen!");
        return 0;
    }
    GSR_CHECK(cudaMemcpyAsync(host_num_rendered, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    GSR_CHECK(cudaStreamSynchronize(stream));
    return 0;
}

int gsr_forward_preprocess(const gsr_view* view, int P, int M, const float* means3D, const float* scales,
                           const float* rotations, const float* opacities, const float* shs,
                           const float* cov3D_precomp, const float* colors_precomp, const gsr_deform* deform,
                           float* means_out, int32_t* radii, void* geom_ws, size_t geom_bytes,
                           uint32_t* host_num_rendered, int debug_dump_cov3D, void* stream_) {
    return forward_preprocess_impl(view, P, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, deform,
                                   means_out, radii, geom_ws, geom_bytes, host_num_rendered, debug_dump_cov3D, stream_, true);
}
int gsr_forward_preprocess_async(const gsr_view* view, int P, int M, const float* means3D, const float* scales,
                                 const float* rotations, const float* opacities, const float* shs,
                                 const float* cov3D_precomp, const float* colors_precomp, const gsr_deform* deform,
                                 float* means_out, int32_t* radii, void* geom_ws, size_t geom_bytes, void* stream_) {
    return forward_preprocess_impl(view, P, M, means3D, scales, rotations, opacities, shs, cov3D_precomp, colors_precomp, deform,
                                   means_out, radii, geom_ws, geom_bytes, nullptr, 0, stream_, false);
}

// ---- view-batched forward preprocess (preprocess.cu: preprocess_fwd_batched_kernel) ----------------------------------------
size_t gsr_forward_batched_slots_bytes(int n_views) { return sizeof(FwdViewSlot) * (size_t)(n_views > 0 ? n_views : 0); }

// Pure host function: slots_host[j] = camera constants of view j + pointers into its geometry workspace.
int gsr_forward_batched_fill_slots(int n_views, const gsr_view_fwd* views, int P, int M, void* slots_host, size_t slots_bytes) {
    if (n_views <= 0 || !views || !slots_host) return gsr_set_error_msg(-1, "batched preprocess: no views");
    if (slots_bytes < gsr_forward_batched_slots_bytes(n_views)) return gsr_set_error_msg(-3, "batched preprocess: slot buffer too small");
    const GeomLayout L = geom_layout(P);
    FwdViewSlot* out = reinterpret_cast<FwdViewSlot*>(slots_host);
    for (int j = 0; j < n_views; j++) {
        const gsr_view_fwd& g = views[j];
        if (!g.view || !g.radii || !g.geom_ws) return gsr_set_error_msg(-1, "batched preprocess: NULL pointer in a view");
        if (g.view->prefiltered || g.view->debug) return gsr_set_error_msg(-1, "batched preprocess: prefiltered / debug views take the per-view path");
        FwdViewSlot sl{};
        if (int rc = fill_view(g.view, M, sl.v)) return rc;
        if (sl.v.scale_modifier != views[0].view->scale_modifier || sl.v.sh_degree != views[0].view->sh_degree)
            return gsr_set_error_msg(-1, "batched preprocess: the views of a batch must share scale_modifier and sh_degree");
        char* ws = reinterpret_cast<char*>(g.geom_ws);
        sl.radii = g.radii;
        sl.depths = reinterpret_cast<float*>(ws + L.depths);
        sl.tiles_touched = reinterpret_cast<uint32_t*>(ws + L.tiles);
        sl.recs = reinterpret_cast<float4*>(ws + L.recs);
        sl.clamped = reinterpret_cast<uint8_t*>(ws + L.clamped);
        sl.block_sums = reinterpret_cast<uint32_t*>(ws + L.block_sums);
        sl.depth_keys = reinterpret_cast<uint32_t*>(ws + L.dkeys);
        sl.depth_state = reinterpret_cast<uint32_t*>(ws + L.dstate);
        sl.rects = reinterpret_cast<uint2*>(ws + L.rects);
        out[j] = sl;
    }
    return 0;
}

int gsr_forward_preprocess_batched(int n_views, const gsr_view_fwd* views, const void* slots_device, int P, int M, const float* means3D,
                                   const float* scales, const float* rotations, const float* opacities, const float* shs,
                                   const gsr_deform* deform, float* means_out, size_t geom_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P <= 0 || n_views <= 0) return 0;
    if (!views || !slots_device || !means3D || !scales || !rotations || !opacities || !shs)
        return gsr_set_error_msg(-1, "batched preprocess: required pointer is NULL");
    if (M != 16 || (reinterpret_cast<uintptr_t>(shs) & 31))
        return gsr_set_error_msg(-1, "batched preprocess: needs M = 16 SH coefficients and 32-byte aligned shs");
    const GeomLayout L = geom_layout(P);
    if (geom_bytes < L.bytes) return gsr_set_error_msg(-3, "geometry workspace too small");
    PreprocessBatchArgs a{};
    a.P = P; a.means = means3D; a.scales = scales; a.rotations = rotations; a.opacities = opacities; a.shs = shs;
    a.deform_mode = deform ? deform->mode : GSR_DEFORM_NONE;
    if (a.deform_mode != GSR_DEFORM_NONE) {
        if (!deform->S || !deform->theta || !means_out) return gsr_set_error_msg(-1, "deform: S, theta and means_out required");
        if (a.deform_mode == GSR_DEFORM_RIGID_BODIES && !deform->body_id) return gsr_set_error_msg(-1, "deform: body_id required");
        a.twist_S = deform->S; a.twist_theta = deform->theta; a.body_id = deform->body_id;
    }
    a.means_out = means_out;
    a.scale_modifier = views[0].view->scale_modifier;
    for (int j = 0; j < n_views; j++)      // every view's depth-sort state (it also holds num_rendered and the flags)
        GSR_CHECK(cudaMemsetAsync(reinterpret_cast<char*>(views[j].geom_ws) + L.dstate, 0, L.dstate_bytes, stream));
    const FwdViewSlot* slots = reinterpret_cast<const FwdViewSlot*>(slots_device);
    for (int j0 = 0; j0 < n_views; j0 += GSR_BATCH_MAX_VIEWS) {
        a.n_views = n_views - j0 < GSR_BATCH_MAX_VIEWS ? n_views - j0 : GSR_BATCH_MAX_VIEWS;
        if (int rc = gsr_launch_preprocess_fwd_batched(a, slots + j0, stream)) return rc;
    }
    return 0;
}

// num_rendered of a finished preprocess (either variant): device -> host, asynchronous (the sync is the caller's)
int gsr_read_num_rendered(const void* geom_ws, int P, uint32_t* host_num_rendered, void* stream_) {
    if (!geom_ws || !host_num_rendered) return gsr_set_error_msg(-1, "read_num_rendered: NULL pointer");
    const GeomLayout L = geom_layout(P);
    const uint32_t* st = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(geom_ws) + L.dstate) + GSR_DS_NUM_RENDERED;
    GSR_CHECK(cudaMemcpyAsync(host_num_rendered, st, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    return 0;
}

static int forward_render_impl(const gsr_view* view, int P, uint32_t R, const int32_t* radii, void* geom_ws, void* binning_ws,
                               size_t binning_bytes, void* image_ws, float* out_color, int materialize_keys, void* stream_,
                               bool capacity_mode, uint32_t* host_status4) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GsrView v;
    if (int rc = fill_view(view, 0, v)) return rc;
    if (!image_ws || !out_color) return gsr_set_error_msg(-1, "forward_render: required pointer is NULL");
    const ImageLayout IL = image_layout(v.W, v.H);
    char* iw = reinterpret_cast<char*>(image_ws);
    uint2* ranges = reinterpret_cast<uint2*>(iw + IL.ranges);
    const int num_tiles = v.grid_x * v.grid_y;
    const uint32_t* point_list = nullptr;
    const float4* recs = nullptr;
    if (P > 0 && R > 0) {
        if (!geom_ws || !binning_ws || !radii) return gsr_set_error_msg(-1, "forward_render: workspace is NULL");
        if (R >= (1u << 30)) return gsr_set_error_msg(-2, "num_rendered must be < 2^30");
        const GeomLayout L = geom_layout(P);
        const BinLayout BL = bin_layout(R, v.W, v.H);
        if (binning_bytes < BL.bytes) return gsr_set_error_msg(-3, "binning workspace too small");
        char* gw = reinterpret_cast<char*>(geom_ws);
        char* bw = reinterpret_cast<char*>(binning_ws);
        recs = reinterpret_cast<const float4*>(gw + L.recs);
        const uint32_t* tiles_touched = reinterpret_cast<const uint32_t*>(gw + L.tiles);
        // 1. the emitting Gaussians in (depth bits, index) order, with their tile rects (srec)
        uint32_t* const dstate = reinterpret_cast<uint32_t*>(gw + L.dstate);
        uint32_t* const order = reinterpret_cast<uint32_t*>(gw + L.dorder);
        if (int rc = gsr_launch_depth_sort((uint32_t)P, reinterpret_cast<const uint32_t*>(gw + L.dkeys), dstate,
                                           reinterpret_cast<uint2*>(gw + L.dpairs_a), reinterpret_cast<uint2*>(gw + L.dpairs_b),
                                           reinterpret_cast<const uint2*>(gw + L.rects), order, reinterpret_cast<uint4*>(gw + L.srec),
                                           true, stream))
            return rc;
        const uint32_t* const n_emit = dstate + GSR_DS_N_EMIT;
        const uint32_t* sorted_tiles = nullptr;
        if (capacity_mode && !BL.sweep)
            return gsr_set_error_msg(-2, "forward_render_capacity: this image size takes the radix binning path, which needs the exact num_rendered");
        if (BL.sweep) {
            // 2+3. stable counting sort of the (never materialised) duplicates by tile id
            const GsrTileBinPlan pl = gsr_make_tile_bin_plan(v.grid_x, v.grid_y);
            uint32_t* plist = reinterpret_cast<uint32_t*>(bw + BL.vals_a);
            if (int rc = gsr_launch_tile_binning(P, n_emit, reinterpret_cast<const uint4*>(gw + L.srec), pl, v.grid_x, v.grid_y,
                                                 reinterpret_cast<uint32_t*>(bw + BL.matrix),
                                                 reinterpret_cast<uint32_t*>(bw + BL.totals),
                                                 reinterpret_cast<uint32_t*>(bw + BL.tile_base), ranges, plist,
                                                 dstate + GSR_DS_SCAN_TICKET, stream, capacity_mode ? R : 0u, dstate + GSR_DS_OVERFLOW))
                return rc;
            point_list = plist;
            if (materialize_keys) {
                uint32_t* tids = reinterpret_cast<uint32_t*>(bw + BL.tkeys_a);
                if (int rc = gsr_launch_expand_tile_ids(num_tiles, ranges, tids, stream)) return rc;
                sorted_tiles = tids;
            }
        } else {
        // 2. offsets in that order, then the duplicates (tile id, Gaussian id)
        uint32_t* sblock = reinterpret_cast<uint32_t*>(gw + L.sblock_sums);
        uint32_t* d_total = reinterpret_cast<uint32_t*>(gw + L.total) + 1;
        if (int rc = gsr_launch_sorted_block_sums(P, n_emit, order, tiles_touched, sblock, stream)) return rc;
        if (int rc = gsr_launch_scan_block_sums(sblock, gsr_div_up(P, 256), d_total, stream)) return rc;
        uint32_t* tkeys_a = reinterpret_cast<uint32_t*>(bw + BL.tkeys_a);
        uint32_t* tkeys_b = reinterpret_cast<uint32_t*>(bw + BL.tkeys_b);
        uint32_t* vals_a = reinterpret_cast<uint32_t*>(bw + BL.vals_a);
        uint32_t* vals_b = reinterpret_cast<uint32_t*>(bw + BL.vals_b);
        const GsrSortPlan tplan = gsr_make_sort_plan(0, BL.tile_bits);
        GSR_CHECK(cudaMemsetAsync(bw + BL.sort_temp, 0, BL.sort_temp_bytes, stream));
        if (int rc = gsr_launch_duplicate(P, n_emit, order, radii, tiles_touched, recs, sblock, tkeys_a, vals_a, v.grid_x,
                                          v.grid_y, tplan, reinterpret_cast<uint32_t*>(bw + BL.sort_temp), stream))
            return rc;
        // 3. stable sort by tile id only
        int in_b = 0;
        if (int rc = gsr_launch_sort_pairs32(tkeys_a, tkeys_b, vals_a, vals_b, R, 0, BL.tile_bits, bw + BL.sort_temp,
                                             BL.sort_temp_bytes, &in_b, stream, 2, true))
            return rc;
        sorted_tiles = in_b ? tkeys_b : tkeys_a;
        point_list = in_b ? vals_b : vals_a;
        if (int rc = gsr_launch_tile_ranges(R, sorted_tiles, ranges, num_tiles, stream)) return rc;
        }
        if (materialize_keys)
            if (int rc = gsr_launch_materialize_keys(R, sorted_tiles, point_list,
                                                     reinterpret_cast<const float*>(gw + L.depths),
                                                     reinterpret_cast<uint64_t*>(bw + BL.keys64), stream))
                return rc;
    } else {
        GSR_CHECK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
    }
    BlendFwdArgs b{};
    b.ranges = ranges; b.point_list = point_list; b.recs = recs;
    b.W = v.W; b.H = v.H; b.grid_x = v.grid_x; b.grid_y = v.grid_y;
    b.bg[0] = view->bg[0]; b.bg[1] = view->bg[1]; b.bg[2] = view->bg[2];
    b.out_color = out_color;
    if (point_list && gsr_blend_region_masks_enabled())
        b.region_masks = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(binning_ws) + bin_layout(R, v.W, v.H).masks);
    b.final_T = reinterpret_cast<float*>(iw + IL.final_T);
    b.n_contrib = reinterpret_cast<uint32_t*>(iw + IL.n_contrib);
    if (int rc = gsr_launch_blend_fwd(b, stream)) return rc;
    if (capacity_mode && host_status4 && P > 0 && geom_ws) {
        // control words 59..62 of the depth-sort state: overflow flag, prefiltered violation, (scan ticket), num_rendered
        const GeomLayout L = geom_layout(P);
        const uint32_t* st = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(geom_ws) + L.dstate) + GSR_DS_OVERFLOW;
        GSR_CHECK(cudaMemcpyAsync(host_status4, st, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    }
    return 0;
}

int gsr_forward_render(const gsr_view* view, int P, uint32_t R, const int32_t* radii, void* geom_ws, void* binning_ws,
                       size_t binning_bytes, void* image_ws, float* out_color, int materialize_keys, void* stream_) {
    return forward_render_impl(view, P, R, radii, geom_ws, binning_ws, binning_bytes, image_ws, out_color, materialize_keys, stream_,
                               false, nullptr);
}
int gsr_forward_render_capacity(const gsr_view* view, int P, uint32_t R_capacity, const int32_t* radii, void* geom_ws, void* binning_ws,
                                size_t binning_bytes, void* image_ws, float* out_color, uint32_t* host_status4, void* stream_) {
    if (R_capacity == 0) return gsr_set_error_msg(-1, "forward_render_capacity: capacity must be > 0");
    return forward_render_impl(view, P, R_capacity, radii, geom_ws, binning_ws, binning_bytes, image_ws, out_color, 0, stream_, true,
                               host_status4);
}

// Debug: workload counters of the blend stage for a finished forward (see blend.cu).
int gsr_ssim_l1_loss_forward(const float* image, const float* gt, int channels, int height, int width,
                             const float* window11, float lambda_dssim, float* dmaps, float* out8, void* stream_) {
    if (!image || !gt || !window11 || !dmaps || !out8) return gsr_set_error_msg(-1, "ssim_l1_loss: NULL pointer");
    if (channels != 3) return gsr_set_error_msg(-1, "ssim_l1_loss: images must be [3,H,W]");
    if (height <= 0 || width <= 0) return gsr_set_error_msg(-1, "ssim_l1_loss: empty image");
    return gsr_launch_ssim_l1_fwd(image, gt, height, width, window11, lambda_dssim, dmaps, out8, (cudaStream_t)stream_);
}

int gsr_ssim_l1_loss_backward(const float* image, const float* gt, int channels, int height, int width,
                              const float* window11, float lambda_dssim, const float* dmaps, const float* upstream,
                              float* dL_dimage, void* stream_) {
    if (!image || !gt || !window11 || !dmaps || !dL_dimage) return gsr_set_error_msg(-1, "ssim_l1_loss: NULL pointer");
    if (channels != 3) return gsr_set_error_msg(-1, "ssim_l1_loss: images must be [3,H,W]");
    if (height <= 0 || width <= 0) return gsr_set_error_msg(-1, "ssim_l1_loss: empty image");
    return gsr_launch_ssim_l1_bwd(image, gt, height, width, window11, lambda_dssim, dmaps, upstream, dL_dimage,
                                  (cudaStream_t)stream_);
}

int gsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int num_groups,
                  const unsigned long long* group_begin, const float* step_size, const float* bias_correction2_sqrt,
                  const double* beta1, const double* beta2, const float* eps, void* stream_) {
    if (num_groups < 0) return gsr_set_error_msg(-1, "adam: negative group count");
    if (num_groups == 0) return 0;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !group_begin || !step_size || !bias_correction2_sqrt || !beta1 ||
        !beta2 || !eps)
        return gsr_set_error_msg(-1, "adam: NULL pointer");
    return gsr_launch_adam(params, grads, exp_avg, exp_avg_sq, num_groups, group_begin, step_size, bias_correction2_sqrt,
                           beta1, beta2, eps, (cudaStream_t)stream_);
}

int gsr_debug_exp_check(float x_max, unsigned long long* out2, void* stream_) {
    if (!out2 || !(x_max > 0.0f)) return gsr_set_error_msg(-1, "exp_check: bad arguments");
    return gsr_launch_exp_check(x_max, out2, (cudaStream_t)stream_);
}

int gsr_debug_blend_stats(const gsr_view* view, int P, uint32_t R, const void* geom_ws, const void* binning_ws,
                          const void* image_ws, unsigned long long* out8, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GsrView v;
    if (int rc = fill_view(view, 0, v)) return rc;
    if (P <= 0 || R == 0 || !geom_ws || !binning_ws || !image_ws || !out8) return gsr_set_error_msg(-1, "blend_stats: bad arguments");
    const GeomLayout L = geom_layout(P);
    const ImageLayout IL = image_layout(v.W, v.H);
    const BinLayout BL = bin_layout(R, v.W, v.H);
    const bool in_b = !BL.sweep && (BL.passes & 1) != 0;
    BlendFwdArgs b{};
    b.ranges = reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(image_ws) + IL.ranges);
    b.point_list = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(binning_ws) + (in_b ? BL.vals_b : BL.vals_a));
    b.recs = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(geom_ws) + L.recs);
    b.W = v.W; b.H = v.H; b.grid_x = v.grid_x; b.grid_y = v.grid_y;
    GSR_CHECK(cudaMemsetAsync(out8, 0, 8 * sizeof(unsigned long long), stream));
    return gsr_launch_blend_stats(b, out8, stream);
}

// The depth sort on its own (tests, diagnostics): order[0 .. n) = indices of the keys != 0xffffffff by ascending (key, index).
size_t gsr_depth_order_ws_bytes(int P) {
    const size_t p = (size_t)(P > 0 ? P : 0);
    return al(gsr_depth_sort_state_bytes((uint32_t)p)) + 2 * al(8 * p) + 256;
}
int gsr_depth_order(const uint32_t* keys, int P, void* ws, size_t ws_bytes, uint32_t* order, uint32_t* info2, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0) return gsr_set_error_msg(-1, "depth_order: P must be >= 0");
    if (P == 0) return 0;
    if (!keys || !ws || !order) return gsr_set_error_msg(-1, "depth_order: NULL pointer");
    if (ws_bytes < gsr_depth_order_ws_bytes(P)) return gsr_set_error_msg(-3, "depth_order: workspace too small");
    char* w = reinterpret_cast<char*>(ws);
    const size_t sb = al(gsr_depth_sort_state_bytes((uint32_t)P));
    uint32_t* state = reinterpret_cast<uint32_t*>(w);
    GSR_CHECK(cudaMemsetAsync(state, 0, sb, stream));
    if (int rc = gsr_launch_depth_sort((uint32_t)P, keys, state, reinterpret_cast<uint2*>(w + sb),
                                       reinterpret_cast<uint2*>(w + sb + al(8 * (size_t)P)), nullptr, order, nullptr, false, stream))
        return rc;
    if (info2) {   // device u32[2]: entries of order[], segments that took the skew path
        GSR_CHECK(cudaMemcpyAsync(info2, state + GSR_DS_N_EMIT, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
        GSR_CHECK(cudaMemcpyAsync(info2 + 1, state + GSR_DS_SLOW_SEGMENTS, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    }
    return 0;
}

// Debug: iteration counts of the blend walk under finer lane-group culls (blend_v2.cu: blend_group_stats_kernel); out32 = 32 u64.
int gsr_debug_blend_group_stats(const gsr_view* view, int P, uint32_t R, const void* geom_ws, const void* binning_ws,
                                const void* image_ws, unsigned long long* out32, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GsrView v;
    if (int rc = fill_view(view, 0, v)) return rc;
    if (P <= 0 || R == 0 || !geom_ws || !binning_ws || !image_ws || !out32) return gsr_set_error_msg(-1, "blend_group_stats: bad arguments");
    const GeomLayout L = geom_layout(P);
    const ImageLayout IL = image_layout(v.W, v.H);
    const BinLayout BL = bin_layout(R, v.W, v.H);
    const bool in_b = !BL.sweep && (BL.passes & 1) != 0;
    BlendFwdArgs b{};
    b.ranges = reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(image_ws) + IL.ranges);
    b.point_list = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(binning_ws) + (in_b ? BL.vals_b : BL.vals_a));
    b.recs = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(geom_ws) + L.recs);
    b.W = v.W; b.H = v.H; b.grid_x = v.grid_x; b.grid_y = v.grid_y;
    GSR_CHECK(cudaMemsetAsync(out32, 0, 32 * sizeof(unsigned long long), stream));
    return gsr_launch_blend_group_stats(b, out32, stream);
}

// renderCUDA backward of one view: zeroes the per-Gaussian gradient records and lets blend_bwd reduce into them.
static int backward_blend_impl(const gsr_view* view, const GsrView& v, int P, uint32_t R, const void* geom_ws, const void* binning_ws,
                               const void* image_ws, float4* grad_recs, const float* dL_dout_color, cudaStream_t stream) {
    const GeomLayout L = geom_layout(P);
    const ImageLayout IL = image_layout(v.W, v.H);
    const char* gw = reinterpret_cast<const char*>(geom_ws);
    const char* iw = reinterpret_cast<const char*>(image_ws);
    GSR_CHECK(cudaMemsetAsync(grad_recs, 0, 48 * (size_t)P, stream));
    if (R > 0) {
        if (!binning_ws) return gsr_set_error_msg(-1, "backward: binning workspace is NULL");
        const BinLayout BL = bin_layout(R, v.W, v.H);
        const char* bw = reinterpret_cast<const char*>(binning_ws);
        const bool in_b = !BL.sweep && (BL.passes & 1) != 0;
        BlendBwdArgs b{};
        b.ranges = reinterpret_cast<const uint2*>(iw + IL.ranges);
        b.point_list = reinterpret_cast<const uint32_t*>(bw + (in_b ? BL.vals_b : BL.vals_a));
        b.recs = reinterpret_cast<const float4*>(gw + L.recs);
        b.W = v.W; b.H = v.H; b.grid_x = v.grid_x; b.grid_y = v.grid_y;
        b.bg[0] = view->bg[0]; b.bg[1] = view->bg[1]; b.bg[2] = view->bg[2];
        b.final_T = reinterpret_cast<const float*>(iw + IL.final_T);
        b.n_contrib = reinterpret_cast<const uint32_t*>(iw + IL.n_contrib);
        b.dL_dpix = dL_dout_color;
        b.grad_recs = grad_recs;
        if (gsr_blend_region_masks_enabled()) b.region_masks = reinterpret_cast<const uint32_t*>(bw + BL.masks);
        if (int rc = gsr_launch_blend_bwd(b, stream)) return rc;
    }
    return 0;
}

// ---- view-batched backward: gsr_backward_blend per view, then ONE gsr_backward_gaussians_batched over the Gaussians ----
int gsr_backward_blend(const gsr_view* view, int P, uint32_t R, const void* geom_ws, const void* binning_ws, const void* image_ws,
                       void* grad_ws, const float* dL_dout_color, void* stream_) {
    if (P <= 0) return 0;
    GsrView v;
    if (int rc = fill_view(view, 0, v)) return rc;
    if (!geom_ws || !image_ws || !grad_ws || !dL_dout_color) return gsr_set_error_msg(-1, "backward_blend: required pointer is NULL");
    return backward_blend_impl(view, v, P, R, geom_ws, binning_ws, image_ws, reinterpret_cast<float4*>(grad_ws), dL_dout_color,
                               (cudaStream_t)stream_);
}

size_t gsr_backward_batched_slots_bytes(int n_views) { return sizeof(BwdViewSlot) * (size_t)(n_views > 0 ? n_views : 0); }

// Pure host function: slots_host[j] = what the batched kernel needs of view j (camera constants + pointers into its workspaces).
int gsr_backward_batched_fill_slots(int n_views, const gsr_view_grads* views, int P, int M, void* slots_host, size_t slots_bytes) {
    if (n_views <= 0 || !views || !slots_host) return gsr_set_error_msg(-1, "batched backward: no views");
    if (slots_bytes < gsr_backward_batched_slots_bytes(n_views)) return gsr_set_error_msg(-3, "batched backward: slot buffer too small");
    const GeomLayout L = geom_layout(P);
    BwdViewSlot* out = reinterpret_cast<BwdViewSlot*>(slots_host);
    for (int j = 0; j < n_views; j++) {
        const gsr_view_grads& g = views[j];
        if (!g.view || !g.radii || !g.geom_ws || !g.grad_ws || !g.dL_dmeans2D) return gsr_set_error_msg(-1, "batched backward: NULL pointer in a view");
        BwdViewSlot sl{};
        if (int rc = fill_view(g.view, M, sl.v)) return rc;
        if (sl.v.scale_modifier != views[0].view->scale_modifier)
            return gsr_set_error_msg(-1, "batched backward: the views of a batch must share scale_modifier");
        const char* gw = reinterpret_cast<const char*>(g.geom_ws);
        sl.radii = g.radii;
        sl.clamped = reinterpret_cast<const uint8_t*>(gw + L.clamped);
        sl.grad_recs = reinterpret_cast<const float4*>(g.grad_ws);
        sl.recs = reinterpret_cast<const float4*>(gw + L.recs);
        sl.dL_dmeans2D = g.dL_dmeans2D;
        out[j] = sl;
    }
    return 0;
}

int gsr_backward_gaussians_batched(int n_views, const void* slots_device, float scale_modifier, int P, int first, int count, int M, const float* means3D,
                                   const float* means_deformed, const float* scales, const float* rotations, const float* shs,
                                   const gsr_deform* deform, float* dL_dmeans3D, float* dL_dopacity, float* dL_dsh, float* dL_dscales,
                                   float* dL_drots, float* dL_dtwist_S, float* dL_dtwist_theta, int accumulate_mask, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P <= 0 || n_views <= 0 || count == 0) return 0;
    if (first < 0 || count < 0 || (long long)first + count > P) return gsr_set_error_msg(-1, "batched backward: Gaussian range outside [0, P)");
    if (!slots_device || !means3D || !scales || !rotations || !shs || !dL_dmeans3D || !dL_dopacity || !dL_dsh || !dL_dscales || !dL_drots)
        return gsr_set_error_msg(-1, "batched backward: required pointer is NULL");
    if (M != 16 || ((reinterpret_cast<uintptr_t>(shs) | reinterpret_cast<uintptr_t>(dL_dsh)) & 31))
        return gsr_set_error_msg(-1, "batched backward: needs M = 16 SH coefficients and 32-byte aligned shs / dL_dsh");
    const int all_acc = GSR_ACC_MEANS3D | GSR_ACC_OPACITY | GSR_ACC_SH | GSR_ACC_SCALES | GSR_ACC_ROTS;
    if (n_views > GSR_BATCH_MAX_VIEWS && (accumulate_mask & all_acc) != all_acc)
        return gsr_set_error_msg(-1, "batched backward: more than 8 views need every output accumulated");
    PreprocessBwdBatchArgs a{};
    a.P = first + count; a.first = first; a.means = means3D;
    a.deform_mode = deform ? deform->mode : GSR_DEFORM_NONE;
    a.means_deformed = (a.deform_mode != GSR_DEFORM_NONE) ? means_deformed : nullptr;
    if (a.deform_mode != GSR_DEFORM_NONE) {
        if (!means_deformed || !deform->S || !deform->theta) return gsr_set_error_msg(-1, "batched backward: deform inputs missing");
        a.twist_S = deform->S; a.twist_theta = deform->theta; a.body_id = deform->body_id; a.num_bodies = deform->num_bodies;
        if (a.deform_mode == GSR_DEFORM_RIGID_BODIES && (!deform->body_id || deform->num_bodies <= 0))
            return gsr_set_error_msg(-1, "batched backward: body_id / num_bodies required");
        if ((dL_dtwist_S == nullptr) != (dL_dtwist_theta == nullptr))
            return gsr_set_error_msg(-1, "batched backward: give both twist gradients or neither");
        if (n_views > GSR_BATCH_MAX_VIEWS && dL_dtwist_S && a.deform_mode == GSR_DEFORM_PER_GAUSSIAN && !(accumulate_mask & GSR_ACC_TWIST))
            return gsr_set_error_msg(-1, "batched backward: more than 8 views need every output accumulated");
    }
    a.scales = scales; a.rotations = rotations; a.shs = shs;
    a.scale_modifier = scale_modifier;
    a.grad_moments = gsr_blend_bwd_writes_moments();
    a.dL_dmeans3D = dL_dmeans3D; a.dL_dopacity = dL_dopacity; a.dL_dsh = dL_dsh; a.dL_dscales = dL_dscales; a.dL_drots = dL_drots;
    a.dL_dtwist_S = dL_dtwist_S; a.dL_dtwist_theta = dL_dtwist_theta;
    a.acc = accumulate_mask;
    const BwdViewSlot* slots = reinterpret_cast<const BwdViewSlot*>(slots_device);
    for (int j0 = 0; j0 < n_views; j0 += GSR_BATCH_MAX_VIEWS) {
        a.n_views = n_views - j0 < GSR_BATCH_MAX_VIEWS ? n_views - j0 : GSR_BATCH_MAX_VIEWS;
        if (int rc = gsr_launch_preprocess_bwd_batched(a, slots + j0, stream)) return rc;
    }
    return 0;
}

int gsr_backward(const gsr_view* view, int P, int M, uint32_t R, const float* means3D, const float* means_deformed,
                 const float* scales, const float* rotations, const float* shs, const float* cov3D_precomp,
                 const float* colors_precomp, const gsr_deform* deform, const int32_t* radii, const void* geom_ws,
                 const void* binning_ws, const void* image_ws, void* grad_ws, const float* dL_dout_color,
                 float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dopacity, float* dL_dcolors, float* dL_dcov3D,
                 float* dL_dsh, float* dL_dscales, float* dL_drots, float* dL_dtwist_S, float* dL_dtwist_theta,
                 int accumulate_mask, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P <= 0) return 0;
    GsrView v;
    if (int rc = fill_view(view, M, v)) return rc;
    if (!means3D || !radii || !geom_ws || !image_ws || !grad_ws || !dL_dout_color || !dL_dmeans3D || !dL_dmeans2D ||
        !dL_dopacity)
        return gsr_set_error_msg(-1, "backward: required pointer is NULL");
    if ((colors_precomp && !dL_dcolors) || (cov3D_precomp && !dL_dcov3D))
        return gsr_set_error_msg(-1, "backward: gradient buffer for a precomputed input is NULL");
    if (shs && !dL_dsh) return gsr_set_error_msg(-1, "backward: dL_dsh required when shs given");
    const GeomLayout L = geom_layout(P);
    const char* gw = reinterpret_cast<const char*>(geom_ws);
    float4* grad_recs = reinterpret_cast<float4*>(grad_ws);
    if (int rc = backward_blend_impl(view, v, P, R, geom_ws, binning_ws, image_ws, grad_recs, dL_dout_color, stream)) return rc;
    PreprocessBwdArgs a{};
    a.P = P; a.means = means3D;
    a.deform_mode = deform ? deform->mode : GSR_DEFORM_NONE;
    a.means_deformed = (a.deform_mode != GSR_DEFORM_NONE) ? means_deformed : nullptr;
    if (a.deform_mode != GSR_DEFORM_NONE) {
        if (!means_deformed || !deform->S || !deform->theta) return gsr_set_error_msg(-1, "backward: deform inputs missing");
        a.twist_S = deform->S; a.twist_theta = deform->theta; a.body_id = deform->body_id; a.num_bodies = deform->num_bodies;
        if (a.deform_mode == GSR_DEFORM_RIGID_BODIES && (!deform->body_id || deform->num_bodies <= 0))
            return gsr_set_error_msg(-1, "backward: body_id / num_bodies required");
        if ((dL_dtwist_S == nullptr) != (dL_dtwist_theta == nullptr))
            return gsr_set_error_msg(-1, "backward: give both twist gradients or neither");
    }
    a.scales = scales; a.rotations = rotations; a.shs = shs; a.cov3D_precomp = cov3D_precomp; a.colors_precomp = colors_precomp;
    a.radii = radii; a.clamped = reinterpret_cast<const uint8_t*>(gw + L.clamped); a.grad_recs = grad_recs;
    a.grad_moments = (R > 0) ? gsr_blend_bwd_writes_moments() : 0;
    a.recs = reinterpret_cast<const float4*>(gw + L.recs); a.half_W = 0.5f * v.W; a.half_H = 0.5f * v.H;
    a.dL_dmeans3D = dL_dmeans3D; a.dL_dmeans2D = dL_dmeans2D; a.dL_dopacity = dL_dopacity; a.dL_dcolors = dL_dcolors;
    a.dL_dcov3D = dL_dcov3D; a.dL_dsh = dL_dsh; a.dL_dscales = scales ? dL_dscales : nullptr;
    a.dL_drots = scales ? dL_drots : nullptr;
    a.dL_dtwist_S = dL_dtwist_S; a.dL_dtwist_theta = dL_dtwist_theta;
    a.acc = accumulate_mask;
    return gsr_launch_preprocess_bwd(a, v, stream);
}

int gsr_mark_visible(const gsr_view* view, int P, const float* means3D, uint8_t* present, void* stream_) {
    if (P <= 0) return 0;
    GsrView v;
    gsr_view tmp = *view;
    if (tmp.image_width <= 0) tmp.image_width = 16;      // markVisible only needs the view matrix
    if (tmp.image_height <= 0) tmp.image_height = 16;
    if (int rc = fill_view(&tmp, 0, v)) return rc;
    if (!means3D || !present) return gsr_set_error_msg(-1, "mark_visible: NULL pointer");
    return gsr_launch_mark_visible(P, means3D, v, present, (cudaStream_t)stream_);
}

int gsr_mlp_gemm(const gsr_gemm* g, void* stream_) {
    if (!g) return gsr_set_error_msg(-1, "mlp_gemm: NULL argument block");
    GsrGemmArgs a{};
    a.M = g->M; a.N = g->N;
    a.A0_hi = g->A0_hi; a.A0_lo = g->A0_lo; a.K0 = g->K0; a.ldA0 = g->ldA0;
    a.A1_hi = g->A1_hi; a.A1_lo = g->A1_lo; a.K1 = g->K1; a.ldA1 = g->ldA1;
    a.B_hi = g->B_hi; a.B_lo = g->B_lo; a.ldB = g->ldB;
    a.mode = g->mode; a.k_splits = g->k_splits; a.bias = g->bias; a.mask_src = g->mask_src; a.ld_mask = g->ld_mask;
    a.out_hi = g->out_hi; a.out_lo = g->out_lo; a.ld_out = g->ld_out;
    a.outT_hi = g->outT_hi; a.outT_lo = g->outT_lo; a.ld_outT = g->ld_outT;
    a.colsum = g->colsum; a.error_flag = g->error_flag; a.mn_major = g->mn_major;
    a.prof_name = g->mode == GSR_GEMM_ATOMIC ? "mlp_gemm_dw" : (g->mask_src ? "mlp_gemm_dx" : "mlp_gemm_fwd");
    if (a.mode < 0 || a.mode > 3) return gsr_set_error_msg(-2, "mlp_gemm: unknown epilogue mode");
    return gsr_launch_mlp_gemm(a, (cudaStream_t)stream_);
}
int gsr_mlp_split(const float* x, int64_t n, float* hi, float* lo, void* stream_) {
    if (n > 0 && (!x || !hi || !lo)) return gsr_set_error_msg(-1, "mlp_split: NULL pointer");
    return gsr_launch_mlp_split(x, n, hi, lo, (cudaStream_t)stream_);
}
int gsr_mlp_split_transpose(const float* x, int rows, int cols, int ld_in, float* hi, float* lo, int ldT, void* stream_) {
    if (rows > 0 && cols > 0 && (!x || !hi || !lo)) return gsr_set_error_msg(-1, "mlp_split_transpose: NULL pointer");
    return gsr_launch_mlp_split_transpose(x, rows, cols, ld_in, hi, lo, ldT, (cudaStream_t)stream_);
}
int gsr_mlp_prepare(const float* x, int rows, int cols, int64_t ld_in, float* hi, float* lo, int64_t ld_out, float* hiT,
                    float* loT, int64_t ldT, float* colsum, void* stream_) {
    if (rows > 0 && cols > 0 && !x) return gsr_set_error_msg(-1, "mlp_prepare: NULL input");
    if ((lo && !hi) || (loT && !hiT)) return gsr_set_error_msg(-1, "mlp_prepare: a lo plane needs its hi plane");
    return gsr_launch_mlp_prepare(x, rows, cols, ld_in, hi, lo, ld_out, hiT, loT, ldT, colsum, (cudaStream_t)stream_);
}
int gsr_mlp_embed(const float* xyz, int P, float* e_hi, float* e_lo, float* eT_hi, float* eT_lo, int64_t ldT, void* stream_) {
    if (P > 0 && (!xyz || !e_hi)) return gsr_set_error_msg(-1, "mlp_embed: NULL pointer");
    if (eT_lo && !eT_hi) return gsr_set_error_msg(-1, "mlp_embed: a lo plane needs its hi plane");
    return gsr_launch_mlp_embed(xyz, P, e_hi, e_lo, eT_hi, eT_lo, ldT, (cudaStream_t)stream_);
}
int gsr_mlp_embed_backward(const float* xyz, int P, const float* d_embed, float* dxyz, int accumulate, void* stream_) {
    if (P > 0 && (!xyz || !d_embed || !dxyz)) return gsr_set_error_msg(-1, "mlp_embed_backward: NULL pointer");
    return gsr_launch_mlp_embed_bwd(xyz, P, d_embed, dxyz, accumulate, (cudaStream_t)stream_);
}

int gsr_densify_stats(int P, const float* viewspace_grad, const int32_t* radii, float* xyz_gradient_accum, float* xyz_gradient_accum_3vec,
                      float* denom, float* max_radii2D, void* stream_) {
    if (P > 0 && (!viewspace_grad || !radii || !xyz_gradient_accum || !denom || !max_radii2D)) return gsr_set_error_msg(-1, "densify_stats: NULL pointer");
    return gsr_launch_densify_stats(P, viewspace_grad, radii, xyz_gradient_accum, xyz_gradient_accum_3vec, denom, max_radii2D, (cudaStream_t)stream_);
}
int gsr_densify_decide(int P, const float* xyz_gradient_accum, const float* denom, const float* scaling_raw, float grad_threshold,
                       float size_threshold, uint8_t* flags, void* stream_) {
    if (P > 0 && (!xyz_gradient_accum || !denom || !scaling_raw || !flags)) return gsr_set_error_msg(-1, "densify_decide: NULL pointer");
    return gsr_launch_densify_decide(P, xyz_gradient_accum, denom, scaling_raw, grad_threshold, size_threshold, flags, (cudaStream_t)stream_);
}
int gsr_densify_split(int n, int N, const float* xyz, const float* scaling_raw, const float* rotation_raw, const float* normals, float* new_xyz,
                      float* new_scaling, void* stream_) {
    if (n > 0 && (!xyz || !scaling_raw || !rotation_raw || !normals || !new_xyz || !new_scaling)) return gsr_set_error_msg(-1, "densify_split: NULL pointer");
    return gsr_launch_densify_split(n, N, xyz, scaling_raw, rotation_raw, normals, new_xyz, new_scaling, (cudaStream_t)stream_);
}
int gsr_densify_prune(int P, const float* opacity_raw, const float* scaling_raw, const float* max_radii2D, float min_opacity,
                      float max_screen_size, float world_size_limit, int use_size, uint8_t* prune, void* stream_) {
    if (P > 0 && (!opacity_raw || !scaling_raw || !prune || (use_size && !max_radii2D))) return gsr_set_error_msg(-1, "densify_prune: NULL pointer");
    return gsr_launch_densify_prune(P, opacity_raw, scaling_raw, max_radii2D, min_opacity, max_screen_size, world_size_limit, use_size, prune,
                                    (cudaStream_t)stream_);
}

int gsr_deform_glue_forward(int P, const float* heads, const float* xyz, const float* scaling, const float* rotation, const float* f_dc,
                            const float* f_rest, float* means3D, float* scales, float* rotations, float* shs, void* stream_) {
    if (P > 0 && (!heads || !xyz || !scaling || !rotation || !f_dc || !f_rest || !means3D || !scales || !rotations || !shs))
        return gsr_set_error_msg(-1, "deform_glue_forward: NULL pointer");
    if (reinterpret_cast<uintptr_t>(heads) & 15) return gsr_set_error_msg(-2, "deform_glue_forward: heads must be 16-byte aligned");
    return gsr_launch_deform_glue_fwd(P, heads, xyz, scaling, rotation, f_dc, f_rest, means3D, scales, rotations, shs, (cudaStream_t)stream_);
}
int gsr_deform_glue_backward(int P, const float* heads, const float* rotation, const float* scales, const float* g_means, const float* g_scales,
                             const float* g_rotations, const float* g_shs, float* d_heads, float* d_xyz, float* d_scaling, float* d_rotation,
                             float* d_f_dc, float* d_f_rest, void* stream_) {
    if (P > 0 && (!heads || !rotation || !scales || !d_heads)) return gsr_set_error_msg(-1, "deform_glue_backward: NULL pointer");
    if ((reinterpret_cast<uintptr_t>(heads) | reinterpret_cast<uintptr_t>(d_heads)) & 15)
        return gsr_set_error_msg(-2, "deform_glue_backward: heads / d_heads must be 16-byte aligned");
    return gsr_launch_deform_glue_bwd(P, heads, rotation, scales, g_means, g_scales, g_rotations, g_shs, d_heads, d_xyz, d_scaling, d_rotation,
                                      d_f_dc, d_f_rest, (cudaStream_t)stream_);
}

size_t gsr_knn_bytes(int P) { return gsr_knn_temp_bytes(P); }
int gsr_knn_dist2(int P, const float* points, float* mean_dist2, void* temp, size_t temp_bytes, void* stream_) {
    if (P > 0 && (!points || !mean_dist2 || !temp)) return gsr_set_error_msg(-1, "knn: NULL pointer");
    return gsr_launch_knn_dist2(P, points, mean_dist2, temp, temp_bytes, (cudaStream_t)stream_);
}

int gsr_exp_se3(int N, const float* S, const float* theta, float* T44, void* stream_) {
    if (N > 0 && (!S || !theta || !T44)) return gsr_set_error_msg(-1, "exp_se3: NULL pointer");
    return gsr_launch_se3_matrices(N, S, theta, T44, (cudaStream_t)stream_);
}
int gsr_exp_se3_backward(int N, const float* S, const float* theta, const float* dT44, float* dS, float* dtheta,
                         void* stream_) {
    if (N > 0 && (!S || !theta || !dT44 || !dS || !dtheta)) return gsr_set_error_msg(-1, "exp_se3_backward: NULL pointer");
    return gsr_launch_se3_matrices_bwd(N, S, theta, dT44, dS, dtheta, (cudaStream_t)stream_);
}

size_t gsr_sort_bytes(uint32_t n, int begin_bit, int end_bit) { return gsr_sort_temp_bytes(n, begin_bit, end_bit); }
int gsr_sort_pairs(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t n, int begin_bit,
                   int end_bit, void* temp, size_t temp_bytes, int* result_in_b, void* stream_) {
    int dummy = 0;
    return gsr_launch_sort_pairs(keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes,
                                 result_in_b ? result_in_b : &dummy, (cudaStream_t)stream_);
}

int gsr_sort_pairs32(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t n, int begin_bit,
                     int end_bit, void* temp, size_t temp_bytes, int* result_in_b, void* stream_) {
    int dummy = 0;
    return gsr_launch_sort_pairs32(keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes,
                                   result_in_b ? result_in_b : &dummy, (cudaStream_t)stream_);
}

}  // extern "C"
