// Fused SE3-deform + preprocess forward, and the frustum test (markVisible).
//
// Replaces, in one pass over the Gaussians:
//   scene/rigid_body.py:86-93  exp_se3  +  gaussian_renderer/__init__.py:92-95 apply
//   cuda_rasterizer/forward.cu:155-256  preprocessCUDA<3>
//   the per-block partial sums of cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:277)
//
// HBM-bound kernel: 236 B in (+28 B twist) and ~84 B out per Gaussian.  One thread
// per Gaussian; the 192-byte SH record and the 48-byte splat record are moved as
// 128-bit vectors; SH is only touched for Gaussians that survive culling.
#include "geom_exact.cuh"
#include "kernels.cuh"

// Spherical-harmonics constants (same values as auxiliary.h:22-38 / utils/sh_utils.py).
__device__ __constant__ float kSH_C0 = 0.28209479177387814f;
__device__ __constant__ float kSH_C1 = 0.4886025119029199f;
__device__ __constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator*(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

// forward.cu:20-71.  `sh` points at this Gaussian's M x 3 coefficients already in
// registers/local (48 floats for M = 16).  Returns the un-clamped colour + 0.5.
template <typename ShFetch>
__device__ __forceinline__ V3 sh_to_rgb(int deg, float3 pos, const float* campos, ShFetch sh) {
    V3 dir = {pos.x - campos[0], pos.y - campos[1], pos.z - campos[2]};
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x = dir.x / len; dir.y = dir.y / len; dir.z = dir.z / len;
    V3 result = kSH_C0 * sh(0);
    if (deg > 0) {
        const float x = dir.x, y = dir.y, z = dir.z;
        result = result - kSH_C1 * y * sh(1) + kSH_C1 * z * sh(2) - kSH_C1 * x * sh(3);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z;
            const float xy = x * y, yz = y * z, xz = x * z;
            result = result + kSH_C2[0] * xy * sh(4) + kSH_C2[1] * yz * sh(5) +
                     kSH_C2[2] * (2.0f * zz - xx - yy) * sh(6) + kSH_C2[3] * xz * sh(7) +
                     kSH_C2[4] * (xx - yy) * sh(8);
            if (deg > 2) {
                result = result + kSH_C3[0] * y * (3.0f * xx - yy) * sh(9) + kSH_C3[1] * xy * z * sh(10) +
                         kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh(11) +
                         kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh(12) +
                         kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh(13) + kSH_C3[5] * z * (xx - yy) * sh(14) +
                         kSH_C3[6] * x * (xx - 3.0f * yy) * sh(15);
            }
        }
    }
    result.x += 0.5f; result.y += 0.5f; result.z += 0.5f;
    return result;
}


__global__ void __launch_bounds__(256, 3) preprocess_fwd_kernel(PreprocessArgs a, GsrView v) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    uint32_t tiles = 0;
    uint32_t dkey = 0xffffffffu;
    if (idx < a.P) {
        // ---- (1) every unconditional load first: no store may sit between them (a store to a
        // non-restrict pointer pins all later loads behind it and serialises DRAM round trips) ----
        float3 p = make_float3(a.means[3 * idx], a.means[3 * idx + 1], a.means[3 * idx + 2]);
        float3 w = make_float3(0, 0, 0), tv = w;
        float th = 0.0f;
        if (a.deform_mode != GSR_DEFORM_NONE) {
            const int t = (a.deform_mode == GSR_DEFORM_RIGID_BODIES) ? a.body_id[idx] : idx;
            const float* S = a.twist_S + 6 * (size_t)t;
            w = make_float3(S[0], S[1], S[2]); tv = make_float3(S[3], S[4], S[5]);
            th = a.twist_theta[t];
        }
        float cov6[6];
        float3 s = make_float3(0, 0, 0);
        float4 q = make_float4(0, 0, 0, 0);
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov6[k] = a.cov3D_precomp[6 * (size_t)idx + k];
        } else {
            s = make_float3(a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2]);
            q = ld_rec4(a.rotations, idx);
        }
        const float op = a.opacities[idx];
        // ---- (2) geometry ----
        if (a.deform_mode != GSR_DEFORM_NONE) p = se3_apply(p, w, tv, th);
        int radius = 0;
        const float depth = xform_row(v.view, 2, p);
        SplatGeom g;
        g.ok = false;
        // auxiliary.h:154-160: with `prefiltered` the caller promises that no point fails this test; the reference
        // printf()s and __trap()s (killing the context).  Here the violation raises a flag that the host turns into
        // the same message as an error return, and the context survives.
        if (a.prefiltered && !(depth > GSR_NEAR)) atomicOr(a.flags, 1u);
        if (depth > GSR_NEAR) {
            if (!a.cov3D_precomp) cov3d_exact(s, v.scale_modifier, q, cov6);
            g = splat_geometry_exact(p, cov6, v);
        }
        // ---- (3) second round trip, survivors only: the 192-byte SH record ----
        V3 rgb = {0, 0, 0};
        uint8_t cl = 0;
        if (g.ok) {
            if (a.colors_precomp) {
                rgb = {a.colors_precomp[3 * idx], a.colors_precomp[3 * idx + 1], a.colors_precomp[3 * idx + 2]};
            } else {
                // 6 x LDG.256 (M == 16, 32-byte aligned), scalar otherwise.
                float shv[48];
                const int M = v.sh_coeffs;
                const int need = (v.sh_degree + 1) * (v.sh_degree + 1) * 3;
                const float* base = a.shs + (size_t)idx * M * 3;
                if (M == 16 && ((reinterpret_cast<uintptr_t>(a.shs) & 31) == 0)) {
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        if (8 * k < need) ld256_nc(base + 8 * k, shv + 8 * k);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 48; k++) shv[k] = (k < need) ? base[k] : 0.0f;
                }
                auto fetch = [&](int k) -> V3 { return {shv[3 * k], shv[3 * k + 1], shv[3 * k + 2]}; };
                rgb = sh_to_rgb(v.sh_degree, p, v.campos, fetch);
                cl = (rgb.x < 0 ? 1 : 0) | (rgb.y < 0 ? 2 : 0) | (rgb.z < 0 ? 4 : 0);
                rgb.x = fmaxf(rgb.x, 0.0f); rgb.y = fmaxf(rgb.y, 0.0f); rgb.z = fmaxf(rgb.z, 0.0f);
            }
        }
        // ---- (4) stores ----
        if (a.deform_mode != GSR_DEFORM_NONE) {
            a.means_out[3 * idx] = p.x; a.means_out[3 * idx + 1] = p.y; a.means_out[3 * idx + 2] = p.z;
        }
        if (a.cov3D_out && depth > GSR_NEAR && !a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) a.cov3D_out[6 * (size_t)idx + k] = cov6[k];
        }
        if (g.ok) {
            // alpha = min(.99, op*exp(power)) >= 1/255 needs power >= log(1/(255 op)).
            // Margin 1e-3 dwarfs the rounding of expf and of this log.
            float cut = (op > 0.0f) ? fmaxf(__logf(1.0f / (255.0f * op)) - 1e-3f, -80.0f) : 1.0f;   // >= -80: see f32x2.cuh exp2_exact
            radius = g.radius;
            tiles = (g.rmax.y - g.rmin.y) * (g.rmax.x - g.rmin.x);
            a.depths[idx] = g.depth;
            float4* r = a.recs + 3 * (size_t)idx;
            r[0] = make_float4(g.pix.x, g.pix.y, g.conic.x, g.conic.y);
            r[1] = make_float4(g.conic.z, op, rgb.x, rgb.y);
            r[2] = make_float4(rgb.z, cut, 0.0f, 0.0f);
            a.clamped[idx] = cl;
        }
        a.radii[idx] = radius;
        a.tiles_touched[idx] = tiles;
        a.rects[idx] = g.ok ? make_uint2(g.rmin.x | (g.rmin.y << 16), g.rmax.x | (g.rmax.y << 16)) : make_uint2(0u, 0u);
        dkey = tiles ? __float_as_uint(depth) : 0xffffffffu;
        a.depth_keys[idx] = dkey;
    }
    // Per-block: sum of tiles_touched (num_rendered; the radix path's offsets scan) and the range of the emitting
    // depth keys, which the depth sort (depth_sort.cu) slices into buckets - one atomic each per block.
    __shared__ uint32_t wsum[8], wmin[8], wmax[8];
    const uint32_t s = __reduce_add_sync(0xffffffffu, tiles);
    const uint32_t kmn = __reduce_min_sync(0xffffffffu, dkey);                       // culled: 0xffffffff
    const uint32_t kmx = __reduce_max_sync(0xffffffffu, tiles ? dkey : 0u);
    if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = s; wmin[threadIdx.x >> 5] = kmn; wmax[threadIdx.x >> 5] = kmx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0, mn = 0xffffffffu, mx = 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) { t += wsum[k]; mn = min(mn, wmin[k]); mx = max(mx, wmax[k]); }
        a.block_sums[blockIdx.x] = t;
        if (t) {
            atomicAdd(a.total, t);              // num_rendered without a scan launch in front of the host read-back
            atomicMax(a.depth_state + GSR_DS_NOT_KMIN, ~mn);
            atomicMax(a.depth_state + GSR_DS_KMAX, mx);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// View-batched forward preprocess: the views of a training step share the parameters, so one thread takes a Gaussian
// through ALL of them - mean / twist / scale / rotation / opacity read once, SE3 and cov3D evaluated once, the 192-byte SH
// record read once (when any view sees the Gaussian) - and writes each view's records into that view's workspace.
// Same device functions, hence the same bits, as preprocess_fwd_kernel (tests compare the workspaces byte for byte).
// Supported: scales + rotations, SH colours; any deform mode.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) preprocess_fwd_batched_kernel(PreprocessBatchArgs a, const FwdViewSlot* __restrict__ g_slots) {
    __shared__ __align__(16) FwdViewSlot slots[GSR_BATCH_MAX_VIEWS];
    __shared__ uint32_t s_red[GSR_BATCH_MAX_VIEWS][8][3];      // per view and warp: sum of tiles, min key, max key
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g_slots);
        uint32_t* dst = reinterpret_cast<uint32_t*>(slots);
        const int words = a.n_views * (int)(sizeof(FwdViewSlot) / 4);
        for (int k = threadIdx.x; k < words; k += 256) dst[k] = __ldg(src + k);
    }
    __syncthreads();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const bool in = idx < a.P;
    float3 p = make_float3(0, 0, 0);
    float3 s = make_float3(0, 0, 0);
    float4 q = make_float4(0, 0, 0, 0);
    float op = 0.0f;
    if (in) {
        p = make_float3(a.means[3 * idx], a.means[3 * idx + 1], a.means[3 * idx + 2]);
        float3 w = make_float3(0, 0, 0), tv = w;
        float th = 0.0f;
        if (a.deform_mode != GSR_DEFORM_NONE) {
            const int t = (a.deform_mode == GSR_DEFORM_RIGID_BODIES) ? a.body_id[idx] : idx;
            const float* S = a.twist_S + 6 * (size_t)t;
            w = make_float3(S[0], S[1], S[2]); tv = make_float3(S[3], S[4], S[5]);
            th = a.twist_theta[t];
        }
        s = make_float3(a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2]);
        q = ld_rec4(a.rotations, idx);
        op = a.opacities[idx];
        if (a.deform_mode != GSR_DEFORM_NONE) {
            p = se3_apply(p, w, tv, th);
            a.means_out[3 * idx] = p.x; a.means_out[3 * idx + 1] = p.y; a.means_out[3 * idx + 2] = p.z;
        }
    }
    float cov6[6];
    bool cov_done = false, sh_done = false;
    float shv[48];
    const float cut = (op > 0.0f) ? fmaxf(__logf(1.0f / (255.0f * op)) - 1e-3f, -80.0f) : 1.0f;
    for (int j = 0; j < a.n_views; j++) {
        const FwdViewSlot* sl = slots + j;
        const GsrView& v = sl->v;
        uint32_t tiles = 0, dkey = 0xffffffffu;
        if (in) {
            int radius = 0;
            const float depth = xform_row(v.view, 2, p);
            SplatGeom g;
            g.ok = false;
            if (depth > GSR_NEAR) {
                if (!cov_done) { cov3d_exact(s, a.scale_modifier, q, cov6); cov_done = true; }
                g = splat_geometry_exact(p, cov6, v);
            }
            if (g.ok) {
                if (!sh_done) {
                    const int need = (v.sh_degree + 1) * (v.sh_degree + 1) * 3;
                    const float* base = a.shs + (size_t)idx * 48;
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        if (8 * k < need) ld256_nc(base + 8 * k, shv + 8 * k);
                    }
                    sh_done = true;
                }
                auto fetch = [&](int k) -> V3 { return {shv[3 * k], shv[3 * k + 1], shv[3 * k + 2]}; };
                V3 rgb = sh_to_rgb(v.sh_degree, p, v.campos, fetch);
                const uint8_t cl = (rgb.x < 0 ? 1 : 0) | (rgb.y < 0 ? 2 : 0) | (rgb.z < 0 ? 4 : 0);
                rgb.x = fmaxf(rgb.x, 0.0f); rgb.y = fmaxf(rgb.y, 0.0f); rgb.z = fmaxf(rgb.z, 0.0f);
                radius = g.radius;
                tiles = (g.rmax.y - g.rmin.y) * (g.rmax.x - g.rmin.x);
                sl->depths[idx] = g.depth;
                float4* r = sl->recs + 3 * (size_t)idx;
                r[0] = make_float4(g.pix.x, g.pix.y, g.conic.x, g.conic.y);
                r[1] = make_float4(g.conic.z, op, rgb.x, rgb.y);
                r[2] = make_float4(rgb.z, cut, 0.0f, 0.0f);
                sl->clamped[idx] = cl;
            }
            sl->radii[idx] = radius;
            sl->tiles_touched[idx] = tiles;
            sl->rects[idx] = g.ok ? make_uint2(g.rmin.x | (g.rmin.y << 16), g.rmax.x | (g.rmax.y << 16)) : make_uint2(0u, 0u);
            dkey = tiles ? __float_as_uint(depth) : 0xffffffffu;
            sl->depth_keys[idx] = dkey;
        }
        const uint32_t sum = __reduce_add_sync(0xffffffffu, tiles);
        const uint32_t kmn = __reduce_min_sync(0xffffffffu, dkey);
        const uint32_t kmx = __reduce_max_sync(0xffffffffu, tiles ? dkey : 0u);
        if ((threadIdx.x & 31) == 0) { s_red[j][threadIdx.x >> 5][0] = sum; s_red[j][threadIdx.x >> 5][1] = kmn; s_red[j][threadIdx.x >> 5][2] = kmx; }
    }
    __syncthreads();
    if (threadIdx.x < a.n_views) {
        const int j = threadIdx.x;
        uint32_t t = 0, mn = 0xffffffffu, mx = 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) { t += s_red[j][k][0]; mn = min(mn, s_red[j][k][1]); mx = max(mx, s_red[j][k][2]); }
        slots[j].block_sums[blockIdx.x] = t;
        if (t) {
            atomicAdd(slots[j].depth_state + GSR_DS_NUM_RENDERED, t);
            atomicMax(slots[j].depth_state + GSR_DS_NOT_KMIN, ~mn);
            atomicMax(slots[j].depth_state + GSR_DS_KMAX, mx);
        }
    }
}

// rasterizer_impl.cu:54-66 checkFrustum: present = (view-space z > 0.2).
__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* means, GsrView v, uint8_t* present) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= P) return;
    const float3 p = make_float3(means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]);
    present[idx] = xform_row(v.view, 2, p) > GSR_NEAR ? 1 : 0;
}

int gsr_launch_preprocess_fwd(const PreprocessArgs& a, const GsrView& v, cudaStream_t stream) {
    if (a.P <= 0) return 0;
    { GsrProfScope prof_("preprocess_fwd", stream);
    preprocess_fwd_kernel<<<gsr_div_up(a.P, 256), 256, 0, stream>>>(a, v); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_preprocess_fwd_batched(const PreprocessBatchArgs& a, const FwdViewSlot* d_slots, cudaStream_t stream) {
    if (a.P <= 0 || a.n_views <= 0) return 0;
    { GsrProfScope prof_("preprocess_fwd_batched", stream);
    preprocess_fwd_batched_kernel<<<gsr_div_up(a.P, 256), 256, 0, stream>>>(a, d_slots); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_mark_visible(int P, const float* means, const GsrView& v, uint8_t* present, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("mark_visible", stream);
    mark_visible_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, means, v, present); }
    GSR_CHECK_LAUNCH();
    return 0;
}
