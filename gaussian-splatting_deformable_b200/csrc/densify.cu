// Densification of the Gaussian set: the per-point statistics and the clone / split / prune decisions
// of scene/gaussian_model.py:1129-1257 and train.py:610-648 as single passes over the flat per-point arrays.
// The optimizer-state surgery that follows the decisions is fused_adam.FusedAdam.prune / append / replace.
#include "kernels.cuh"

namespace {

// train.py:613,618 + gaussian_model.py:1252-1257: for the points the view saw (radii > 0)
//   max_radii2D = max(max_radii2D, radii);  accum3 += grad;  accum += |grad.xy|;  denom += 1
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const float* __restrict__ grad2d /*[P,3]*/, const int* __restrict__ radii,
                                                           float* __restrict__ accum, float* __restrict__ accum3,
                                                           float* __restrict__ denom, float* __restrict__ max_radii) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1], gz = grad2d[3 * (size_t)i + 2];
    max_radii[i] = fmaxf(max_radii[i], (float)r);
    if (accum3) { accum3[3 * (size_t)i] += gx; accum3[3 * (size_t)i + 1] += gy; accum3[3 * (size_t)i + 2] += gz; }
    accum[i] += sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));       // torch.norm(grad[:, :2], dim=-1)
    denom[i] += 1.0f;
}

// gaussian_model.py:1220-1221 + 1187-1190 + 1132-1137: grads = accum / denom (NaN -> 0), and per point
//   bit 0 (clone): grads >= thr  and  max(exp(scaling)) <= percent_dense * extent
//   bit 1 (split): grads >= thr  and  max(exp(scaling)) >  percent_dense * extent
// (the reference evaluates the split test after appending the clones, on padded gradients that are 0 for the new
// rows: a clone is never split in the same call, so both tests are decided here, on the original rows).
__global__ void __launch_bounds__(256) densify_decide_kernel(int P, const float* __restrict__ accum, const float* __restrict__ denom,
                                                            const float* __restrict__ scaling_raw /*[P,3] log-scales*/,
                                                            float grad_threshold, float size_threshold, uint8_t* __restrict__ flags) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    float g = accum[i] / denom[i];
    if (g != g) g = 0.0f;
    const float s = fmaxf(fmaxf(expf(scaling_raw[3 * (size_t)i]), expf(scaling_raw[3 * (size_t)i + 1])), expf(scaling_raw[3 * (size_t)i + 2]));
    const bool sel = fabsf(g) >= grad_threshold;                 // torch.norm over the size-1 last dimension = |g|
    uint8_t f = 0;
    if (sel && s <= size_threshold) f |= 1;
    if (g >= grad_threshold && s > size_threshold) f |= 2;
    flags[i] = f;
}

// gaussian_model.py:1139-1145 for the n selected points, N samples each (row r of the output = sample r / n of point r % n,
// the order torch's .repeat(N, 1) produces):  xyz' = R(q) (z * exp(scaling)) + xyz,  scaling' = log(exp(scaling) / (0.8 N))
// with R = build_rotation(q) (utils/general_utils.py:78-99, q normalised inside).
__global__ void __launch_bounds__(256) densify_split_kernel(int n, int N, const float* __restrict__ xyz, const float* __restrict__ scaling_raw,
                                                           const float* __restrict__ rot_raw, const float* __restrict__ z /*[n N,3] standard normals*/,
                                                           float* __restrict__ new_xyz, float* __restrict__ new_scaling) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n * N) return;
    const int i = r % n;
    const float sx = expf(scaling_raw[3 * (size_t)i]), sy = expf(scaling_raw[3 * (size_t)i + 1]), sz = expf(scaling_raw[3 * (size_t)i + 2]);
    const float vx = z[3 * (size_t)r] * sx, vy = z[3 * (size_t)r + 1] * sy, vz = z[3 * (size_t)r + 2] * sz;
    float qr = rot_raw[4 * (size_t)i], qx = rot_raw[4 * (size_t)i + 1], qy = rot_raw[4 * (size_t)i + 2], qz = rot_raw[4 * (size_t)i + 3];
    const float nrm = sqrtf(qr * qr + qx * qx + qy * qy + qz * qz);
    qr /= nrm; qx /= nrm; qy /= nrm; qz /= nrm;
    const float R00 = 1 - 2 * (qy * qy + qz * qz), R01 = 2 * (qx * qy - qr * qz), R02 = 2 * (qx * qz + qr * qy);
    const float R10 = 2 * (qx * qy + qr * qz), R11 = 1 - 2 * (qx * qx + qz * qz), R12 = 2 * (qy * qz - qr * qx);
    const float R20 = 2 * (qx * qz - qr * qy), R21 = 2 * (qy * qz + qr * qx), R22 = 1 - 2 * (qx * qx + qy * qy);
    new_xyz[3 * (size_t)r] = R00 * vx + R01 * vy + R02 * vz + xyz[3 * (size_t)i];
    new_xyz[3 * (size_t)r + 1] = R10 * vx + R11 * vy + R12 * vz + xyz[3 * (size_t)i + 1];
    new_xyz[3 * (size_t)r + 2] = R20 * vx + R21 * vy + R22 * vz + xyz[3 * (size_t)i + 2];
    const float inv = 1.0f / (0.8f * (float)N);
    new_scaling[3 * (size_t)r] = logf(sx * inv); new_scaling[3 * (size_t)r + 1] = logf(sy * inv); new_scaling[3 * (size_t)r + 2] = logf(sz * inv);
}

// gaussian_model.py:1226-1231: prune = sigmoid(opacity) < min_opacity  or  max_radii2D > max_screen_size  or
// max(exp(scaling)) > 0.1 extent  (the last two only when max_screen_size is given); `also` ORs in the split originals.
__global__ void __launch_bounds__(256) densify_prune_kernel(int P, const float* __restrict__ opacity_raw, const float* __restrict__ scaling_raw,
                                                           const float* __restrict__ max_radii, float min_opacity, float max_screen_size,
                                                           float world_size_limit, int use_size, uint8_t* __restrict__ prune) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float o = 1.0f / (1.0f + expf(-opacity_raw[i]));
    bool p = o < min_opacity;
    if (use_size) {
        const float s = fmaxf(fmaxf(expf(scaling_raw[3 * (size_t)i]), expf(scaling_raw[3 * (size_t)i + 1])), expf(scaling_raw[3 * (size_t)i + 2]));
        p = p || max_radii[i] > max_screen_size || s > world_size_limit;
    }
    prune[i] = p ? 1 : 0;
}

}  // namespace

int gsr_launch_densify_stats(int P, const float* grad2d, const int* radii, float* accum, float* accum3, float* denom, float* max_radii,
                             cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("densify_stats", stream);
    densify_stats_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, grad2d, radii, accum, accum3, denom, max_radii); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_densify_decide(int P, const float* accum, const float* denom, const float* scaling_raw, float grad_threshold,
                              float size_threshold, uint8_t* flags, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("densify_decide", stream);
    densify_decide_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, accum, denom, scaling_raw, grad_threshold, size_threshold, flags); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_densify_split(int n, int N, const float* xyz, const float* scaling_raw, const float* rot_raw, const float* z, float* new_xyz,
                             float* new_scaling, cudaStream_t stream) {
    if (n <= 0 || N <= 0) return 0;
    { GsrProfScope prof_("densify_split", stream);
    densify_split_kernel<<<gsr_div_up((long long)n * N, 256), 256, 0, stream>>>(n, N, xyz, scaling_raw, rot_raw, z, new_xyz, new_scaling); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_densify_prune(int P, const float* opacity_raw, const float* scaling_raw, const float* max_radii, float min_opacity,
                             float max_screen_size, float world_size_limit, int use_size, uint8_t* prune, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("densify_prune", stream);
    densify_prune_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, opacity_raw, scaling_raw, max_radii, min_opacity, max_screen_size,
                                                                 world_size_limit, use_size, prune); }
    GSR_CHECK_LAUNCH();
    return 0;
}

// ---- activation glue between the deformation network and the rasterizer (gaussian_renderer/__init__.py:79,116,122,140) ------
// render() turns the network's four heads into the rasterizer's inputs with ~10 torch element-wise kernels per view:
//   means3D = _xyz + dx;  scales = exp(_scaling + dscale);  rotations = normalize(_rotation + drot);
//   shs = cat(_features_dc, _features_rest) + dshs.reshape(-1, 16, 3)
// One pass here (488 B read, 232 B written per Gaussian), and one pass for the backward.
namespace {

__global__ void __launch_bounds__(256) glue_fwd_kernel(int P, const float* __restrict__ heads /*[P x 64]: dx 0-2, dscale 3-5, drot 6-9, dshs 10-57*/,
                                                      const float* __restrict__ xyz, const float* __restrict__ scaling, const float* __restrict__ rotation,
                                                      const float* __restrict__ f_dc /*[P,1,3]*/, const float* __restrict__ f_rest /*[P,15,3]*/,
                                                      float* __restrict__ means3D, float* __restrict__ scales, float* __restrict__ rotations,
                                                      float* __restrict__ shs /*[P,16,3]*/) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float4* h4 = reinterpret_cast<const float4*>(heads + 64 * (size_t)i);
    float h[64];
#pragma unroll
    for (int j = 0; j < 15; j++) { const float4 q = __ldg(h4 + j); h[4 * j] = q.x; h[4 * j + 1] = q.y; h[4 * j + 2] = q.z; h[4 * j + 3] = q.w; }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        means3D[3 * (size_t)i + k] = xyz[3 * (size_t)i + k] + h[k];
        scales[3 * (size_t)i + k] = expf(scaling[3 * (size_t)i + k] + h[3 + k]);
    }
    float q[4], n2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) { q[k] = rotation[4 * (size_t)i + k] + h[6 + k]; n2 += q[k] * q[k]; }
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);            // torch.nn.functional.normalize: x / max(|x|, eps)
#pragma unroll
    for (int k = 0; k < 4; k++) rotations[4 * (size_t)i + k] = q[k] * inv;
    float* o = shs + 48 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = f_dc[3 * (size_t)i + k] + h[10 + k];
#pragma unroll
    for (int k = 3; k < 48; k++) o[k] = f_rest[45 * (size_t)i + k - 3] + h[10 + k];
}

// gradients: d heads [P x 64] (columns 58-63 zero), d _xyz = g_means, d _scaling = g_scales * scales,
// d _rotation = (g - r (r.g)) / max(|q|, eps) with r = normalised q, d features = g_shs
__global__ void __launch_bounds__(256) glue_bwd_kernel(int P, const float* __restrict__ heads, const float* __restrict__ rotation,
                                                      const float* __restrict__ scales, const float* __restrict__ g_means,
                                                      const float* __restrict__ g_scales, const float* __restrict__ g_rot,
                                                      const float* __restrict__ g_shs, float* __restrict__ d_heads,
                                                      float* __restrict__ d_xyz, float* __restrict__ d_scaling, float* __restrict__ d_rotation,
                                                      float* __restrict__ d_f_dc, float* __restrict__ d_f_rest) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    float d[64];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float gm = g_means ? g_means[3 * (size_t)i + k] : 0.0f;
        const float gs = g_scales ? g_scales[3 * (size_t)i + k] * scales[3 * (size_t)i + k] : 0.0f;
        d[k] = gm; d[3 + k] = gs;
        if (d_xyz) d_xyz[3 * (size_t)i + k] = gm;
        if (d_scaling) d_scaling[3 * (size_t)i + k] = gs;
    }
    float q[4], g[4], n2 = 0.0f, rg = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) { q[k] = rotation[4 * (size_t)i + k] + heads[64 * (size_t)i + 6 + k]; n2 += q[k] * q[k]; g[k] = g_rot ? g_rot[4 * (size_t)i + k] : 0.0f; }
    const float nrm = sqrtf(n2);
    const float inv = 1.0f / fmaxf(nrm, 1e-12f);
#pragma unroll
    for (int k = 0; k < 4; k++) rg += q[k] * inv * g[k];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float dq = nrm > 1e-12f ? (g[k] - q[k] * inv * rg) * inv : g[k] * inv;
        d[6 + k] = dq;
        if (d_rotation) d_rotation[4 * (size_t)i + k] = dq;
    }
#pragma unroll
    for (int k = 0; k < 48; k++) {
        const float gv = g_shs ? g_shs[48 * (size_t)i + k] : 0.0f;
        d[10 + k] = gv;
        if (k < 3) { if (d_f_dc) d_f_dc[3 * (size_t)i + k] = gv; }
        else if (d_f_rest) d_f_rest[45 * (size_t)i + k - 3] = gv;
    }
#pragma unroll
    for (int k = 58; k < 64; k++) d[k] = 0.0f;
    float4* o = reinterpret_cast<float4*>(d_heads + 64 * (size_t)i);
#pragma unroll
    for (int j = 0; j < 16; j++) o[j] = make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
}

}  // namespace

int gsr_launch_deform_glue_fwd(int P, const float* heads, const float* xyz, const float* scaling, const float* rotation, const float* f_dc,
                               const float* f_rest, float* means3D, float* scales, float* rotations, float* shs, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("deform_glue_fwd", stream);
    glue_fwd_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, heads, xyz, scaling, rotation, f_dc, f_rest, means3D, scales, rotations, shs); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_deform_glue_bwd(int P, const float* heads, const float* rotation, const float* scales, const float* g_means, const float* g_scales,
                               const float* g_rot, const float* g_shs, float* d_heads, float* d_xyz, float* d_scaling, float* d_rotation,
                               float* d_f_dc, float* d_f_rest, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("deform_glue_bwd", stream);
    glue_bwd_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, heads, rotation, scales, g_means, g_scales, g_rot, g_shs, d_heads, d_xyz, d_scaling,
                                                            d_rotation, d_f_dc, d_f_rest); }
    GSR_CHECK_LAUNCH();
    return 0;
}
