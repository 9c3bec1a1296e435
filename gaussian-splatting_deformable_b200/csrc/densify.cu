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
