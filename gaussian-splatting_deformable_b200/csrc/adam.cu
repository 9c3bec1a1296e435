// Fused Adam over ONE flat fp32 parameter buffer with per-group hyper-parameters.
//
// The reference optimises six Gaussian tensors (+ the deformation network) with torch.optim.Adam(lr=0,
// eps=1e-15) and one learning rate per group (scene/gaussian_model.py:834-886): with the parameters, their
// gradients (view_parallel.FlatGradBuffer - also the all-reduce buffer) and both moments laid out flat, the
// whole optimizer step is one pass at HBM speed: 16 B read + 12 B written per parameter (28 B x 66 x P).
// Arithmetic follows torch.optim.Adam (torch 2.11, `_single_tensor_adam`, no amsgrad / weight decay /
// maximize):   m = lerp(m, g, 1 - b1);  v = b2 v + (1 - b2) g g;
//              p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with the bias corrections computed on the host in double and rounded as torch rounds them.
#include "kernels.cuh"

namespace {

struct AdamGroups {
    int n;
    unsigned long long begin[GSR_ADAM_MAX_GROUPS + 1];      // element offsets, begin[n] = total
    float step_size[GSR_ADAM_MAX_GROUPS];                   // lr / bias_correction1
    float bc2_sqrt[GSR_ADAM_MAX_GROUPS];                // sqrt(bias_correction2)
    float one_minus_beta1[GSR_ADAM_MAX_GROUPS];             // float(1 - beta1) rounded from DOUBLE as torch rounds its scalars
    float beta2[GSR_ADAM_MAX_GROUPS], one_minus_beta2[GSR_ADAM_MAX_GROUPS], eps[GSR_ADAM_MAX_GROUPS];
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float omb1, float b2, float omb2, float eps,
                                         float step_size, float bc2_sqrt) {
    m = __fmaf_rn(omb1, g - m, m);                          // torch: exp_avg.lerp_(grad, 1 - beta1)
    v = __fmaf_rn(__fmul_rn(omb2, g), g, __fmul_rn(v, b2)); // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);   // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = __fmaf_rn(-step_size, __fdiv_rn(m, denom), p);      // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, AdamGroups G) {
    const unsigned long long total = G.begin[G.n];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * 4ull;
    for (unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 4ull; i < total; i += stride) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < GSR_ADAM_MAX_GROUPS; j++) k += (j < G.n && i >= G.begin[j]) ? 1 : 0;
        const bool whole = (i + 4ull <= G.begin[k + 1]);
        if (whole) {
            float4 pp = *reinterpret_cast<float4*>(p + i);
            const float4 gg = *reinterpret_cast<const float4*>(g + i);
            float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
            const float o1 = G.one_minus_beta1[k], b2 = G.beta2[k], o2 = G.one_minus_beta2[k], e = G.eps[k], ss = G.step_size[k], ib = G.bc2_sqrt[k];
            adam_one(pp.x, gg.x, mm.x, vv.x, o1, b2, o2, e, ss, ib);
            adam_one(pp.y, gg.y, mm.y, vv.y, o1, b2, o2, e, ss, ib);
            adam_one(pp.z, gg.z, mm.z, vv.z, o1, b2, o2, e, ss, ib);
            adam_one(pp.w, gg.w, mm.w, vv.w, o1, b2, o2, e, ss, ib);
            *reinterpret_cast<float4*>(p + i) = pp;
            *reinterpret_cast<float4*>(m + i) = mm;
            *reinterpret_cast<float4*>(v + i) = vv;
        } else {                                             // a group boundary (or the tail) inside these 4 elements
            for (unsigned long long e4 = i; e4 < i + 4ull && e4 < total; e4++) {
                int kk = 0;
#pragma unroll
                for (int j = 1; j < GSR_ADAM_MAX_GROUPS; j++) kk += (j < G.n && e4 >= G.begin[j]) ? 1 : 0;
                float pp = p[e4], mm = m[e4], vv = v[e4];
                adam_one(pp, g[e4], mm, vv, G.one_minus_beta1[kk], G.beta2[kk], G.one_minus_beta2[kk], G.eps[kk], G.step_size[kk], G.bc2_sqrt[kk]);
                p[e4] = pp; m[e4] = mm; v[e4] = vv;
            }
        }
    }
}

}  // namespace

int gsr_launch_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int num_groups,
                    const unsigned long long* begin /*[num_groups + 1]*/, const float* step_size, const float* bc2_sqrt,
                    const double* beta1, const double* beta2, const float* eps, cudaStream_t stream) {
    if (num_groups <= 0) return 0;
    if (num_groups > GSR_ADAM_MAX_GROUPS) return gsr_set_error_msg(-2, "adam: too many parameter groups");
    AdamGroups G{};
    G.n = num_groups;
    for (int k = 0; k <= num_groups; k++) G.begin[k] = begin[k];
    for (int k = 0; k < num_groups; k++) {
        if (begin[k + 1] < begin[k]) return gsr_set_error_msg(-2, "adam: group offsets must be non-decreasing");
        G.step_size[k] = step_size[k]; G.bc2_sqrt[k] = bc2_sqrt[k];
        G.one_minus_beta1[k] = (float)(1.0 - beta1[k]); G.beta2[k] = (float)beta2[k];
        G.one_minus_beta2[k] = (float)(1.0 - beta2[k]); G.eps[k] = eps[k];
    }
    const unsigned long long total = begin[num_groups];
    if (total == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return gsr_set_error_msg(-2, "adam: buffers must be 16-byte aligned");
    unsigned long long blocks = (total / 4 + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    if (blocks == 0) blocks = 1;
    { GsrProfScope prof_("adam_step", stream);
    adam_kernel<<<(unsigned)blocks, 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, G); }
    GSR_CHECK_LAUNCH();
    return 0;
}
