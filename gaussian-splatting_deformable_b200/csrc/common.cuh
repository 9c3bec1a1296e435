// Shared declarations for the sm_100a deformable-Gaussian rasterizer kernels.
//
// Everything here is device-side plumbing for the C-ABI in include/gsr_b200.h.
// No torch, no glm, no CUB: the library links against the CUDA runtime only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GSR_TILE 16              // reference: BLOCK_X = BLOCK_Y = 16 (cuda_rasterizer/config.h:16-17)
#define GSR_TILE_PIX 256
#define GSR_NEAR 0.2f            // reference near plane (auxiliary.h:154)

// ---------------------------------------------------------------------------
// Camera / view constants, passed by value (lives in the constant bank, so the
// 35 matrix words are LDC broadcasts instead of global loads).
// Matrices keep the reference's memory convention: 16 floats, true row i is
// (m[i], m[i+4], m[i+8], m[i+12])  (auxiliary.h:58-76).
// ---------------------------------------------------------------------------
struct GsrView {
    float view[16];
    float proj[16];
    float campos[3];
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;      // W / (2 tan_fovx), H / (2 tan_fovy)  (rasterizer_impl.cu:222-223)
    float scale_modifier;
    int W, H;
    int grid_x, grid_y;          // ceil(W/16), ceil(H/16)
    int sh_degree;               // active degree D
    int sh_coeffs;               // M = coefficients stored per Gaussian (0 => no SH)
};

// Per-Gaussian "splat record" consumed by the blend kernels: 3 x float4 = 48 B,
// contiguous so one list entry is one 48-byte gather (one TMA bulk copy).
//   q0 = (x, y, conic.x, conic.y)   q1 = (conic.z, opacity, r, g)
//   q2 = (b, power_cut, 0, 0)
// power_cut: a pixel whose power is below it can never reach alpha >= 1/255
// (log(1/(255*opacity)) minus a safety margin) -- lets the blend loops reject
// without evaluating expf, with decisions identical to the reference's.
#define GSR_REC_F4 3

// 12-float gradient record accumulated by blend backward, consumed by the
// fused per-Gaussian backward:
//   g0 = (dL_dmean2D.x, dL_dmean2D.y, dL_dconic.x, dL_dconic.y)
//   g1 = (dL_dconic.w(zz), dL_dopacity, dL_dcolor.r, dL_dcolor.g)
//   g2 = (dL_dcolor.b, 0, 0, 0)
// Moment form (blend_v2.cu): with w = G * dL/dalpha per blended (pixel, Gaussian) pair and
// d = mean2D - pixel, the record accumulates
//   g0 = (sum w, sum w dx, sum w dy, sum w dx^2)   g1 = (sum w dx dy, sum w dy^2, dL_dcolor.r, dL_dcolor.g)
//   g2 = (dL_dcolor.b, 0, 0, 0)
// and the per-Gaussian backward turns the six moments into dL_dopacity, dL_dmean2D and dL_dconic
// with the Gaussian's own opacity and conic (linear, so it commutes with the sum over pixels):
// ten fewer instructions per pair in the issue-bound blend loop.
#define GSR_GRAD_F4 3

// Deformation modes fused into preprocess (scene/rigid_body.py:86-93 + the
// apply recipe gaussian_renderer/__init__.py:92-95).
#define GSR_DEFORM_NONE 0
#define GSR_DEFORM_PER_GAUSSIAN 1   // S[P,6], theta[P]
#define GSR_DEFORM_RIGID_BODIES 2   // body_id[P], S[B,6], theta[B]

#define GSR_CHECK(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return gsr_set_error(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define GSR_CHECK_LAUNCH() GSR_CHECK(cudaGetLastError())

int gsr_set_error(cudaError_t e, const char* what, const char* file, int line);
int gsr_set_error_msg(int code, const char* msg);

// Launch accounting + optional per-kernel timing with CUDA events on the launching
// stream (gsr_profile_enable / gsr_profile_dump in include/gsr_b200.h).
struct GsrProfScope {
    GsrProfScope(const char* name, cudaStream_t stream);
    ~GsrProfScope();
    int slot; cudaStream_t stream; int pending;
};

static inline int gsr_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// 256-bit global loads/stores (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256).  A thread that
// owns a 32-byte-aligned chunk moves one whole DRAM sector per instruction, so AoS
// records (192-byte SH, its gradient) are read without the L2->L1 sector re-fetch that
// 16-byte accesses to the same sector from different instructions cause.
__device__ __forceinline__ void ld256_nc(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld256(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// [P,4] records (quaternions and their gradients) move as one 128-bit access when the array is
// 16-byte aligned and as four scalars otherwise: the caller's tensors may be views into a packed
// flat buffer (fused_adam.FusedAdam / view_parallel.FlatGradBuffer) at any 4-byte offset.
__device__ __forceinline__ float4 ld_rec4(const float* base, size_t idx) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) return reinterpret_cast<const float4*>(base)[idx];
    const float* p = base + 4 * idx;
    return make_float4(p[0], p[1], p[2], p[3]);
}
__device__ __forceinline__ void st_rec4(float* base, size_t idx, float4 o, bool accumulate) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) {
        float4* d = reinterpret_cast<float4*>(base) + idx;
        if (accumulate) atomicAdd(d, o); else *d = o;
    } else {
        float* p = base + 4 * idx;
        if (accumulate) { atomicAdd(p, o.x); atomicAdd(p + 1, o.y); atomicAdd(p + 2, o.z); atomicAdd(p + 3, o.w); }
        else { p[0] = o.x; p[1] = o.y; p[2] = o.z; p[3] = o.w; }
    }
}

// ---------------------------------------------------------------------------
// SE3 exponential map applied to a point (closed form of rigid_body.exp_se3
// followed by y = (T [x;1])[:3]):
//   R = I + sin(th) W + (1-cos(th)) W^2
//   p = (th I + (1-cos(th)) W + (th - sin(th)) W^2) v
//   y = R x + p,  with  W u = w x u  and  W^2 u = w (w.u) - (w.w) u
// ---------------------------------------------------------------------------
struct Se3Coef { float a, b, c; };  // sin(th), 1-cos(th), th-sin(th)

__device__ __forceinline__ Se3Coef se3_coef(float th) {
    float s, c;
    sincosf(th, &s, &c);
    Se3Coef k; k.a = s; k.b = 1.0f - c; k.c = th - s; return k;
}
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// W^2 u
__device__ __forceinline__ float3 skew2(float3 w, float3 u) {
    float wu = dot3(w, u), ww = dot3(w, w);
    return make_float3(w.x * wu - ww * u.x, w.y * wu - ww * u.y, w.z * wu - ww * u.z);
}
__device__ __forceinline__ float3 se3_apply(float3 x, float3 w, float3 v, float th) {
    Se3Coef k = se3_coef(th);
    float3 wx = cross3(w, x), wwx = skew2(w, x);
    float3 wv = cross3(w, v), wwv = skew2(w, v);
    float3 y;
    y.x = x.x + k.a * wx.x + k.b * wwx.x + th * v.x + k.b * wv.x + k.c * wwv.x;
    y.y = x.y + k.a * wx.y + k.b * wwx.y + th * v.y + k.b * wv.y + k.c * wwv.y;
    y.z = x.z + k.a * wx.z + k.b * wwx.z + th * v.z + k.b * wv.z + k.c * wwv.z;
    return y;
}
