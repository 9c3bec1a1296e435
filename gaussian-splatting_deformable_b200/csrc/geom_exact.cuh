// Bit-exact per-Gaussian geometry: view/projection transform, 3D covariance,
// EWA 2D covariance, conic, radius, pixel centre and tile rectangle.
//
// Tile keys embed float bits of depth and a rect derived from ceil(3*sqrt(lambda))
// and an FP64 ndc->pixel map, and the blend decides alpha >= 1/255 on the conic,
// so these values must carry EXACTLY the roundings of the reference binary
// (forward.cu:74-152,192-237; auxiliary.h:41-76).  nvcc contracts the reference's
// glm expressions into a specific mix of FMUL/FFMA/FADD; that mix was read off
// the reference's PTX+SASS (tools/ptx_ssa.py, sm_100, nvcc 12.9.86) and is
// reproduced here with explicit round-to-nearest intrinsics, which the compiler
// may neither contract nor re-associate.  Rules observed:
//   a*b + c*d + e*f        -> fma(e,f, fma(a,b, c*d))          (glm mat3*mat3 element)
//   m0*x + m4*y + m8*z + m12 -> m12 + fma(z,m8, fma(x,m0, y*m4))
//   p*q - r*s (both products otherwise unused) -> fma(p,q, -(r*s))
#pragma once
#include "common.cuh"

#define FMUL(a, b) __fmul_rn((a), (b))
#define FADD(a, b) __fadd_rn((a), (b))
#define FSUB(a, b) __fsub_rn((a), (b))
#define FFMA(a, b, c) __fmaf_rn((a), (b), (c))

// a*b + c*d + e*f as nvcc emits it for a glm matrix-product element.
__device__ __forceinline__ float glm_dot3(float a, float b, float c, float d, float e, float f) {
    return FFMA(e, f, FFMA(a, b, FMUL(c, d)));
}

// Row `r` (0..3) of a 4x4 transform applied to [p,1]  (auxiliary.h:58-77).
__device__ __forceinline__ float xform_row(const float* m, int r, float3 p) {
    return FADD(m[12 + r], FFMA(p.z, m[8 + r], FFMA(p.x, m[r], FMUL(p.y, m[4 + r]))));
}

// forward.cu:118-152.  Returns the six unique entries of Sigma = M^T M with
// M = S * R(q); q = (r,x,y,z) is NOT normalised here (forward.cu:127).
__device__ __forceinline__ void cov3d_exact(float3 scale, float mod, float4 q, float* cov6) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float sx = FMUL(mod, scale.x), sy = FMUL(mod, scale.y), sz = FMUL(mod, scale.z);
    const float yy = FMUL(y, y), zz = FMUL(z, z);
    const float rz = FMUL(r, z), xz = FMUL(x, z), rx = FMUL(r, x);
    const float yy_zz = FADD(yy, zz);
    const float xx_zz = FFMA(x, x, zz);
    const float xx_yy = FFMA(x, x, yy);
    const float xy_m_rz = FFMA(x, y, -rz);
    const float xy_p_rz = FFMA(x, y, rz);
    const float xz_p_ry = FFMA(r, y, xz);
    const float xz_m_ry = FFMA(-r, y, xz);
    const float yz_m_rx = FFMA(y, z, -rx);
    const float yz_p_rx = FFMA(y, z, rx);
    // glm::mat3 constructor is column-major: R[c][r'] below is "column c, row r'".
    const float R00 = FSUB(1.0f, FADD(yy_zz, yy_zz));
    const float R01 = FADD(xy_m_rz, xy_m_rz);
    const float R02 = FADD(xz_p_ry, xz_p_ry);
    const float R10 = FADD(xy_p_rz, xy_p_rz);
    const float R11 = FSUB(1.0f, FADD(xx_zz, xx_zz));
    const float R12 = FADD(yz_m_rx, yz_m_rx);
    const float R20 = FADD(xz_m_ry, xz_m_ry);
    const float R21 = FADD(yz_p_rx, yz_p_rx);
    const float R22 = FSUB(1.0f, FADD(xx_yy, xx_yy));
    // M = S * R with S diagonal: M[c][r'] = s_r' * R[c][r'] (the 0*x terms of the
    // glm product add exact zeros).
    const float M00 = FMUL(sx, R00), M01 = FMUL(sy, R01), M02 = FMUL(sz, R02);
    const float M10 = FMUL(sx, R10), M11 = FMUL(sy, R11), M12 = FMUL(sz, R12);
    const float M20 = FMUL(sx, R20), M21 = FMUL(sy, R21), M22 = FMUL(sz, R22);
    // Sigma[c][r'] = M[r'][0]*M[c][0] + M[r'][1]*M[c][1] + M[r'][2]*M[c][2]
    cov6[0] = glm_dot3(M00, M00, M01, M01, M02, M02);  // Sigma[0][0]
    cov6[1] = glm_dot3(M10, M00, M11, M01, M12, M02);  // Sigma[0][1]
    cov6[2] = glm_dot3(M20, M00, M21, M01, M22, M02);  // Sigma[0][2]
    cov6[3] = glm_dot3(M10, M10, M11, M11, M12, M12);  // Sigma[1][1]
    cov6[4] = glm_dot3(M20, M10, M21, M11, M22, M12);  // Sigma[1][2]
    cov6[5] = glm_dot3(M20, M20, M21, M21, M22, M22);  // Sigma[2][2]
}

// The 2x3 non-zero part of T = W * J (forward.cu:80-99) for view-space point t.
struct EwaT { float T00, T01, T02, T10, T11, T12; float tx_ratio, ty_ratio; };

__device__ __forceinline__ EwaT ewa_T_exact(float3 t, const GsrView& v) {
    const float limx = FMUL(v.tan_fovx, 1.3f), limy = FMUL(v.tan_fovy, 1.3f);
    const float txtz = __fdiv_rn(t.x, t.z), tytz = __fdiv_rn(t.y, t.z);
    const float cx = fminf(limx, fmaxf(-limx, txtz));
    const float cy = fminf(limy, fmaxf(-limy, tytz));
    const float ntz = -t.z;
    const float tz2 = FMUL(t.z, t.z);
    const float J00 = __fdiv_rn(v.focal_x, t.z);
    const float J02 = __fdiv_rn(FMUL(v.focal_x, FMUL(cx, ntz)), tz2);   // -(fx * tx) / tz^2
    const float J11 = __fdiv_rn(v.focal_y, t.z);
    const float J12 = __fdiv_rn(FMUL(v.focal_y, FMUL(cy, ntz)), tz2);
    const float* V = v.view;
    EwaT o;
    // T[0][r] = W[0][r]*J00 + W[1][r]*0 + W[2][r]*J02, W[k][r] = V[k + 4r]
    o.T00 = FFMA(V[2], J02, FMUL(V[0], J00));
    o.T01 = FFMA(V[6], J02, FMUL(V[4], J00));
    o.T02 = FFMA(V[10], J02, FMUL(V[8], J00));
    // T[1][r] = W[0][r]*0 + W[1][r]*J11 + W[2][r]*J12
    o.T10 = FFMA(V[2], J12, FMUL(J11, V[1]));
    o.T11 = FFMA(V[6], J12, FMUL(J11, V[5]));
    o.T12 = FFMA(V[10], J12, FMUL(J11, V[9]));
    o.tx_ratio = txtz; o.ty_ratio = tytz;
    return o;
}

// cov2D = T^T Vrk^T T, upper-left 2x2, with the +0.3 low-pass (forward.cu:101-112).
__device__ __forceinline__ float3 cov2d_exact(const EwaT& e, const float* c) {
    const float X00 = glm_dot3(e.T00, c[0], e.T01, c[1], e.T02, c[2]);
    const float X01 = glm_dot3(e.T10, c[0], e.T11, c[1], e.T12, c[2]);
    const float X10 = glm_dot3(e.T00, c[1], e.T01, c[3], e.T02, c[4]);
    const float X11 = glm_dot3(e.T10, c[1], e.T11, c[3], e.T12, c[4]);
    const float X20 = glm_dot3(e.T00, c[2], e.T01, c[4], e.T02, c[5]);
    const float X21 = glm_dot3(e.T10, c[2], e.T11, c[4], e.T12, c[5]);
    const float cov00 = glm_dot3(e.T00, X00, e.T01, X10, e.T02, X20);
    const float cov01 = glm_dot3(e.T00, X01, e.T01, X11, e.T02, X21);
    const float cov11 = glm_dot3(e.T10, X01, e.T11, X11, e.T12, X21);
    return make_float3(FADD(cov00, 0.3f), cov01, FADD(cov11, 0.3f));
}

// auxiliary.h:41-44, evaluated in double: ((v + 1.0) * S - 1.0) * 0.5
__device__ __forceinline__ float ndc2pix_exact(float v, int S) {
    return __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5));
}

// auxiliary.h:46-56 (BLOCK_X = BLOCK_Y = 16; the /16 is an exact *0.0625).
__device__ __forceinline__ void tile_rect_exact(float2 p, int radius, int grid_x, int grid_y,
                                                uint2& rmin, uint2& rmax) {
    const float rf = (float)radius;
    rmin.x = min((unsigned)grid_x, (unsigned)max(0, (int)FMUL(FSUB(p.x, rf), 0.0625f)));
    rmin.y = min((unsigned)grid_y, (unsigned)max(0, (int)FMUL(FSUB(p.y, rf), 0.0625f)));
    rmax.x = min((unsigned)grid_x, (unsigned)max(0, (int)FMUL(FADD(FADD(FADD(p.x, rf), 16.0f), -1.0f), 0.0625f)));
    rmax.y = min((unsigned)grid_y, (unsigned)max(0, (int)FMUL(FADD(FADD(FADD(p.y, rf), 16.0f), -1.0f), 0.0625f)));
}

// Everything the binning and blend stages need from one Gaussian's geometry.
struct SplatGeom {
    float depth;        // p_view.z
    float2 pix;         // pixel centre
    float3 cov;         // (a, b, c) incl. low-pass
    float3 conic;
    int radius;
    uint2 rmin, rmax;
    bool ok;            // passed near-cull, det != 0 and non-empty rect
};

// forward.cu:186-237 for one already-deformed mean `p`.
__device__ __forceinline__ SplatGeom splat_geometry_exact(float3 p, const float* cov6, const GsrView& v) {
    SplatGeom g;
    g.ok = false; g.radius = 0;
    g.depth = xform_row(v.view, 2, p);
    if (g.depth <= GSR_NEAR) return g;
    const float hx = xform_row(v.proj, 0, p), hy = xform_row(v.proj, 1, p), hw = xform_row(v.proj, 3, p);
    const float p_w = __frcp_rn(FADD(hw, 0.0000001f));
    const float projx = FMUL(hx, p_w), projy = FMUL(hy, p_w);
    float3 t = make_float3(xform_row(v.view, 0, p), xform_row(v.view, 1, p), g.depth);
    const EwaT e = ewa_T_exact(t, v);
    g.cov = cov2d_exact(e, cov6);
    const float det = FFMA(g.cov.x, g.cov.z, -FMUL(g.cov.y, g.cov.y));
    if (det == 0.0f) return g;
    const float det_inv = __frcp_rn(det);
    g.conic = make_float3(FMUL(g.cov.z, det_inv), FMUL(det_inv, -g.cov.y), FMUL(g.cov.x, det_inv));
    const float mid = FMUL(FADD(g.cov.x, g.cov.z), 0.5f);
    const float root = __fsqrt_rn(fmaxf(FFMA(mid, mid, -det), 0.1f));
    const float lam = fmaxf(FADD(mid, root), FSUB(mid, root));
    const float my_radius = ceilf(FMUL(__fsqrt_rn(lam), 3.0f));
    g.pix = make_float2(ndc2pix_exact(projx, v.W), ndc2pix_exact(projy, v.H));
    const int radius = (int)my_radius;
    tile_rect_exact(g.pix, radius, v.grid_x, v.grid_y, g.rmin, g.rmax);
    if ((g.rmax.x - g.rmin.x) * (g.rmax.y - g.rmin.y) == 0) return g;
    g.radius = radius;
    g.ok = true;
    return g;
}

// Blend exponent in the reference's rounding (forward.cu:335 as compiled):
//   power = fma(fma(dx, dx*cx, dy*(dy*cz)), -0.5, -(dy*(dx*cy)))
__device__ __forceinline__ float blend_power_exact(float dx, float dy, float cx, float cy, float cz) {
    return FFMA(FFMA(dx, FMUL(dx, cx), FMUL(dy, FMUL(dy, cz))), -0.5f, -FMUL(dy, FMUL(dx, cy)));
}
