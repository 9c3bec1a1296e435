// Fused per-Gaussian backward: conic/cov2D, projection, SH, cov3D and SE3.
//
// Replaces, in ONE pass over the Gaussians (the reference makes two kernel
// passes plus nine torch::zeros fills, rasterize_points.cu:151-159):
//   backward.cu:144-274  computeCov2DCUDA
//   backward.cu:346-396  preprocessCUDA (bwd) with SH bwd :20-139, cov3D bwd :278-341
//   autograd of scene/rigid_body.py:86-93 + the apply recipe (closed form)
// Every output element is written here (zeros for culled Gaussians), so no
// pre-zeroing of the 75-float-per-Gaussian gradient tensors is needed.
// HBM-bound: ~280 B read + ~236 B written per Gaussian.
#include "geom_exact.cuh"
#include "kernels.cuh"
#include <stdlib.h>

namespace {

__device__ __constant__ float bSH_C0 = 0.28209479177387814f;
__device__ __constant__ float bSH_C1 = 0.4886025119029199f;
__device__ __constant__ float bSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float bSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator*(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// d(v/|v|)/dv applied to dv  (auxiliary.h:107-117)
__device__ __forceinline__ V3 dnormvdv(V3 v, V3 dv) {
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    V3 o;
    o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return o;
}


// Output of one value: plain store, or (accumulate mode) a fire-and-forget RED so that
// no load of the running gradient sits in the thread's dependency chain.
__device__ __forceinline__ void emit(float* p, float val, bool accumulate) {
    if (accumulate) atomicAdd(p, val); else *p = val;
}

// Structure: ONE dependent chain of two DRAM round trips per thread.  (1) radii -> visible;
// (2) every input of a visible Gaussian is loaded back to back (gradient record, mean, scale,
// quaternion, clamp bits, the 192-byte SH record as 6 x LDG.256, twist) BEFORE the first store -
// a store to a non-restrict pointer would otherwise pin all later loads behind it; (3) math;
// (4) all stores.  Accumulate mode uses vector reductions (RED.ADD.F32x4 for the SH gradient).
template <int MINB>
__global__ void __launch_bounds__(256, MINB) preprocess_bwd_kernel(PreprocessBwdArgs a, GsrView v) {
    extern __shared__ float s_body[];   // [num_bodies][7] when accumulating rigid-body twists in smem
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const bool body_smem = (a.deform_mode == GSR_DEFORM_RIGID_BODIES) && a.dL_dtwist_S && (a.num_bodies * 7 * 4 <= 32768);
    if (body_smem) {
        for (int k = threadIdx.x; k < a.num_bodies * 7; k += 256) s_body[k] = 0.0f;
        __syncthreads();
    }
    const int M = v.sh_coeffs;
    const int acc = a.acc;
    const bool visible = (idx < a.P) && (a.radii[idx] > 0);
    if (idx < a.P && !visible) {
        // Culled: zeros to every output that is not a running sum; nothing is read.
        a.dL_dmeans2D[3 * idx] = 0.0f; a.dL_dmeans2D[3 * idx + 1] = 0.0f; a.dL_dmeans2D[3 * idx + 2] = 0.0f;
        if (!(acc & GSR_ACC_OPACITY)) a.dL_dopacity[idx] = 0.0f;
        if (a.dL_dcolors) { a.dL_dcolors[3 * idx] = 0.0f; a.dL_dcolors[3 * idx + 1] = 0.0f; a.dL_dcolors[3 * idx + 2] = 0.0f; }
        if (a.dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * (size_t)idx + k] = 0.0f;
        }
        if (a.dL_dsh && !(acc & GSR_ACC_SH)) {
            float* dsh = a.dL_dsh + (size_t)idx * M * 3;
            if (M == 16 && ((reinterpret_cast<uintptr_t>(a.dL_dsh) & 31) == 0)) {
                const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 6; k++) st256(dsh + 8 * k, z8);
            } else {
                for (int k = 0; k < M * 3; k++) dsh[k] = 0.0f;
            }
        }
        if (a.dL_dscales && !(acc & GSR_ACC_SCALES)) { a.dL_dscales[3 * idx] = 0.0f; a.dL_dscales[3 * idx + 1] = 0.0f; a.dL_dscales[3 * idx + 2] = 0.0f; }
        if (a.dL_drots && !(acc & GSR_ACC_ROTS)) st_rec4(a.dL_drots, idx, make_float4(0.f, 0.f, 0.f, 0.f), false);
        if (a.dL_dtwist_S && a.deform_mode == GSR_DEFORM_PER_GAUSSIAN && !(acc & GSR_ACC_TWIST)) {
#pragma unroll
            for (int k = 0; k < 6; k++) a.dL_dtwist_S[6 * (size_t)idx + k] = 0.0f;
            a.dL_dtwist_theta[idx] = 0.0f;
        }
        if (!(acc & GSR_ACC_MEANS3D)) { a.dL_dmeans3D[3 * idx] = 0.0f; a.dL_dmeans3D[3 * idx + 1] = 0.0f; a.dL_dmeans3D[3 * idx + 2] = 0.0f; }
    }
    if (visible) {
        // ================= (2) all loads =================
        const float4* gr = a.grad_recs + 3 * (size_t)idx;
        float4 g0 = gr[0], g1 = gr[1];
        const float g2x = reinterpret_cast<const float*>(gr + 2)[0];
        float4 sp0 = make_float4(0, 0, 0, 0);
        float2 sp1 = make_float2(0, 0);
        if (a.grad_moments) {
            const float4* sp = a.recs + 3 * (size_t)idx;
            sp0 = sp[0]; sp1 = *reinterpret_cast<const float2*>(sp + 1);
        }
        const float* mp = (a.means_deformed ? a.means_deformed : a.means) + 3 * (size_t)idx;
        const float3 mean = make_float3(mp[0], mp[1], mp[2]);
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
        const int deg = v.sh_degree;
        float shv[48];
        uint8_t cl = 0;
        const bool wide = a.shs && (M == 16) &&
                          (((reinterpret_cast<uintptr_t>(a.shs) | reinterpret_cast<uintptr_t>(a.dL_dsh)) & 31) == 0);
        if (a.shs) {
            cl = a.clamped[idx];
            const float* shp = a.shs + (size_t)idx * M * 3;
            if (wide) {
                const int need = (deg + 1) * (deg + 1) * 3;
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    if (8 * k < need) ld256_nc(shp + 8 * k, shv + 8 * k);
                    else {
#pragma unroll
                        for (int e = 0; e < 8; e++) shv[8 * k + e] = 0.0f;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 48; k++) shv[k] = (k < M * 3) ? shp[k] : 0.0f;
            }
        }
        const int tix = (a.deform_mode == GSR_DEFORM_RIGID_BODIES) ? a.body_id[idx] : idx;
        float3 w = make_float3(0, 0, 0), tv = w, x0 = w;
        float th = 0.0f;
        if (a.deform_mode != GSR_DEFORM_NONE) {
            const float* S = a.twist_S + 6 * (size_t)tix;
            w = make_float3(S[0], S[1], S[2]); tv = make_float3(S[3], S[4], S[5]);
            th = a.twist_theta[tix];
            x0 = make_float3(a.means[3 * idx], a.means[3 * idx + 1], a.means[3 * idx + 2]);
        }

        // ================= (3) math =================
        float gm[3];                            // dL/d(deformed mean)
        float dcov[6] = {0, 0, 0, 0, 0, 0};
        float dscale[3] = {0, 0, 0};
        float drot[4] = {0, 0, 0, 0};
        // ---- computeCov2DCUDA (backward.cu:144-274) ----
        if (!a.cov3D_precomp) cov3d_exact(s, v.scale_modifier, q, cov6);   // recomputed: saves a 24 B/Gaussian round trip
        float3 t = make_float3(xform_row(v.view, 0, mean), xform_row(v.view, 1, mean), xform_row(v.view, 2, mean));
        const float limx = 1.3f * v.tan_fovx, limy = 1.3f * v.tan_fovy;
        const float txtz = t.x / t.z, tytz = t.y / t.z;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
        const EwaT e = ewa_T_exact(t, v);
        t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
        t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
        const float3 cov = cov2d_exact(e, cov6);
        const float ca = cov.x, cb = cov.y, cc = cov.z;
        const float denom = ca * cc - cb * cb;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (a.grad_moments) {
            // moments -> (dL_dmean2D.xy, dL_dconic.xyz, dL_dopacity); common.cuh, backward.cu:504-528
            const float M0 = g0.x, Mx = g0.y, My = g0.z, Mxx = g0.w, Mxy = g1.x, Myy = g1.y;
            const float cx = sp0.z, cy = sp0.w, cz = sp1.x, op = sp1.y;
            const float hop = -0.5f * op;
            g0.x = -op * a.half_W * (cx * Mx + cy * My);
            g0.y = -op * a.half_H * (cz * My + cy * Mx);
            g0.z = hop * Mxx; g0.w = hop * Mxy; g1.x = hop * Myy;
            g1.y = M0;
        }
        const float dcx = g0.z, dcy = g0.w, dcz = g1.x;   // dL_dconic (xx, xy, yy)
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        // glm column-major T[c][r]: T[0][*] = (T00,T01,T02), T[1][*] = (T10,T11,T12)
        const float T00 = e.T00, T01 = e.T01, T02 = e.T02, T10 = e.T10, T11 = e.T11, T12 = e.T12;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-cc * cc * dcx + 2 * cb * cc * dcy + (denom - ca * cc) * dcz);
            dL_dc = denom2inv * (-ca * ca * dcz + 2 * ca * cb * dcy + (denom - ca * cc) * dcx);
            dL_db = denom2inv * 2 * (cb * cc * dcx - (denom + 2 * cb * cb) * dcy + ca * cb * dcz);
            dcov[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
            dcov[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
            dcov[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
            dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
            dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
            dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
        }
        // Vrk (symmetric) rows
        const float V00 = cov6[0], V01 = cov6[1], V02 = cov6[2], V11 = cov6[3], V12 = cov6[4], V22 = cov6[5];
        const float dL_dT00 = 2 * (T00 * V00 + T01 * V01 + T02 * V02) * dL_da + (T10 * V00 + T11 * V01 + T12 * V02) * dL_db;
        const float dL_dT01 = 2 * (T00 * V01 + T01 * V11 + T02 * V12) * dL_da + (T10 * V01 + T11 * V11 + T12 * V12) * dL_db;
        const float dL_dT02 = 2 * (T00 * V02 + T01 * V12 + T02 * V22) * dL_da + (T10 * V02 + T11 * V12 + T12 * V22) * dL_db;
        const float dL_dT10 = 2 * (T10 * V00 + T11 * V01 + T12 * V02) * dL_dc + (T00 * V00 + T01 * V01 + T02 * V02) * dL_db;
        const float dL_dT11 = 2 * (T10 * V01 + T11 * V11 + T12 * V12) * dL_dc + (T00 * V01 + T01 * V11 + T02 * V12) * dL_db;
        const float dL_dT12 = 2 * (T10 * V02 + T11 * V12 + T12 * V22) * dL_dc + (T00 * V02 + T01 * V12 + T02 * V22) * dL_db;
        // W[c][r] = view[c + 4 r]
        const float* Vm = v.view;
        const float dL_dJ00 = Vm[0] * dL_dT00 + Vm[4] * dL_dT01 + Vm[8] * dL_dT02;
        const float dL_dJ02 = Vm[2] * dL_dT00 + Vm[6] * dL_dT01 + Vm[10] * dL_dT02;
        const float dL_dJ11 = Vm[1] * dL_dT10 + Vm[5] * dL_dT11 + Vm[9] * dL_dT12;
        const float dL_dJ12 = Vm[2] * dL_dT10 + Vm[6] * dL_dT11 + Vm[10] * dL_dT12;
        const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
        const float h_x = v.focal_x, h_y = v.focal_y;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                             (2 * h_y * t.y) * tz3 * dL_dJ12;
        // transformVec4x3Transpose (auxiliary.h:89-97)
        gm[0] = Vm[0] * dL_dtx + Vm[1] * dL_dty + Vm[2] * dL_dtz;
        gm[1] = Vm[4] * dL_dtx + Vm[5] * dL_dty + Vm[6] * dL_dtz;
        gm[2] = Vm[8] * dL_dtx + Vm[9] * dL_dty + Vm[10] * dL_dtz;

        // ---- projection part (backward.cu:370-387) ----
        const float* proj = v.proj;
        const float m_hom_w = proj[3] * mean.x + proj[7] * mean.y + proj[11] * mean.z + proj[15];
        const float m_w = 1.0f / (m_hom_w + 0.0000001f);
        const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        gm[0] += (proj[0] * m_w - proj[3] * mul1) * g0.x + (proj[1] * m_w - proj[3] * mul2) * g0.y;
        gm[1] += (proj[4] * m_w - proj[7] * mul1) * g0.x + (proj[5] * m_w - proj[7] * mul2) * g0.y;
        gm[2] += (proj[8] * m_w - proj[11] * mul1) * g0.x + (proj[9] * m_w - proj[11] * mul2) * g0.y;

        // ---- SH backward (backward.cu:20-139) ----
        float coef[16];     // dRGB/dsh_k is a scalar per coefficient: dL_dsh[k] = coef[k] * dL_dRGB
        V3 dRGB = {0, 0, 0};
        if (a.shs) {
            dRGB = {(cl & 1) ? 0.0f : g1.z, (cl & 2) ? 0.0f : g1.w, (cl & 4) ? 0.0f : g2x};
            auto sh = [&](int k) -> V3 { return {shv[3 * k], shv[3 * k + 1], shv[3 * k + 2]}; };
            V3 dir_orig = {mean.x - v.campos[0], mean.y - v.campos[1], mean.z - v.campos[2]};
            const float len = sqrtf(dot(dir_orig, dir_orig));
            const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
            V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
#pragma unroll
            for (int k = 0; k < 16; k++) coef[k] = 0.0f;
            coef[0] = bSH_C0;
            if (deg > 0) {
                coef[1] = -bSH_C1 * y; coef[2] = bSH_C1 * z; coef[3] = -bSH_C1 * x;
                dRGBdx = (-bSH_C1) * sh(3);
                dRGBdy = (-bSH_C1) * sh(1);
                dRGBdz = bSH_C1 * sh(2);
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    coef[4] = bSH_C2[0] * xy; coef[5] = bSH_C2[1] * yz; coef[6] = bSH_C2[2] * (2.f * zz - xx - yy);
                    coef[7] = bSH_C2[3] * xz; coef[8] = bSH_C2[4] * (xx - yy);
                    const V3 s4 = sh(4), s5 = sh(5), s6 = sh(6), s7 = sh(7), s8 = sh(8);
                    dRGBdx = dRGBdx + (bSH_C2[0] * y) * s4 + (bSH_C2[2] * 2.f * -x) * s6 + (bSH_C2[3] * z) * s7 + (bSH_C2[4] * 2.f * x) * s8;
                    dRGBdy = dRGBdy + (bSH_C2[0] * x) * s4 + (bSH_C2[1] * z) * s5 + (bSH_C2[2] * 2.f * -y) * s6 + (bSH_C2[4] * 2.f * -y) * s8;
                    dRGBdz = dRGBdz + (bSH_C2[1] * y) * s5 + (bSH_C2[2] * 2.f * 2.f * z) * s6 + (bSH_C2[3] * x) * s7;
                    if (deg > 2) {
                        coef[9] = bSH_C3[0] * y * (3.f * xx - yy); coef[10] = bSH_C3[1] * xy * z;
                        coef[11] = bSH_C3[2] * y * (4.f * zz - xx - yy);
                        coef[12] = bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                        coef[13] = bSH_C3[4] * x * (4.f * zz - xx - yy); coef[14] = bSH_C3[5] * z * (xx - yy);
                        coef[15] = bSH_C3[6] * x * (xx - 3.f * yy);
                        const V3 s9 = sh(9), s10 = sh(10), s11 = sh(11), s12 = sh(12), s13 = sh(13), s14 = sh(14), s15 = sh(15);
                        dRGBdx = dRGBdx + (bSH_C3[0] * 3.f * 2.f * xy) * s9 + (bSH_C3[1] * yz) * s10 + (bSH_C3[2] * -2.f * xy) * s11 +
                                 (bSH_C3[3] * -3.f * 2.f * xz) * s12 + (bSH_C3[4] * (-3.f * xx + 4.f * zz - yy)) * s13 +
                                 (bSH_C3[5] * 2.f * xz) * s14 + (bSH_C3[6] * 3.f * (xx - yy)) * s15;
                        dRGBdy = dRGBdy + (bSH_C3[0] * 3.f * (xx - yy)) * s9 + (bSH_C3[1] * xz) * s10 +
                                 (bSH_C3[2] * (-3.f * yy + 4.f * zz - xx)) * s11 + (bSH_C3[3] * -3.f * 2.f * yz) * s12 +
                                 (bSH_C3[4] * -2.f * xy) * s13 + (bSH_C3[5] * -2.f * yz) * s14 + (bSH_C3[6] * -3.f * 2.f * xy) * s15;
                        dRGBdz = dRGBdz + (bSH_C3[1] * xy) * s10 + (bSH_C3[2] * 4.f * 2.f * yz) * s11 +
                                 (bSH_C3[3] * 3.f * (2.f * zz - xx - yy)) * s12 + (bSH_C3[4] * 4.f * 2.f * xz) * s13 +
                                 (bSH_C3[5] * (xx - yy)) * s14;
                    }
                }
            }
            const V3 dL_ddir = {dot(dRGBdx, dRGB), dot(dRGBdy, dRGB), dot(dRGBdz, dRGB)};
            const V3 dmean_sh = dnormvdv(dir_orig, dL_ddir);
            gm[0] += dmean_sh.x; gm[1] += dmean_sh.y; gm[2] += dmean_sh.z;
        }

        // ---- cov3D backward (backward.cu:278-341) ----
        if (a.scales) {
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            // R as glm columns; Rt[k] = row k of that column-major matrix (= column k of transpose)
            const float R[3][3] = {
                {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};   // R[c][r']
            const float sv[3] = {v.scale_modifier * s.x, v.scale_modifier * s.y, v.scale_modifier * s.z};
            float Mm[3][3];     // M = S * R : M[c][r'] = sv[r'] * R[c][r']
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int rr = 0; rr < 3; rr++) Mm[c][rr] = sv[rr] * R[c][rr];
            // dL_dSigma (column-major, symmetric)
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            // dL_dM = 2 * M * dL_dSigma  (glm product: (A*B)[c][r'] = sum_k A[k][r'] B[c][k])
            float dM[3][3];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int rr = 0; rr < 3; rr++)
                    dM[c][rr] = 2.0f * (Mm[0][rr] * dS[c][0] + Mm[1][rr] * dS[c][1] + Mm[2][rr] * dS[c][2]);
            // Rt = transpose(R): Rt[k][j] = R[j][k];  dL_dMt[k][j] = dM[j][k]
            float dMt[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < 3; j++) dMt[k][j] = dM[j][k];
#pragma unroll
            for (int k = 0; k < 3; k++)
                dscale[k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < 3; j++) dMt[k][j] *= sv[k];
            drot[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            drot[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            drot[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            drot[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }

        // ---- SE3 backward (closed form) ----
        float gx[3] = {gm[0], gm[1], gm[2]};
        float dS6[6] = {0, 0, 0, 0, 0, 0};
        float dth = 0.0f;
        if (a.deform_mode != GSR_DEFORM_NONE) {
            const float3 g = make_float3(gm[0], gm[1], gm[2]);
            float sn, cs;
            sincosf(th, &sn, &cs);
            const float ka = sn, kb = 1.0f - cs, kc = th - sn;
            const float3 wg = cross3(w, g), wwg = skew2(w, g);
            // dL/dx = R^T g = g - a (w x g) + b W^2 g
            gx[0] = g.x - ka * wg.x + kb * wwg.x;
            gx[1] = g.y - ka * wg.y + kb * wwg.y;
            gx[2] = g.z - ka * wg.z + kb * wwg.z;
            // dL/dv = th g - b (w x g) + c W^2 g
            dS6[3] = th * g.x - kb * wg.x + kc * wwg.x;
            dS6[4] = th * g.y - kb * wg.y + kc * wwg.y;
            dS6[5] = th * g.z - kb * wg.z + kc * wwg.z;
            // dL/dth = g . ( cos (w x x) + sin W^2 x + v + sin (w x v) + (1-cos) W^2 v )
            const float3 wx = cross3(w, x0), wwx = skew2(w, x0), wv = cross3(w, tv), wwv = skew2(w, tv);
            dth = g.x * (cs * wx.x + sn * wwx.x + tv.x + sn * wv.x + kb * wwv.x) +
                  g.y * (cs * wx.y + sn * wwx.y + tv.y + sn * wv.y + kb * wwv.y) +
                  g.z * (cs * wx.z + sn * wwx.z + tv.z + sn * wv.z + kb * wwv.z);
            // dL/dw = a (x x g) + b D(x) + b (v x g) + c D(v),  D(u) = g (w.u) + u (w.g) - 2 w (g.u)
            const float3 xg = cross3(x0, g), vg = cross3(tv, g);
            const float wdx = dot3(w, x0), wdv = dot3(w, tv), wdg = dot3(w, g), gdx = dot3(g, x0), gdv = dot3(g, tv);
            const float3 Dx = make_float3(g.x * wdx + x0.x * wdg - 2.f * w.x * gdx, g.y * wdx + x0.y * wdg - 2.f * w.y * gdx,
                                          g.z * wdx + x0.z * wdg - 2.f * w.z * gdx);
            const float3 Dv = make_float3(g.x * wdv + tv.x * wdg - 2.f * w.x * gdv, g.y * wdv + tv.y * wdg - 2.f * w.y * gdv,
                                          g.z * wdv + tv.z * wdg - 2.f * w.z * gdv);
            dS6[0] = ka * xg.x + kb * Dx.x + kb * vg.x + kc * Dv.x;
            dS6[1] = ka * xg.y + kb * Dx.y + kb * vg.y + kc * Dv.y;
            dS6[2] = ka * xg.z + kb * Dx.z + kb * vg.z + kc * Dv.z;
        }

        // ================= (4) all stores =================
        a.dL_dmeans2D[3 * idx] = g0.x; a.dL_dmeans2D[3 * idx + 1] = g0.y; a.dL_dmeans2D[3 * idx + 2] = 0.0f;
        emit(a.dL_dopacity + idx, g1.y, acc & GSR_ACC_OPACITY);
        if (a.dL_dcolors) { a.dL_dcolors[3 * idx] = g1.z; a.dL_dcolors[3 * idx + 1] = g1.w; a.dL_dcolors[3 * idx + 2] = g2x; }
        if (a.dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
        }
        if (a.shs) {
            // dL_dsh record: 48 floats (element e = coef[e/3] * dRGB[e%3]) as 6 x STG.256, or,
            // accumulating, 12 x RED.ADD.F32x4
            float* dsh = a.dL_dsh + (size_t)idx * M * 3;
            const float dr[3] = {dRGB.x, dRGB.y, dRGB.z};
            if (wide) {
                if (acc & GSR_ACC_SH) {
#pragma unroll
                    for (int k = 0; k < 12; k++) {
                        float4 o;
                        o.x = coef[(4 * k) / 3] * dr[(4 * k) % 3];
                        o.y = coef[(4 * k + 1) / 3] * dr[(4 * k + 1) % 3];
                        o.z = coef[(4 * k + 2) / 3] * dr[(4 * k + 2) % 3];
                        o.w = coef[(4 * k + 3) / 3] * dr[(4 * k + 3) % 3];
                        if (4 * k < (deg + 1) * (deg + 1) * 3) atomicAdd(reinterpret_cast<float4*>(dsh) + k, o);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        float o[8];
#pragma unroll
                        for (int e8 = 0; e8 < 8; e8++) o[e8] = coef[(8 * k + e8) / 3] * dr[(8 * k + e8) % 3];
                        st256(dsh + 8 * k, o);
                    }
                }
            } else {
#pragma unroll
                for (int e1 = 0; e1 < 48; e1++)
                    if (e1 < M * 3) emit(dsh + e1, coef[e1 / 3] * dr[e1 % 3], acc & GSR_ACC_SH);
            }
        }
        if (a.dL_dscales) {
#pragma unroll
            for (int k = 0; k < 3; k++) emit(a.dL_dscales + 3 * idx + k, dscale[k], acc & GSR_ACC_SCALES);
        }
        if (a.dL_drots) {
            st_rec4(a.dL_drots, idx, make_float4(drot[0], drot[1], drot[2], drot[3]), (acc & GSR_ACC_ROTS) != 0);
        }
        if (a.dL_dtwist_S && a.deform_mode != GSR_DEFORM_NONE) {
            if (a.deform_mode == GSR_DEFORM_PER_GAUSSIAN) {
#pragma unroll
                for (int k = 0; k < 6; k++) emit(a.dL_dtwist_S + 6 * (size_t)idx + k, dS6[k], acc & GSR_ACC_TWIST);
                emit(a.dL_dtwist_theta + idx, dth, acc & GSR_ACC_TWIST);
            } else if (body_smem) {
                float* dstS = s_body + 7 * tix;
#pragma unroll
                for (int k = 0; k < 6; k++) atomicAdd(dstS + k, dS6[k]);
                atomicAdd(dstS + 6, dth);
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) atomicAdd(a.dL_dtwist_S + 6 * (size_t)tix + k, dS6[k]);
                atomicAdd(a.dL_dtwist_theta + tix, dth);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) emit(a.dL_dmeans3D + 3 * idx + k, gx[k], acc & GSR_ACC_MEANS3D);
    }
    if (body_smem) {
        __syncthreads();
        for (int k = threadIdx.x; k < a.num_bodies * 7; k += 256) {
            const float val = s_body[k];
            if (val != 0.0f) {
                const int b = k / 7, c = k % 7;
                if (c < 6) atomicAdd(a.dL_dtwist_S + 6 * (size_t)b + c, val);
                else atomicAdd(a.dL_dtwist_theta + b, val);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// View-batched per-Gaussian backward.  A training step renders several views of the SAME parameters and sums their
// gradients; run per view, the chain rule above re-reads every parameter record (330 B) and read-modify-writes the 264-byte
// running gradient for each view.  Here one thread takes a Gaussian through ALL views of the batch: the parameters are
// read once, only the view's 84 bytes (moment record, conic + opacity, clamp bits, radius) per view, and the gradient is
// updated once.  What makes the sum cheap: for fixed scale / rotation / twist the cov3D and SE3 backward are LINEAR in
// dL/dSigma and dL/d(deformed mean), so those are summed over the views and pushed through cov3D / SE3 once.
// Same formulas as preprocess_bwd_kernel; supported configuration: scales + rotations, SH colours (M = 16, 32-byte
// aligned), any deform mode.  dL_dmeans2D is written per view (zeros where the view culled the Gaussian).
// ---------------------------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(128, MINB) preprocess_bwd_batched_kernel(PreprocessBwdBatchArgs a, const BwdViewSlot* __restrict__ g_slots) {
    extern __shared__ float s_body[];
    // the views' camera constants and workspace pointers: shared memory (broadcast reads, no pointer chase through global)
    __shared__ __align__(16) BwdViewSlot slots[GSR_BATCH_MAX_VIEWS];
    const int idx = a.first + blockIdx.x * 128 + threadIdx.x;       // this launch covers Gaussians [first, P)
    const bool body_smem = (a.deform_mode == GSR_DEFORM_RIGID_BODIES) && a.dL_dtwist_S && (a.num_bodies * 7 * 4 <= 32768);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g_slots);
        uint32_t* dst = reinterpret_cast<uint32_t*>(slots);
        const int words = a.n_views * (int)(sizeof(BwdViewSlot) / 4);
        for (int k = threadIdx.x; k < words; k += 128) dst[k] = __ldg(src + k);
    }
    if (body_smem)
        for (int k = threadIdx.x; k < a.num_bodies * 7; k += 128) s_body[k] = 0.0f;
    __syncthreads();
    const int acc = a.acc;
    unsigned vis = 0;
    if (idx < a.P) {
        for (int j = 0; j < a.n_views; j++) vis |= (__ldg(slots[j].radii + idx) > 0 ? 1u : 0u) << j;
        for (int j = 0; j < a.n_views; j++) {
            if (!((vis >> j) & 1u)) {
                float* m2 = slots[j].dL_dmeans2D + 3 * (size_t)idx;
                m2[0] = 0.0f; m2[1] = 0.0f; m2[2] = 0.0f;
            }
        }
    }
    if (idx < a.P && vis == 0u) {
        // culled by every view: zeros to the outputs that are not running sums
        if (!(acc & GSR_ACC_OPACITY)) a.dL_dopacity[idx] = 0.0f;
        if (!(acc & GSR_ACC_SH)) {
            const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 6; k++) st256(a.dL_dsh + 48 * (size_t)idx + 8 * k, z8);
        }
        if (!(acc & GSR_ACC_SCALES)) { a.dL_dscales[3 * idx] = 0.0f; a.dL_dscales[3 * idx + 1] = 0.0f; a.dL_dscales[3 * idx + 2] = 0.0f; }
        if (!(acc & GSR_ACC_ROTS)) st_rec4(a.dL_drots, idx, make_float4(0.f, 0.f, 0.f, 0.f), false);
        if (a.dL_dtwist_S && a.deform_mode == GSR_DEFORM_PER_GAUSSIAN && !(acc & GSR_ACC_TWIST)) {
#pragma unroll
            for (int k = 0; k < 6; k++) a.dL_dtwist_S[6 * (size_t)idx + k] = 0.0f;
            a.dL_dtwist_theta[idx] = 0.0f;
        }
        if (!(acc & GSR_ACC_MEANS3D)) { a.dL_dmeans3D[3 * idx] = 0.0f; a.dL_dmeans3D[3 * idx + 1] = 0.0f; a.dL_dmeans3D[3 * idx + 2] = 0.0f; }
    }
    if (vis != 0u) {
        // ---- the Gaussian's own records, once ----
        const float* mp = (a.means_deformed ? a.means_deformed : a.means) + 3 * (size_t)idx;
        const float3 mean = make_float3(mp[0], mp[1], mp[2]);
        const float3 s = make_float3(a.scales[3 * idx], a.scales[3 * idx + 1], a.scales[3 * idx + 2]);
        const float4 q = ld_rec4(a.rotations, idx);
        float shv[48];
        {
            const float* shp = a.shs + 48 * (size_t)idx;
#pragma unroll
            for (int k = 0; k < 6; k++) ld256_nc(shp + 8 * k, shv + 8 * k);
        }
        const int tix = (a.deform_mode == GSR_DEFORM_RIGID_BODIES) ? a.body_id[idx] : idx;
        float3 w = make_float3(0, 0, 0), tv = w, x0 = w;
        float th = 0.0f;
        if (a.deform_mode != GSR_DEFORM_NONE) {
            const float* S = a.twist_S + 6 * (size_t)tix;
            w = make_float3(S[0], S[1], S[2]); tv = make_float3(S[3], S[4], S[5]);
            th = a.twist_theta[tix];
            x0 = make_float3(a.means[3 * idx], a.means[3 * idx + 1], a.means[3 * idx + 2]);
        }
        float cov6[6];
        cov3d_exact(s, a.scale_modifier, q, cov6);
        const float V00 = cov6[0], V01 = cov6[1], V02 = cov6[2], V11 = cov6[3], V12 = cov6[4], V22 = cov6[5];

        // ---- sums over the views ----
        float gm[3] = {0, 0, 0};                 // dL/d(deformed mean)
        float dcov[6] = {0, 0, 0, 0, 0, 0};      // dL/dSigma
        float dsh[48];
#pragma unroll
        for (int k = 0; k < 48; k++) dsh[k] = 0.0f;
        float dop = 0.0f;
        // the view's records (84 B) are requested one view ahead of the math that consumes them
        float4 n_g0, n_g1, n_sp0 = make_float4(0, 0, 0, 0);
        float2 n_sp1 = make_float2(0, 0);
        float n_g2x;
        uint8_t n_cl;
        auto fetch = [&](int j) {
            const BwdViewSlot* sl = slots + j;
            const float4* gr = sl->grad_recs + 3 * (size_t)idx;
            n_g0 = gr[0]; n_g1 = gr[1];
            n_g2x = reinterpret_cast<const float*>(gr + 2)[0];
            n_cl = sl->clamped[idx];
            if (a.grad_moments) {
                const float4* sp = sl->recs + 3 * (size_t)idx;
                n_sp0 = sp[0];
                n_sp1 = *reinterpret_cast<const float2*>(sp + 1);
            }
        };
        fetch(__ffs(vis) - 1);
        for (unsigned rest = vis; rest; rest &= rest - 1) {
            const int j = __ffs(rest) - 1;
            const BwdViewSlot* sl = slots + j;
            const GsrView& v = sl->v;
            float4 g0 = n_g0, g1 = n_g1;
            const float g2x = n_g2x;
            const uint8_t cl = n_cl;
            const float4 sp0 = n_sp0;
            const float2 sp1 = n_sp1;
            if (rest & (rest - 1)) fetch(__ffs(rest & (rest - 1)) - 1);
            if (a.grad_moments) {
                const float M0 = g0.x, Mx = g0.y, My = g0.z, Mxx = g0.w, Mxy = g1.x, Myy = g1.y;
                const float cx = sp0.z, cy = sp0.w, cz = sp1.x, op = sp1.y;
                const float hop = -0.5f * op;
                g0.x = -op * (0.5f * v.W) * (cx * Mx + cy * My);
                g0.y = -op * (0.5f * v.H) * (cz * My + cy * Mx);
                g0.z = hop * Mxx; g0.w = hop * Mxy; g1.x = hop * Myy;
                g1.y = M0;
            }
            {
                float* m2 = sl->dL_dmeans2D + 3 * (size_t)idx;
                m2[0] = g0.x; m2[1] = g0.y; m2[2] = 0.0f;
            }
            dop += g1.y;
            // ---- computeCov2DCUDA (backward.cu:144-274) ----
            // (gradients carry a 1e-4 tolerance: reciprocals by MUFU.RCP + multiply here, where the per-view kernel and the
            // forward keep IEEE divisions for bit-exact geometry - this loop is issue-bound, ~1000 instructions per pair)
            float3 t = make_float3(xform_row(v.view, 0, mean), xform_row(v.view, 1, mean), xform_row(v.view, 2, mean));
            const float limx = 1.3f * v.tan_fovx, limy = 1.3f * v.tan_fovy;
            const float tz = __frcp_rn(t.z);
            const float txtz = t.x * tz, tytz = t.y * tz;
            const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
            const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
            t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
            t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
            const float tz2 = tz * tz, tz3 = tz2 * tz;
            EwaT e;
            {
                const float J00 = v.focal_x * tz, J02 = -v.focal_x * t.x * tz2, J11 = v.focal_y * tz, J12 = -v.focal_y * t.y * tz2;
                const float* V = v.view;
                e.T00 = V[2] * J02 + V[0] * J00; e.T01 = V[6] * J02 + V[4] * J00; e.T02 = V[10] * J02 + V[8] * J00;
                e.T10 = V[2] * J12 + J11 * V[1]; e.T11 = V[6] * J12 + J11 * V[5]; e.T12 = V[10] * J12 + J11 * V[9];
            }
            const float3 cov = cov2d_exact(e, cov6);
            const float ca = cov.x, cb = cov.y, cc = cov.z;
            const float denom = ca * cc - cb * cb;
            const float denom2inv = __frcp_rn((denom * denom) + 0.0000001f);
            const float dcx = g0.z, dcy = g0.w, dcz = g1.x;
            float dL_da = 0, dL_db = 0, dL_dc = 0;
            const float T00 = e.T00, T01 = e.T01, T02 = e.T02, T10 = e.T10, T11 = e.T11, T12 = e.T12;
            if (denom2inv != 0) {
                dL_da = denom2inv * (-cc * cc * dcx + 2 * cb * cc * dcy + (denom - ca * cc) * dcz);
                dL_dc = denom2inv * (-ca * ca * dcz + 2 * ca * cb * dcy + (denom - ca * cc) * dcx);
                dL_db = denom2inv * 2 * (cb * cc * dcx - (denom + 2 * cb * cb) * dcy + ca * cb * dcz);
                dcov[0] += (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
                dcov[3] += (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
                dcov[5] += (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
                dcov[1] += 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
                dcov[2] += 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
                dcov[4] += 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
            }
            const float dL_dT00 = 2 * (T00 * V00 + T01 * V01 + T02 * V02) * dL_da + (T10 * V00 + T11 * V01 + T12 * V02) * dL_db;
            const float dL_dT01 = 2 * (T00 * V01 + T01 * V11 + T02 * V12) * dL_da + (T10 * V01 + T11 * V11 + T12 * V12) * dL_db;
            const float dL_dT02 = 2 * (T00 * V02 + T01 * V12 + T02 * V22) * dL_da + (T10 * V02 + T11 * V12 + T12 * V22) * dL_db;
            const float dL_dT10 = 2 * (T10 * V00 + T11 * V01 + T12 * V02) * dL_dc + (T00 * V00 + T01 * V01 + T02 * V02) * dL_db;
            const float dL_dT11 = 2 * (T10 * V01 + T11 * V11 + T12 * V12) * dL_dc + (T00 * V01 + T01 * V11 + T02 * V12) * dL_db;
            const float dL_dT12 = 2 * (T10 * V02 + T11 * V12 + T12 * V22) * dL_dc + (T00 * V02 + T01 * V12 + T02 * V22) * dL_db;
            const float* Vm = v.view;
            const float dL_dJ00 = Vm[0] * dL_dT00 + Vm[4] * dL_dT01 + Vm[8] * dL_dT02;
            const float dL_dJ02 = Vm[2] * dL_dT00 + Vm[6] * dL_dT01 + Vm[10] * dL_dT02;
            const float dL_dJ11 = Vm[1] * dL_dT10 + Vm[5] * dL_dT11 + Vm[9] * dL_dT12;
            const float dL_dJ12 = Vm[2] * dL_dT10 + Vm[6] * dL_dT11 + Vm[10] * dL_dT12;
            const float h_x = v.focal_x, h_y = v.focal_y;
            const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
            const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
            const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                                 (2 * h_y * t.y) * tz3 * dL_dJ12;
            float gv[3];
            gv[0] = Vm[0] * dL_dtx + Vm[1] * dL_dty + Vm[2] * dL_dtz;
            gv[1] = Vm[4] * dL_dtx + Vm[5] * dL_dty + Vm[6] * dL_dtz;
            gv[2] = Vm[8] * dL_dtx + Vm[9] * dL_dty + Vm[10] * dL_dtz;
            // ---- projection part (backward.cu:370-387) ----
            const float* proj = v.proj;
            const float m_hom_w = proj[3] * mean.x + proj[7] * mean.y + proj[11] * mean.z + proj[15];
            const float m_w = __frcp_rn(m_hom_w + 0.0000001f);
            const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
            const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
            gv[0] += (proj[0] * m_w - proj[3] * mul1) * g0.x + (proj[1] * m_w - proj[3] * mul2) * g0.y;
            gv[1] += (proj[4] * m_w - proj[7] * mul1) * g0.x + (proj[5] * m_w - proj[7] * mul2) * g0.y;
            gv[2] += (proj[8] * m_w - proj[11] * mul1) * g0.x + (proj[9] * m_w - proj[11] * mul2) * g0.y;
            // ---- SH backward (backward.cu:20-139) ----
            {
                const int deg = v.sh_degree;
                const V3 dRGB = {(cl & 1) ? 0.0f : g1.z, (cl & 2) ? 0.0f : g1.w, (cl & 4) ? 0.0f : g2x};
                // dL/ddir = sum_k (d coef_k / d dir) (sh_k . dRGB): the dot products first (one scalar per coefficient)
                float pk[16];
#pragma unroll
                for (int k = 0; k < 16; k++) pk[k] = shv[3 * k] * dRGB.x + shv[3 * k + 1] * dRGB.y + shv[3 * k + 2] * dRGB.z;
                const V3 dir_orig = {mean.x - v.campos[0], mean.y - v.campos[1], mean.z - v.campos[2]};
                const float sum2 = dot(dir_orig, dir_orig);
                const float rlen = rsqrtf(sum2);
                const float x = dir_orig.x * rlen, y = dir_orig.y * rlen, z = dir_orig.z * rlen;
                float ddx = 0.0f, ddy = 0.0f, ddz = 0.0f;
                float coef[16];
#pragma unroll
                for (int k = 0; k < 16; k++) coef[k] = 0.0f;
                coef[0] = bSH_C0;
                if (deg > 0) {
                    coef[1] = -bSH_C1 * y; coef[2] = bSH_C1 * z; coef[3] = -bSH_C1 * x;
                    ddx = -bSH_C1 * pk[3]; ddy = -bSH_C1 * pk[1]; ddz = bSH_C1 * pk[2];
                    if (deg > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        coef[4] = bSH_C2[0] * xy; coef[5] = bSH_C2[1] * yz; coef[6] = bSH_C2[2] * (2.f * zz - xx - yy);
                        coef[7] = bSH_C2[3] * xz; coef[8] = bSH_C2[4] * (xx - yy);
                        ddx += (bSH_C2[0] * y) * pk[4] + (bSH_C2[2] * 2.f * -x) * pk[6] + (bSH_C2[3] * z) * pk[7] + (bSH_C2[4] * 2.f * x) * pk[8];
                        ddy += (bSH_C2[0] * x) * pk[4] + (bSH_C2[1] * z) * pk[5] + (bSH_C2[2] * 2.f * -y) * pk[6] + (bSH_C2[4] * 2.f * -y) * pk[8];
                        ddz += (bSH_C2[1] * y) * pk[5] + (bSH_C2[2] * 2.f * 2.f * z) * pk[6] + (bSH_C2[3] * x) * pk[7];
                        if (deg > 2) {
                            coef[9] = bSH_C3[0] * y * (3.f * xx - yy); coef[10] = bSH_C3[1] * xy * z;
                            coef[11] = bSH_C3[2] * y * (4.f * zz - xx - yy);
                            coef[12] = bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                            coef[13] = bSH_C3[4] * x * (4.f * zz - xx - yy); coef[14] = bSH_C3[5] * z * (xx - yy);
                            coef[15] = bSH_C3[6] * x * (xx - 3.f * yy);
                            ddx += (bSH_C3[0] * 3.f * 2.f * xy) * pk[9] + (bSH_C3[1] * yz) * pk[10] + (bSH_C3[2] * -2.f * xy) * pk[11] +
                                   (bSH_C3[3] * -3.f * 2.f * xz) * pk[12] + (bSH_C3[4] * (-3.f * xx + 4.f * zz - yy)) * pk[13] +
                                   (bSH_C3[5] * 2.f * xz) * pk[14] + (bSH_C3[6] * 3.f * (xx - yy)) * pk[15];
                            ddy += (bSH_C3[0] * 3.f * (xx - yy)) * pk[9] + (bSH_C3[1] * xz) * pk[10] +
                                   (bSH_C3[2] * (-3.f * yy + 4.f * zz - xx)) * pk[11] + (bSH_C3[3] * -3.f * 2.f * yz) * pk[12] +
                                   (bSH_C3[4] * -2.f * xy) * pk[13] + (bSH_C3[5] * -2.f * yz) * pk[14] + (bSH_C3[6] * -3.f * 2.f * xy) * pk[15];
                            ddz += (bSH_C3[1] * xy) * pk[10] + (bSH_C3[2] * 4.f * 2.f * yz) * pk[11] +
                                   (bSH_C3[3] * 3.f * (2.f * zz - xx - yy)) * pk[12] + (bSH_C3[4] * 4.f * 2.f * xz) * pk[13] +
                                   (bSH_C3[5] * (xx - yy)) * pk[14];
                        }
                    }
                }
                // dnormvdv (auxiliary.h:107-117) with 1 / |dir|^3 from the reciprocal square root
                {
                    const float inv32 = rlen * rlen * rlen;
                    const V3 d = dir_orig;
                    gv[0] += ((sum2 - d.x * d.x) * ddx - d.y * d.x * ddy - d.z * d.x * ddz) * inv32;
                    gv[1] += (-d.x * d.y * ddx + (sum2 - d.y * d.y) * ddy - d.z * d.y * ddz) * inv32;
                    gv[2] += (-d.x * d.z * ddx - d.y * d.z * ddy + (sum2 - d.z * d.z) * ddz) * inv32;
                }
                const float dr[3] = {dRGB.x, dRGB.y, dRGB.z};
#pragma unroll
                for (int k = 0; k < 48; k++) dsh[k] = fmaf(coef[k / 3], dr[k % 3], dsh[k]);
            }
            gm[0] += gv[0]; gm[1] += gv[1]; gm[2] += gv[2];
        }

        // ---- cov3D backward (backward.cu:278-341), once, on the summed dL/dSigma ----
        float dscale[3], drot[4];
        {
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            const float R[3][3] = {
                {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float sv[3] = {a.scale_modifier * s.x, a.scale_modifier * s.y, a.scale_modifier * s.z};
            float Mm[3][3];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int rr = 0; rr < 3; rr++) Mm[c][rr] = sv[rr] * R[c][rr];
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float dM[3][3];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int rr = 0; rr < 3; rr++)
                    dM[c][rr] = 2.0f * (Mm[0][rr] * dS[c][0] + Mm[1][rr] * dS[c][1] + Mm[2][rr] * dS[c][2]);
            float dMt[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < 3; j++) dMt[k][j] = dM[j][k];
#pragma unroll
            for (int k = 0; k < 3; k++)
                dscale[k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int j = 0; j < 3; j++) dMt[k][j] *= sv[k];
            drot[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            drot[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            drot[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            drot[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
        // ---- SE3 backward (closed form), once, on the summed dL/d(deformed mean) ----
        float gx[3] = {gm[0], gm[1], gm[2]};
        float dS6[6] = {0, 0, 0, 0, 0, 0};
        float dth = 0.0f;
        if (a.deform_mode != GSR_DEFORM_NONE) {
            const float3 g = make_float3(gm[0], gm[1], gm[2]);
            float sn, cs;
            sincosf(th, &sn, &cs);
            const float ka = sn, kb = 1.0f - cs, kc = th - sn;
            const float3 wg = cross3(w, g), wwg = skew2(w, g);
            gx[0] = g.x - ka * wg.x + kb * wwg.x;
            gx[1] = g.y - ka * wg.y + kb * wwg.y;
            gx[2] = g.z - ka * wg.z + kb * wwg.z;
            dS6[3] = th * g.x - kb * wg.x + kc * wwg.x;
            dS6[4] = th * g.y - kb * wg.y + kc * wwg.y;
            dS6[5] = th * g.z - kb * wg.z + kc * wwg.z;
            const float3 wx = cross3(w, x0), wwx = skew2(w, x0), wv = cross3(w, tv), wwv = skew2(w, tv);
            dth = g.x * (cs * wx.x + sn * wwx.x + tv.x + sn * wv.x + kb * wwv.x) +
                  g.y * (cs * wx.y + sn * wwx.y + tv.y + sn * wv.y + kb * wwv.y) +
                  g.z * (cs * wx.z + sn * wwx.z + tv.z + sn * wv.z + kb * wwv.z);
            const float3 xg = cross3(x0, g), vg = cross3(tv, g);
            const float wdx = dot3(w, x0), wdv = dot3(w, tv), wdg = dot3(w, g), gdx = dot3(g, x0), gdv = dot3(g, tv);
            const float3 Dx = make_float3(g.x * wdx + x0.x * wdg - 2.f * w.x * gdx, g.y * wdx + x0.y * wdg - 2.f * w.y * gdx,
                                          g.z * wdx + x0.z * wdg - 2.f * w.z * gdx);
            const float3 Dv = make_float3(g.x * wdv + tv.x * wdg - 2.f * w.x * gdv, g.y * wdv + tv.y * wdg - 2.f * w.y * gdv,
                                          g.z * wdv + tv.z * wdg - 2.f * w.z * gdv);
            dS6[0] = ka * xg.x + kb * Dx.x + kb * vg.x + kc * Dv.x;
            dS6[1] = ka * xg.y + kb * Dx.y + kb * vg.y + kc * Dv.y;
            dS6[2] = ka * xg.z + kb * Dx.z + kb * vg.z + kc * Dv.z;
        }
        // ---- stores: running sums by fire-and-forget RED (measured: a plain load-add-store per output is 12 % slower -
        // the extra DRAM round trip at the end of every thread outweighs the atomics) ----
        auto upd = [&](float* p, float val, bool accumulate) { emit(p, val, accumulate); };
        {
            float* dst = a.dL_dsh + 48 * (size_t)idx;
            if (acc & GSR_ACC_SH) {
#pragma unroll
                for (int k = 0; k < 12; k++)
                    atomicAdd(reinterpret_cast<float4*>(dst) + k, make_float4(dsh[4 * k], dsh[4 * k + 1], dsh[4 * k + 2], dsh[4 * k + 3]));
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) st256(dst + 8 * k, dsh + 8 * k);
            }
        }
        upd(a.dL_dopacity + idx, dop, acc & GSR_ACC_OPACITY);
#pragma unroll
        for (int k = 0; k < 3; k++) upd(a.dL_dscales + 3 * idx + k, dscale[k], acc & GSR_ACC_SCALES);
        st_rec4(a.dL_drots, idx, make_float4(drot[0], drot[1], drot[2], drot[3]), (acc & GSR_ACC_ROTS) != 0);
        if (a.dL_dtwist_S && a.deform_mode != GSR_DEFORM_NONE) {
            if (a.deform_mode == GSR_DEFORM_PER_GAUSSIAN) {
#pragma unroll
                for (int k = 0; k < 6; k++) upd(a.dL_dtwist_S + 6 * (size_t)idx + k, dS6[k], acc & GSR_ACC_TWIST);
                upd(a.dL_dtwist_theta + idx, dth, acc & GSR_ACC_TWIST);
            } else if (body_smem) {
                float* dstS = s_body + 7 * tix;
#pragma unroll
                for (int k = 0; k < 6; k++) atomicAdd(dstS + k, dS6[k]);
                atomicAdd(dstS + 6, dth);
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) atomicAdd(a.dL_dtwist_S + 6 * (size_t)tix + k, dS6[k]);
                atomicAdd(a.dL_dtwist_theta + tix, dth);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) upd(a.dL_dmeans3D + 3 * idx + k, gx[k], acc & GSR_ACC_MEANS3D);
    }
    if (body_smem) {
        __syncthreads();
        for (int k = threadIdx.x; k < a.num_bodies * 7; k += 128) {
            const float val = s_body[k];
            if (val != 0.0f) {
                const int b = k / 7, c = k % 7;
                if (c < 6) atomicAdd(a.dL_dtwist_S + 6 * (size_t)b + c, val);
                else atomicAdd(a.dL_dtwist_theta + b, val);
            }
        }
    }
}

// ---- standalone exp_se3: S[N,6], theta[N] -> T[N,4,4] (rigid_body.exp_se3) ----
__global__ void __launch_bounds__(256) se3_matrices_kernel(int N, const float* __restrict__ S,
                                                           const float* __restrict__ theta, float* __restrict__ T44) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float3 w = make_float3(S[6 * i], S[6 * i + 1], S[6 * i + 2]);
    const float3 tv = make_float3(S[6 * i + 3], S[6 * i + 4], S[6 * i + 5]);
    const float th = theta[i];
    float sn, cs;
    sincosf(th, &sn, &cs);
    const float ka = sn, kb = 1.0f - cs, kc = th - sn;
    const float ww = dot3(w, w);
    // W = skew(w); W2 = w w^T - (w.w) I
    const float W[3][3] = {{0, -w.z, w.y}, {w.z, 0, -w.x}, {-w.y, w.x, 0}};
    const float wv[3] = {w.x, w.y, w.z};
    const float vv[3] = {tv.x, tv.y, tv.z};
    float4* out = reinterpret_cast<float4*>(T44 + 16 * (size_t)i);
    float row[3][4];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float p = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float W2 = wv[r] * wv[c] - (r == c ? ww : 0.0f);
            row[r][c] = (r == c ? 1.0f : 0.0f) + ka * W[r][c] + kb * W2;
            p += ((r == c ? th : 0.0f) + kb * W[r][c] + kc * W2) * vv[c];
        }
        row[r][3] = p;
    }
#pragma unroll
    for (int r = 0; r < 3; r++) out[r] = make_float4(row[r][0], row[r][1], row[r][2], row[r][3]);
    out[3] = make_float4(0.f, 0.f, 0.f, 1.f);
}

// Backward of exp_se3 for an arbitrary upstream gradient dT[N,4,4] (rows 0..2 used).
__global__ void __launch_bounds__(256) se3_matrices_bwd_kernel(int N, const float* __restrict__ S,
                                                               const float* __restrict__ theta,
                                                               const float* __restrict__ dT44, float* __restrict__ dS,
                                                               float* __restrict__ dtheta) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float wv[3] = {S[6 * i], S[6 * i + 1], S[6 * i + 2]};
    const float vv[3] = {S[6 * i + 3], S[6 * i + 4], S[6 * i + 5]};
    const float th = theta[i];
    float sn, cs;
    sincosf(th, &sn, &cs);
    const float ka = sn, kb = 1.0f - cs, kc = th - sn;
    const float ww = wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2];
    const float W[3][3] = {{0, -wv[2], wv[1]}, {wv[2], 0, -wv[0]}, {-wv[1], wv[0], 0}};
    float GR[3][3], gp[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) GR[r][c] = dT44[16 * (size_t)i + 4 * r + c];
        gp[r] = dT44[16 * (size_t)i + 4 * r + 3];
    }
    // A = th I + kb W + kc W2 ; p = A v ; R = I + ka W + kb W2
    // dL/dW(r,c)  = ka GR + kb gp v^T ;  dL/dW2(r,c) = kb GR + kc gp v^T
    float dth = 0.0f, dv[3] = {0, 0, 0}, dw[3] = {0, 0, 0};
    float dW[3][3], dW2[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float W2 = wv[r] * wv[c] - (r == c ? ww : 0.0f);
            const float A = (r == c ? th : 0.0f) + kb * W[r][c] + kc * W2;
            dv[c] += A * gp[r];
            const float gpv = gp[r] * vv[c];
            dW[r][c] = ka * GR[r][c] + kb * gpv;
            dW2[r][c] = kb * GR[r][c] + kc * gpv;
            // d/dth: R' = cos W + sin W2 ; A' = I + sin W + (1-cos) W2
            dth += GR[r][c] * (cs * W[r][c] + sn * W2) + gpv * ((r == c ? 1.0f : 0.0f) + sn * W[r][c] + kb * W2);
        }
    // W = skew(w): dL/dw_x = dW[2][1] - dW[1][2], etc.
    dw[0] = dW[2][1] - dW[1][2];
    dw[1] = dW[0][2] - dW[2][0];
    dw[2] = dW[1][0] - dW[0][1];
    // W2 = w w^T - (w.w) I : dL/dw_k = sum_c (dW2[k][c] + dW2[c][k]) w_c - 2 w_k tr(dW2)
    const float tr = dW2[0][0] + dW2[1][1] + dW2[2][2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) acc += (dW2[k][c] + dW2[c][k]) * wv[c];
        dw[k] += acc - 2.0f * wv[k] * tr;
    }
    dS[6 * i] = dw[0]; dS[6 * i + 1] = dw[1]; dS[6 * i + 2] = dw[2];
    dS[6 * i + 3] = dv[0]; dS[6 * i + 4] = dv[1]; dS[6 * i + 5] = dv[2];
    dtheta[i] = dth;
}

}  // namespace

int gsr_launch_preprocess_bwd_batched(const PreprocessBwdBatchArgs& a, const BwdViewSlot* d_slots, cudaStream_t stream) {
    if (a.P <= a.first || a.n_views <= 0) return 0;
    size_t smem = 0;
    if (a.deform_mode == GSR_DEFORM_RIGID_BODIES && a.dL_dtwist_S && a.num_bodies * 7 * 4 <= 32768)
        smem = (size_t)a.num_bodies * 7 * 4;
    { GsrProfScope prof_("preprocess_bwd_batched", stream);
    // measured at C2 (8 views): 234 registers / 2 CTAs per SM 0.376 ms, 168 / 3 (spills) 0.447 ms, 128 / 4 0.530 ms: issue-bound
    static const int minb = getenv("GSR_PRE_BWDB_MINB") ? atoi(getenv("GSR_PRE_BWDB_MINB")) : 2;
    const int blocks = gsr_div_up(a.P - a.first, 128);
    if (minb >= 4) preprocess_bwd_batched_kernel<4><<<blocks, 128, smem, stream>>>(a, d_slots);
    else if (minb == 2) preprocess_bwd_batched_kernel<2><<<blocks, 128, smem, stream>>>(a, d_slots);
    else preprocess_bwd_batched_kernel<3><<<blocks, 128, smem, stream>>>(a, d_slots); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_preprocess_bwd(const PreprocessBwdArgs& a, const GsrView& v, cudaStream_t stream) {
    if (a.P <= 0) return 0;
    size_t smem = 0;
    if (a.deform_mode == GSR_DEFORM_RIGID_BODIES && a.dL_dtwist_S && a.num_bodies * 7 * 4 <= 32768)
        smem = (size_t)a.num_bodies * 7 * 4;
    { GsrProfScope prof_("preprocess_bwd", stream);
    static const int minb = getenv("GSR_PRE_BWD_MINB") ? atoi(getenv("GSR_PRE_BWD_MINB")) : 2;
    if (minb >= 3) preprocess_bwd_kernel<3><<<gsr_div_up(a.P, 256), 256, smem, stream>>>(a, v);
    else preprocess_bwd_kernel<2><<<gsr_div_up(a.P, 256), 256, smem, stream>>>(a, v); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_se3_matrices(int N, const float* S, const float* theta, float* T44, cudaStream_t stream) {
    if (N <= 0) return 0;
    { GsrProfScope prof_("se3_matrices", stream);
    se3_matrices_kernel<<<gsr_div_up(N, 256), 256, 0, stream>>>(N, S, theta, T44); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_se3_matrices_bwd(int N, const float* S, const float* theta, const float* dT44, float* dS,
                                float* dtheta, cudaStream_t stream) {
    if (N <= 0) return 0;
    { GsrProfScope prof_("se3_matrices_bwd", stream);
    se3_matrices_bwd_kernel<<<gsr_div_up(N, 256), 256, 0, stream>>>(N, S, theta, dT44, dS, dtheta); }
    GSR_CHECK_LAUNCH();
    return 0;
}
