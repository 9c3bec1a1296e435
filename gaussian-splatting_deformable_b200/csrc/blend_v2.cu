// Tile blending, second generation: warp-autonomous regions + packed FP32x2 pixel pairs.
//
// Replaces cuda_rasterizer/forward.cu:261-374 (renderCUDA fwd) and backward.cu:399-557
// (renderCUDA bwd) like blend.cu, with the same bit-identical per-pixel decisions, but:
//
//   * a WARP owns an 8 x (8 NP) pixel region of the 16x16 tile and walks the tile's list on
//     its own: it gathers 32 list entries at a time (one 48-byte record per lane, next batch in
//     flight while the current one is blended), culls them against ITS region (the closed-form
//     bound of blend.cu on a rectangle a quarter or half the size, so far fewer survivors),
//     ballot-compacts the survivors into a warp-private shared ring, and stops as soon as its
//     own pixels are done.  No CTA barrier anywhere.
//   * a lane owns NP pairs of vertically adjacent pixels.  The two pixels of a pair share dx,
//     and everything that differs is computed on both at once with sm_100's packed
//     FFMA2/FMUL2/FADD2 (f32x2.cuh): same IEEE roundings as the reference's scalar
//     sequence, half the issue slots.  expf is CUDA's own algorithm restated on packed
//     values.  These kernels are issue-bound (profiles/), so issue slots are the currency.
//   * the loop bodies are STRAIGHT-LINE code: after the region cull 92 % of the iterations blend, so a
//     pixel that does not blend simply runs with alpha = 0 (and a finished pixel is parked far away so
//     that the ordinary power test rejects everything) - no skips, no phi copies, 31 of 32 threads active.
//   * backward: per-Gaussian sums are accumulated as six moments (common.cuh) plus the colour
//     gradient, parked for four consecutive Gaussians in a warp-private shared buffer and reduced
//     together (8 x LDS.128 + FADD2 per lane), then leave the SM as 9 native RED.ADD.F32 per
//     (region, Gaussian).  Template switches keep the measured alternatives selectable.
#include "f32x2.cuh"
#include "geom_exact.cuh"
#include "kernels.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr unsigned FULL = 0xffffffffu;

// see blend.cu: true unless NO pixel centre in [x0,x1]x[y0,y1] can have power >= cut
__device__ __forceinline__ bool rect_may_contribute(float mx, float my, float a, float b, float c, float cut,
                                                    float x0, float y0, float x1, float y1) {
    const float dxl = x0 - mx, dxh = x1 - mx, dyl = y0 - my, dyh = y1 - my;
    const bool in_x = (dxl <= 0.0f) && (dxh >= 0.0f);
    const bool in_y = (dyl <= 0.0f) && (dyh >= 0.0f);
    if (in_x && in_y) return true;
    if (!(a > 0.0f && c > 0.0f && a * c > b * b)) return true;   // not PD (or NaN): never cull
    float qmin = 3.0e38f;
    if (!in_x) {
        const float dxe = dxl > 0.0f ? dxl : dxh;
        const float dys = fminf(fmaxf(-b * dxe / c, dyl), dyh);
        qmin = fminf(qmin, a * dxe * dxe + 2.0f * b * dxe * dys + c * dys * dys);
    }
    if (!in_y) {
        const float dye = dyl > 0.0f ? dyl : dyh;
        const float dxs = fminf(fmaxf(-b * dye / a, dxl), dxh);
        qmin = fminf(qmin, a * dxs * dxs + 2.0f * b * dxs * dye + c * dye * dye);
    }
    const float dxm = fmaxf(fabsf(dxl), fabsf(dxh)), dym = fmaxf(fabsf(dyl), fabsf(dyh));
    const float mag = 0.5f * (a * dxm * dxm + c * dym * dym) + fabsf(b) * dxm * dym;
    const float margin = 1e-2f + 4e-6f * mag;
    return !(-0.5f * qmin < cut - margin);
}

// The reference's exponent (forward.cu:335 as compiled, geom_exact.cuh:blend_power_exact)
//   power = fma(fma(dx, dx*cx, dy*(dy*cz)), -0.5, -(dy*(dx*cy)))
// for two pixels of one column: dx, s1 = dx*cx and s2 = dx*(-cy) are shared scalars.
__device__ __forceinline__ f32x2 power2_exact(float dx, float s1, float s2, f32x2 dy, float cz) {
    const f32x2 v = mul2(dy, mul2(dy, pk1(cz)));
    const f32x2 w = fma2(pk1(dx), pk1(s1), v);
    return fma2(w, pk1(-0.5f), mul2(dy, pk1(s2)));
}

// ---------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------
// WPC = warps per CTA: 4 / NP (a CTA = a tile) or 1 (a CTA = one 8x8 region, NP == 1: a finished region gives its
// registers back at once instead of holding them until the tile's slowest region is done).
template <int NP, int MINB, bool TMUL, bool STRAIGHT, int WPC = 4 / NP>
__global__ void __launch_bounds__(32 * WPC, MINB) blend_fwd_v2_kernel(BlendFwdArgs a) {
    constexpr int NW = 4 / NP;                       // warps (regions) per tile
    __shared__ float4 s_q0[WPC][32];                 // x, y, conic.x, -conic.y
    __shared__ float4 s_q1[WPC][32];                 // conic.z, opacity, cut, (list position + 1) as bits
    __shared__ float4 s_q2[WPC][32];                 // r, g, b, -
    const int lane = threadIdx.x & 31;
    // WPC == 1: 1-D grid, block = 4 tile + region, so the four regions of a tile are dispatched together and share L2 / L1
    const int warp = (WPC == 1) ? (int)(blockIdx.x & 3) : (int)(threadIdx.x >> 5);   // region of the tile
    const int sw = (WPC == 1) ? 0 : warp;
    const int tile1 = (int)(blockIdx.x >> 2);
    const int tbx = (WPC == 1) ? tile1 % a.grid_x : (int)blockIdx.x, tby = (WPC == 1) ? tile1 / a.grid_x : (int)blockIdx.y;
    const int tile = tby * a.grid_x + tbx;
    const int X0 = tbx * GSR_TILE + ((NP == 1) ? ((warp & 1) << 3) : (warp << 3));
    const int Y0 = tby * GSR_TILE + ((NP == 1) ? ((warp >> 1) << 3) : 0);
    if (X0 >= a.W || Y0 >= a.H) return;             // region outside the image: warps are independent
    const int px = X0 + (lane & 7);
    const float pxf = (float)px;
    const float rx0 = (float)X0, ry0 = (float)Y0;
    const float rx1 = fminf(rx0 + 7.0f, (float)(a.W - 1)), ry1 = fminf(ry0 + (float)(8 * NP - 1), (float)(a.H - 1));

    // A pixel that is done (T (1 - alpha) < 1e-4, forward.cu:345-350) - or lies outside the image - is
    // parked FAR_Y pixels away: -power = d^T Q d / 2 >= FAR_Y^2 / (2 sigma_max^2) then exceeds every cut
    // (|cut| <= 80) for any Gaussian with sigma_max < 7e7 pixels, and FAR_Y^2 lambda_max stays finite up to
    // lambda_max = 3e20 (the 0.3-pixel dilation bounds it by 3.4), so the pixel rejects every later Gaussian
    // in the ordinary power test and needs no flag in the loop.
    constexpr float FAR_Y = 1.0e9f;
    int pyi[NP];
    f32x2 npy[NP], T[NP], C0[NP], C1[NP], C2[NP];
    uint32_t lastA[NP], lastB[NP];
    bool inA[NP], inB[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) {
        pyi[q] = Y0 + 8 * q + 2 * (lane >> 3);
        inA[q] = px < a.W && pyi[q] < a.H;
        inB[q] = px < a.W && pyi[q] + 1 < a.H;
        npy[q] = pk(inA[q] ? -(float)pyi[q] : -FAR_Y, inB[q] ? -(float)(pyi[q] + 1) : -FAR_Y);
        T[q] = pk1(1.0f); C0[q] = C1[q] = C2[q] = pk1(0.0f);
        lastA[q] = lastB[q] = 0;
    }
    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);
    float4* const q0s = s_q0[sw]; float4* const q1s = s_q1[sw]; float4* const q2s = s_q2[sw];
    // region masks: word (batch, region) of this tile at NW * (range.x / 32 + tile + batch) + region
    uint32_t* const mask_out = a.region_masks ? a.region_masks + (size_t)NW * ((size_t)(range.x >> 5) + (size_t)tile) + warp : nullptr;

    float4 n0, n1, n2;
    bool nvalid = false;
    auto fetch = [&](int base) {
        nvalid = base + lane < todo;
        if (nvalid) {
            const uint32_t id = a.point_list[range.x + base + lane];
            const float4* r = a.recs + 3 * (size_t)id;
            n0 = __ldg(r); n1 = __ldg(r + 1); n2 = __ldg(r + 2);
        }
    };
    if (todo > 0) fetch(0);
    for (int base = 0; base < todo; base += 32) {
        bool all_done = true;
#pragma unroll
        for (int q = 0; q < NP; q++) {
            float yA, yB;
            upk(npy[q], yA, yB);
            all_done = all_done && yA < -0.5f * FAR_Y && yB < -0.5f * FAR_Y;
        }
        if (__all_sync(FULL, all_done)) break;
        const float4 c0 = n0, c1 = n1, c2 = n2;
        const bool keep = nvalid && rect_may_contribute(c0.x, c0.y, c0.z, c0.w, c1.x, c2.y, rx0, ry0, rx1, ry1);
        if (base + 32 < todo) fetch(base + 32);
        const unsigned m = __ballot_sync(FULL, keep);
        const int n = __popc(m);
        // the backward walks the same list for the same region: hand it the survivors of this batch (bit l = entry base + l)
        if (mask_out && lane == 0) mask_out[(size_t)(base >> 5) * NW] = m;
        __syncwarp();                                // every lane is done reading the previous batch
        if (keep) {
            const int slot = __popc(m & ((1u << lane) - 1u));
            q0s[slot] = make_float4(c0.x, c0.y, c0.z, -c0.w);
            q1s[slot] = make_float4(c1.x, c1.y, c2.y, __uint_as_float((uint32_t)(base + lane) + 1u));
            q2s[slot] = make_float4(c1.z, c1.w, c2.x, 0.0f);
        }
        __syncwarp();
        for (int j = 0; j < n; j++) {
            const float4 g0 = q0s[j];
            const float4 g1 = q1s[j];
            const float dx = g0.x - pxf;
            const float s1 = FMUL(dx, g0.z), s2 = FMUL(dx, g0.w);
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const f32x2 dy = add2(pk1(g0.y), npy[q]);
                const f32x2 pw = power2_exact(dx, s1, s2, dy, g1.x);
                float pA, pB;
                upk(pw, pA, pB);
                const bool okA = !(pA > 0.0f || pA < g1.z);
                const bool okB = !(pB > 0.0f || pB < g1.z);
                // warp-uniform skips: a lane whose pixels reject runs the same straight-line code with alpha 0
                if (!STRAIGHT) { if (!__any_sync(FULL, okA || okB)) continue; }
                const f32x2 al = mul2(pk1(g1.y), exp2_exact(pw));
                float aA, aB, tA, tB;
                upk(al, aA, aB);
                aA = fminf(0.99f, aA); aB = fminf(0.99f, aB);
                const bool bA = okA && !(aA < 1.0f / 255.0f);
                const bool bB = okB && !(aB < 1.0f / 255.0f);
                upk(mul2(T[q], rsub2(pk(aA, aB), 1.0f)), tA, tB);      // test_T
                const bool dA = bA && tA < 0.0001f, dB = bB && tB < 0.0001f;   // done now: park the pixel
                const bool cA = bA && !dA, cB = bB && !dB;
                float yA, yB;
                upk(npy[q], yA, yB);
                npy[q] = pk(dA ? -FAR_Y : yA, dB ? -FAR_Y : yB);
                if (!STRAIGHT) { if (!__any_sync(FULL, cA || cB)) continue; }
                // a pixel that does not blend this Gaussian runs with alpha 0: T * 1, C + T * (0 * colour)
                const f32x2 w = pk(cA ? aA : 0.0f, cB ? aB : 0.0f);
                const float4 col = q2s[j];
                C0[q] = fma2(T[q], mul2(w, pk1(col.x)), C0[q]);
                C1[q] = fma2(T[q], mul2(w, pk1(col.y)), C1[q]);
                C2[q] = fma2(T[q], mul2(w, pk1(col.z)), C2[q]);
                if (TMUL) {
                    T[q] = mul2(T[q], rsub2(w, 1.0f));
                } else {
                    float TA, TB;
                    upk(T[q], TA, TB);
                    T[q] = pk(cA ? tA : TA, cB ? tB : TB);
                }
                const uint32_t pos1 = __float_as_uint(g1.w);
                lastA[q] = cA ? pos1 : lastA[q];
                lastB[q] = cB ? pos1 : lastB[q];
            }
        }
    }
    const size_t HW = (size_t)a.H * a.W;
#pragma unroll
    for (int q = 0; q < NP; q++) {
        float TA, TB, r0, r1, g0, g1, b0, b1;
        upk(T[q], TA, TB); upk(C0[q], r0, r1); upk(C1[q], g0, g1); upk(C2[q], b0, b1);
        if (inA[q]) {
            const size_t pix = (size_t)pyi[q] * a.W + px;
            a.final_T[pix] = TA;
            a.n_contrib[pix] = lastA[q];
            a.out_color[pix] = fmaf(TA, a.bg[0], r0);
            a.out_color[HW + pix] = fmaf(TA, a.bg[1], g0);
            a.out_color[2 * HW + pix] = fmaf(TA, a.bg[2], b0);
        }
        if (inB[q]) {
            const size_t pix = (size_t)(pyi[q] + 1) * a.W + px;
            a.final_T[pix] = TB;
            a.n_contrib[pix] = lastB[q];
            a.out_color[pix] = fmaf(TB, a.bg[0], r1);
            a.out_color[HW + pix] = fmaf(TB, a.bg[1], g1);
            a.out_color[2 * HW + pix] = fmaf(TB, a.bg[2], b1);
        }
    }
}

// ---------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------
// Sum v[0..8] over the warp.  On return lane 4k (k = 0..7) holds the total of v[k] in v[0];
// every lane holds the total of v[8] in v[8].  (Transposing butterfly: 14 shuffles.)
__device__ __forceinline__ void warp_reduce9_v2(float (&v)[9], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = up ? v[i] : v[i + 4];
            const float keep = up ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = up ? v[i] : v[i + 2];
            const float keep = up ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 8);
        }
    }
    {
        const bool up = lane & 4;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 2);
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[8] += __shfl_xor_sync(FULL, v[8], o);
}

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Per pixel, back to front (backward.cu:443-548).  With d_j = colour_j . dL/dpixel the reference's
// three-channel "accumulated colour behind j" recurrence (with its pending last_alpha/last_color)
// collapses to one scalar per pixel, updated right after use:
//   dL/dalpha_j = T_j (d_j - a_j),   a_{j-1} = alpha_j d_j + (1 - alpha_j) a_j = fma(alpha_j, d_j - a_j, a_j)
// (same sum, associated per pixel instead of per channel; gradients carry a 1e-4 tolerance).
// A pixel that does not blend Gaussian j runs with alpha = G = 0: T, a and every sum are then
// unchanged exactly, so the loop body is straight-line code behind two warp-uniform skips.
// SMEM_RED: the nine per-lane sums of up to four consecutive Gaussians are parked in a warp-private
// shared buffer and reduced together - lane (e, o) = (entry, octant) adds the 32 lane values of sum o of
// entry e with eight LDS.128 - instead of a 14-shuffle butterfly per Gaussian (33 vs 59 instructions
// per blended (region, Gaussian)).
constexpr int RED_E = 4, RED_STRIDE = 36;             // entries per flush; padded row (ncu: bank conflicts on 1.6 % of the kernel's shared wavefronts)

// (A CTA per region, as in the forward, measured 3-4 % SLOWER here: 0.609 vs 0.585 ms.)
template <int NP, int MINB, bool SMEM_RED, bool STRAIGHT, bool MASKS = false>
__global__ void __launch_bounds__(128 / NP, MINB) blend_bwd_v2_kernel(BlendBwdArgs a) {
    constexpr int NW = 4 / NP;
    __shared__ __align__(16) float s_red[SMEM_RED ? NW : 1][SMEM_RED ? RED_E * 9 * RED_STRIDE : 1];
    __shared__ uint32_t s_rid[NW][RED_E];
    __shared__ float4 s_q0[NW][32];                  // x, y, conic.x, -conic.y
    __shared__ float4 s_q1[NW][32];                  // conic.z, opacity, cut, list position as bits
    __shared__ float4 s_q2[NW][32];                  // r, g, b, Gaussian id as bits
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.y * a.grid_x + blockIdx.x;
    const int X0 = blockIdx.x * GSR_TILE + ((NP == 1) ? ((warp & 1) << 3) : (warp << 3));
    const int Y0 = blockIdx.y * GSR_TILE + ((NP == 1) ? ((warp >> 1) << 3) : 0);
    if (X0 >= a.W || Y0 >= a.H) return;
    const int px = X0 + (lane & 7);
    const float pxf = (float)px;
    const float rx0 = (float)X0, ry0 = (float)Y0;
    const float rx1 = fminf(rx0 + 7.0f, (float)(a.W - 1)), ry1 = fminf(ry0 + (float)(8 * NP - 1), (float)(a.H - 1));
    const size_t HW = (size_t)a.H * a.W;
    const bool has_bg = (a.bg[0] != 0.0f) || (a.bg[1] != 0.0f) || (a.bg[2] != 0.0f);

    f32x2 npy[NP], T[NP], nTfin[NP], dp0[NP], dp1[NP], dp2[NP], bgdot[NP], acc[NP];
    uint32_t lastA[NP], lastB[NP];
    uint32_t wlast = 0;
#pragma unroll
    for (int q = 0; q < NP; q++) {
        const int py = Y0 + 8 * q + 2 * (lane >> 3);
        npy[q] = pk(-(float)py, -(float)(py + 1));
        const bool inA = px < a.W && py < a.H, inB = px < a.W && py + 1 < a.H;
        const size_t pA = (size_t)py * a.W + px, pB = pA + a.W;
        const float TA = inA ? a.final_T[pA] : 0.0f, TB = inB ? a.final_T[pB] : 0.0f;
        lastA[q] = inA ? a.n_contrib[pA] : 0u;
        lastB[q] = inB ? a.n_contrib[pB] : 0u;
        float d[6] = {0, 0, 0, 0, 0, 0};
        if (inA) { d[0] = a.dL_dpix[pA]; d[2] = a.dL_dpix[HW + pA]; d[4] = a.dL_dpix[2 * HW + pA]; }
        if (inB) { d[1] = a.dL_dpix[pB]; d[3] = a.dL_dpix[HW + pB]; d[5] = a.dL_dpix[2 * HW + pB]; }
        T[q] = pk(TA, TB); nTfin[q] = pk(-TA, -TB);
        dp0[q] = pk(d[0], d[1]); dp1[q] = pk(d[2], d[3]); dp2[q] = pk(d[4], d[5]);
        bgdot[q] = pk(a.bg[0] * d[0] + a.bg[1] * d[2] + a.bg[2] * d[4], a.bg[0] * d[1] + a.bg[1] * d[3] + a.bg[2] * d[5]);
        acc[q] = pk1(0.0f);
        wlast = max(wlast, max(lastA[q], lastB[q]));
    }
    // nothing behind the region's deepest contributor matters to any of its pixels
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wlast = max(wlast, __shfl_xor_sync(FULL, wlast, o));
    const uint2 range = a.ranges[tile];
    float4* const q0s = s_q0[warp]; float4* const q1s = s_q1[warp]; float4* const q2s = s_q2[warp];
    float* const grad_base = reinterpret_cast<float*>(a.grad_recs);

    // With the forward's region masks the list is walked in the forward's 32-entry batches (back to front: lane l takes
    // position 32 b + 31 - l) and only the survivors of the forward's cull are fetched at all; without them the cull is
    // repeated here on batches counted back from the deepest contributor.
    const uint32_t* const mask_in = MASKS ? a.region_masks + (size_t)NW * ((size_t)(range.x >> 5) + (size_t)tile) + warp : nullptr;
    float4 n0, n1, n2;
    uint32_t nid = 0;
    bool nvalid = false;
    int npos = 0;
    auto fetch = [&](int start) {          // positions start-1, start-2, ... (back to front)
        const int pos = start - 1 - lane;
        npos = pos;
        nvalid = pos >= 0;
        if (MASKS) {
            const uint32_t mk = __ldg(mask_in + (size_t)((start - 1) >> 5) * NW);       // start is a multiple of 32
            nvalid = nvalid && ((mk >> (pos & 31)) & 1u) && pos < (int)wlast;
        }
        if (nvalid) {
            nid = a.point_list[range.x + pos];
            const float4* r = a.recs + 3 * (size_t)nid;
            n0 = __ldg(r); n1 = __ldg(r + 1); n2 = __ldg(r + 2);
        }
    };
    float* const red = s_red[SMEM_RED ? warp : 0];
    int ne = 0;
    auto flush = [&](int count) {
        __syncwarp();
        const int e = lane >> 3, o = lane & 7;
        // every lane computes (rows of entries >= count hold stale values); only the atomics are predicated
        const float* rows = red + e * (9 * RED_STRIDE);
        const float4* r4 = reinterpret_cast<const float4*>(rows + o * RED_STRIDE);
        f32x2 s0 = pk1(0.0f), s1 = pk1(0.0f);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 x = r4[i];
            s0 = add2(s0, pk(x.x, x.y));
            s1 = add2(s1, pk(x.z, x.w));
        }
        float a0, a1, b0, b1;
        upk(s0, a0, a1); upk(s1, b0, b1);
        const float sum = (a0 + a1) + (b0 + b1);
        // ninth sum: each octant lane adds four of its 32 values, then three shuffles inside the octet
        const float4 y = reinterpret_cast<const float4*>(rows + 8 * RED_STRIDE)[o];
        float s8 = (y.x + y.y) + (y.z + y.w);
        s8 += __shfl_xor_sync(FULL, s8, 4);
        s8 += __shfl_xor_sync(FULL, s8, 2);
        s8 += __shfl_xor_sync(FULL, s8, 1);
        if (e < count) {
            float* dst = grad_base + 12 * (size_t)s_rid[warp][e];
            atomicAdd(dst + o, sum);
            if (o == 0) atomicAdd(dst + 8, s8);
        }
        __syncwarp();
    };
    const int start0 = MASKS ? (int)((wlast + 31u) & ~31u) : (int)wlast;
    if (wlast > 0) fetch(start0);
    for (int start = start0; start > 0; start -= 32) {
        const float4 c0 = n0, c1 = n1, c2 = n2;
        const uint32_t cid = nid;
        const int cpos = npos;
        bool keep = nvalid;
        if (!MASKS) keep = keep && rect_may_contribute(c0.x, c0.y, c0.z, c0.w, c1.x, c2.y, rx0, ry0, rx1, ry1);
        if (start - 32 > 0) fetch(start - 32);
        const unsigned m = __ballot_sync(FULL, keep);
        const int n = __popc(m);
        __syncwarp();
        if (keep) {
            const int slot = __popc(m & ((1u << lane) - 1u));
            q0s[slot] = make_float4(c0.x, c0.y, c0.z, -c0.w);
            q1s[slot] = make_float4(c1.x, c1.y, c2.y, __uint_as_float((uint32_t)cpos));
            q2s[slot] = make_float4(c1.z, c1.w, c2.x, __uint_as_float(cid));
        }
        __syncwarp();
        for (int j = 0; j < n; j++) {
            const float4 g0 = q0s[j];
            const float4 g1 = q1s[j];
            const uint32_t pos_j = __float_as_uint(g1.w);
            const float dx = g0.x - pxf;
            const float s1 = FMUL(dx, g0.z), s2 = FMUL(dx, g0.w);
            // Every branch below is warp-uniform: a lane whose pixels do not blend Gaussian j runs the
            // same straight-line code with alpha = G = 0 (a per-lane branch would save no issue slot).
            f32x2 dy[NP], pw[NP];
            bool okA[NP], okB[NP];
            bool any_ok = false;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                dy[q] = add2(pk1(g0.y), npy[q]);
                pw[q] = power2_exact(dx, s1, s2, dy[q], g1.x);
                float pA, pB;
                upk(pw[q], pA, pB);
                okA[q] = pos_j < lastA[q] && !(pA > 0.0f || pA < g1.z);
                okB[q] = pos_j < lastB[q] && !(pB > 0.0f || pB < g1.z);
                any_ok = any_ok || okA[q] || okB[q];
            }
            if (!STRAIGHT) { if (!__any_sync(FULL, any_ok)) continue; }
            f32x2 G[NP], alpha[NP];
            bool any_act = false;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const f32x2 Gp = exp2_exact(pw[q]);
                float GA, GB, aA, aB;
                upk(Gp, GA, GB);
                upk(mul2(pk1(g1.y), Gp), aA, aB);
                aA = fminf(0.99f, aA); aB = fminf(0.99f, aB);
                const bool actA = okA[q] && !(aA < 1.0f / 255.0f);
                const bool actB = okB[q] && !(aB < 1.0f / 255.0f);
                any_act = any_act || actA || actB;
                G[q] = pk(actA ? GA : 0.0f, actB ? GB : 0.0f);
                alpha[q] = pk(actA ? aA : 0.0f, actB ? aB : 0.0f);
            }
            if (!STRAIGHT) { if (!__any_sync(FULL, any_act)) continue; }
            const float4 col = q2s[j];
            f32x2 V[9];
#pragma unroll
            for (int q = 0; q < NP; q++) {
                // T_j = T_{j+1} / (1 - alpha): MUFU.RCP + one Newton step
                const f32x2 om = rsub2(alpha[q], 1.0f);
                float o0, o1;
                upk(om, o0, o1);
                const float r0 = rcp_approx(o0), r1 = rcp_approx(o1);
                const f32x2 r = pk(r0, r1);
                const f32x2 inv = fma2(r, fma2(om, pk(-r0, -r1), pk1(1.0f)), r);
                T[q] = mul2(T[q], inv);
                const f32x2 dcd = mul2(alpha[q], T[q]);                   // dchannel_dcolor
                const f32x2 d = fma2(pk1(col.z), dp2[q], fma2(pk1(col.y), dp1[q], mul2(pk1(col.x), dp0[q])));
                const f32x2 diff = fma2(acc[q], pk1(-1.0f), d);           // d_j - a_j
                f32x2 dL_dalpha = mul2(diff, T[q]);
                acc[q] = fma2(alpha[q], diff, acc[q]);                    // a_{j-1} = alpha_j d_j + (1 - alpha_j) a_j
                if (has_bg) dL_dalpha = fma2(mul2(nTfin[q], inv), bgdot[q], dL_dalpha);
                // moment form of the per-Gaussian sums (common.cuh): w = G dL/dalpha
                const f32x2 w = mul2(G[q], dL_dalpha);
                const f32x2 wx = mul2(w, pk1(dx)), wy = mul2(w, dy[q]);
                if (q == 0) {
                    V[0] = w; V[1] = wx; V[2] = wy;
                    V[3] = mul2(wx, pk1(dx));
                    V[4] = mul2(wx, dy[q]);
                    V[5] = mul2(wy, dy[q]);
                    V[6] = mul2(dcd, dp0[q]);
                    V[7] = mul2(dcd, dp1[q]);
                    V[8] = mul2(dcd, dp2[q]);
                } else {
                    V[0] = add2(V[0], w); V[1] = add2(V[1], wx); V[2] = add2(V[2], wy);
                    V[3] = fma2(wx, pk1(dx), V[3]);
                    V[4] = fma2(wx, dy[q], V[4]);
                    V[5] = fma2(wy, dy[q], V[5]);
                    V[6] = fma2(dcd, dp0[q], V[6]);
                    V[7] = fma2(dcd, dp1[q], V[7]);
                    V[8] = fma2(dcd, dp2[q], V[8]);
                }
            }
            float v[9];
#pragma unroll
            for (int k = 0; k < 9; k++) { float x0, x1; upk(V[k], x0, x1); v[k] = x0 + x1; }
            if (SMEM_RED) {
                float* slot = red + ne * (9 * RED_STRIDE) + lane;
#pragma unroll
                for (int k = 0; k < 9; k++) slot[k * RED_STRIDE] = v[k];
                if (lane == 0) s_rid[warp][ne] = __float_as_uint(col.w);
                if (++ne == RED_E) { flush(ne); ne = 0; }
            } else {
                warp_reduce9_v2(v, lane);
                float* dst = grad_base + 12 * (size_t)__float_as_uint(col.w);
                if ((lane & 3) == 0) atomicAdd(dst + (lane >> 2), v[0]);
                else if (lane == 1) atomicAdd(dst + 8, v[8]);
            }
        }
    }
    if (SMEM_RED && ne) flush(ne);
}

// Debug: replay blend_fwd_v2's walk of every 8x8 region and count the loop iterations that a finer-grained cull would
// leave, for several ways of splitting the warp's 32 lanes into groups that walk separately compacted lists.
// Configs c (groups x shape in pixels): 0: 1 x 8x8, 1: 2 x 8x4, 2: 4 x 4x4, 3: 4 x 8x2, 4: 2 x 4x8, 5: 8 x 4x2.
// out[4c+0] = sum over groups of surviving entries; out[4c+1] = sum over batches of the largest group count (groups in
// lockstep per 32-entry batch); out[4c+2] = sum over regions of the largest group total (ideal queues); out[4c+3] = batches.
// out[24] = (lane, entry) with a blending pixel, out[25] = (pixel, entry) blended, out[26] = (pixel, entry) evaluated by config 0.
__global__ void __launch_bounds__(128) blend_group_stats_kernel(BlendFwdArgs a, unsigned long long* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.y * a.grid_x + blockIdx.x;
    const int X0 = blockIdx.x * GSR_TILE + ((warp & 1) << 3), Y0 = blockIdx.y * GSR_TILE + ((warp >> 1) << 3);
    if (X0 >= a.W || Y0 >= a.H) return;
    const int col = lane & 7, prow = lane >> 3;
    const int px = X0 + col, pyA = Y0 + 2 * prow, pyB = pyA + 1;
    bool doneA = !(px < a.W && pyA < a.H), doneB = !(px < a.W && pyB < a.H);
    float TA = 1.0f, TB = 1.0f;
    constexpr int NC = 6;
    const int ngroups[NC] = {1, 2, 4, 4, 2, 8};
    // group of a lane / pixel rectangle of a group, per config
    auto group_of = [&](int c, int l) -> int {
        const int cl = l & 7, pr = l >> 3;
        switch (c) {
            case 0: return 0;
            case 1: return pr >> 1;
            case 2: return (pr >> 1) * 2 + (cl >> 2);
            case 3: return pr;
            case 4: return cl >> 2;
            default: return pr * 2 + (cl >> 2);
        }
    };
    auto group_rect = [&](int c, int g, float& x0, float& y0, float& x1, float& y1) {
        int gx = 0, gy = 0, w = 8, h = 8;
        switch (c) {
            case 0: break;
            case 1: gy = 4 * g; h = 4; break;
            case 2: gx = 4 * (g & 1); gy = 4 * (g >> 1); w = 4; h = 4; break;
            case 3: gy = 2 * g; h = 2; break;
            case 4: gx = 4 * g; w = 4; break;
            default: gx = 4 * (g & 1); gy = 2 * (g >> 1); w = 4; h = 2; break;
        }
        x0 = (float)(X0 + gx); y0 = (float)(Y0 + gy);
        x1 = fminf(x0 + (float)(w - 1), (float)(a.W - 1)); y1 = fminf(y0 + (float)(h - 1), (float)(a.H - 1));
    };
    unsigned gmask[NC][8];
    for (int c = 0; c < NC; c++)
        for (int g = 0; g < ngroups[c]; g++) gmask[c][g] = __ballot_sync(FULL, group_of(c, lane) == g);
    unsigned long long sum[NC], bmax[NC], tot[NC][8], batches = 0, lane_blend = 0, pix_blend = 0, pix_eval = 0;
    for (int c = 0; c < NC; c++) { sum[c] = bmax[c] = 0; for (int g = 0; g < 8; g++) tot[c][g] = 0; }
    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);
    for (int base = 0; base < todo; base += 32) {
        const unsigned done_m = __ballot_sync(FULL, doneA && doneB);
        if (done_m == FULL) break;
        batches++;
        const bool valid = base + lane < todo;
        float4 c0 = make_float4(0, 0, 0, 0), c1 = c0, c2 = c0;
        if (valid) {
            const uint32_t id = a.point_list[range.x + base + lane];
            const float4* r = a.recs + 3 * (size_t)id;
            c0 = r[0]; c1 = r[1]; c2 = r[2];
        }
        unsigned m0 = 0;
        for (int c = 0; c < NC; c++) {
            unsigned long long mx = 0;
            for (int g = 0; g < ngroups[c]; g++) {
                float x0, y0, x1, y1;
                group_rect(c, g, x0, y0, x1, y1);
                const bool gdone = (done_m & gmask[c][g]) == gmask[c][g];
                const bool keep = valid && !gdone && rect_may_contribute(c0.x, c0.y, c0.z, c0.w, c1.x, c2.y, x0, y0, x1, y1);
                const unsigned m = __ballot_sync(FULL, keep);
                const unsigned long long n = __popc(m);
                if (c == 0) m0 = m;
                sum[c] += n; tot[c][g] += n; mx = n > mx ? n : mx;
            }
            bmax[c] += mx;
        }
        // blend the survivors of the whole-region cull in order, exactly as the forward does
        for (unsigned m = m0; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const float gx = __shfl_sync(FULL, c0.x, src), gy = __shfl_sync(FULL, c0.y, src);
            const float cx = __shfl_sync(FULL, c0.z, src), cy = __shfl_sync(FULL, c0.w, src), cz = __shfl_sync(FULL, c1.x, src);
            const float op = __shfl_sync(FULL, c1.y, src), cut = __shfl_sync(FULL, c2.y, src);
            const float dx = gx - (float)px;
            bool blA = false, blB = false;
            if (!doneA) {
                pix_eval++;
                const float power = blend_power_exact(dx, gy - (float)pyA, cx, cy, cz);
                if (!(power > 0.0f || power < cut)) {
                    const float alpha = fminf(0.99f, op * expf(power));
                    if (!(alpha < 1.0f / 255.0f)) {
                        const float t = TA * (1.0f - alpha);
                        if (t < 0.0001f) doneA = true; else { TA = t; blA = true; }
                    }
                }
            }
            if (!doneB) {
                pix_eval++;
                const float power = blend_power_exact(dx, gy - (float)pyB, cx, cy, cz);
                if (!(power > 0.0f || power < cut)) {
                    const float alpha = fminf(0.99f, op * expf(power));
                    if (!(alpha < 1.0f / 255.0f)) {
                        const float t = TB * (1.0f - alpha);
                        if (t < 0.0001f) doneB = true; else { TB = t; blB = true; }
                    }
                }
            }
            lane_blend += (blA || blB); pix_blend += (int)blA + (int)blB;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lane_blend += __shfl_xor_sync(FULL, lane_blend, o);
        pix_blend += __shfl_xor_sync(FULL, pix_blend, o);
        pix_eval += __shfl_xor_sync(FULL, pix_eval, o);
    }
    if (lane == 0) {
        for (int c = 0; c < NC; c++) {
            unsigned long long mx = 0;
            for (int g = 0; g < ngroups[c]; g++) mx = tot[c][g] > mx ? tot[c][g] : mx;
            atomicAdd(&out[4 * c + 0], sum[c]); atomicAdd(&out[4 * c + 1], bmax[c]); atomicAdd(&out[4 * c + 2], mx);
            atomicAdd(&out[4 * c + 3], batches);
        }
        atomicAdd(&out[24], lane_blend); atomicAdd(&out[25], pix_blend); atomicAdd(&out[26], pix_eval);
    }
}

// exhaustive check kernel: exp1_exact / exp2_exact against expf on every float in [lo_bits, hi_bits]
__global__ void exp_check_kernel(uint32_t lo_bits, uint32_t hi_bits, unsigned long long* mismatches) {
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long bad1 = 0, bad2 = 0;
    for (uint64_t b = (uint64_t)lo_bits + blockIdx.x * blockDim.x + threadIdx.x; b <= hi_bits; b += stride) {
        const float x = __uint_as_float((uint32_t)b);
        const float r = expf(x);
        const float e1 = exp1_exact(x);
        float e2a, e2b;
        const float xb = x * 0.5f;
        upk(exp2_exact(pk(x, xb)), e2a, e2b);
        const float rb = expf(xb);
        bad1 += __float_as_uint(r) != __float_as_uint(e1);
        bad2 += (__float_as_uint(r) != __float_as_uint(e2a)) || (__float_as_uint(rb) != __float_as_uint(e2b));
    }
    if (bad1) atomicAdd(&mismatches[0], bad1);
    if (bad2) atomicAdd(&mismatches[1], bad2);
}

int env_int_v2(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

// The forward's region masks are usable by the backward only when both run the second-generation kernels with the same
// region shape (GSR_BLEND_MASKS=0 switches them off: the backward then repeats the cull, as in round 1).
int gsr_blend_region_masks_enabled() {
    static const int on = env_int_v2("GSR_BLEND_MASKS", 1) && env_int_v2("GSR_BLEND_FWD_V", 2) == 2 && env_int_v2("GSR_BLEND_BWD_V", 2) == 2 &&
                          env_int_v2("GSR_FWD_NP", 1) == 1 && env_int_v2("GSR_BWD_NP", 1) == 1 && env_int_v2("GSR_BWD_MINB", 0) < 8 &&
                          env_int_v2("GSR_BWD_SMEM_RED", 1) && env_int_v2("GSR_BWD_STRAIGHT", 1);      // the default backward variant reads them
    return on;
}

int gsr_launch_blend_fwd_v2(const BlendFwdArgs& a, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    static const int np = env_int_v2("GSR_FWD_NP", 1);
    { GsrProfScope prof_("blend_fwd", stream);
    static const int tmul = env_int_v2("GSR_FWD_TMUL", 0);   // T *= (1 - w) instead of a select: measured 1 % slower
    static const int straight = env_int_v2("GSR_FWD_STRAIGHT", 1);
    static const int minb = env_int_v2("GSR_FWD_MINB", 7);   // 72 regs: 7 CTAs/SM
    // one 8x8 region per CTA by default, regions of a tile adjacent in a 1-D grid: 0.343 vs 0.354 ms per tile CTA at C2
    // (2-D region grid: 0.349; the same variant of the backward is 3-4 % SLOWER)
    static const int wpc = env_int_v2("GSR_FWD_WPC", 1);
    if (np != 2 && wpc == 1 && straight && !tmul) {
        const dim3 rgrid(4u * (unsigned)a.grid_x * (unsigned)a.grid_y, 1, 1);
        blend_fwd_v2_kernel<1, 28, false, true, 1><<<rgrid, 32, 0, stream>>>(a);
    } else
    if (np == 2) blend_fwd_v2_kernel<2, 0, false, false><<<grid, 64, 0, stream>>>(a);
    else if (straight && tmul) blend_fwd_v2_kernel<1, 0, true, true><<<grid, 128, 0, stream>>>(a);
    else if (straight && minb == 7) blend_fwd_v2_kernel<1, 7, false, true><<<grid, 128, 0, stream>>>(a);
    else if (straight && minb == 8) blend_fwd_v2_kernel<1, 8, false, true><<<grid, 128, 0, stream>>>(a);
    else if (straight) blend_fwd_v2_kernel<1, 0, false, true><<<grid, 128, 0, stream>>>(a);
    else if (tmul) blend_fwd_v2_kernel<1, 0, true, false><<<grid, 128, 0, stream>>>(a);
    else blend_fwd_v2_kernel<1, 0, false, false><<<grid, 128, 0, stream>>>(a); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_blend_bwd_v2(const BlendBwdArgs& a, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    static const int np = env_int_v2("GSR_BWD_NP", 1);
    { GsrProfScope prof_("blend_bwd", stream);
    static const int minb = env_int_v2("GSR_BWD_MINB", 0);
    static const int sred = env_int_v2("GSR_BWD_SMEM_RED", 1);
    static const int straight = env_int_v2("GSR_BWD_STRAIGHT", 1);
    if (np == 2) {
        if (sred) blend_bwd_v2_kernel<2, 0, true, false><<<grid, 64, 0, stream>>>(a);
        else blend_bwd_v2_kernel<2, 0, false, false><<<grid, 64, 0, stream>>>(a);
    } else if (minb >= 8) {
        blend_bwd_v2_kernel<1, 8, false, false><<<grid, 128, 0, stream>>>(a);
    } else if (straight && sred) {
        if (a.region_masks) blend_bwd_v2_kernel<1, 0, true, true, true><<<grid, 128, 0, stream>>>(a);
        else blend_bwd_v2_kernel<1, 0, true, true><<<grid, 128, 0, stream>>>(a);
    } else {
        if (sred) blend_bwd_v2_kernel<1, 0, true, false><<<grid, 128, 0, stream>>>(a);
        else blend_bwd_v2_kernel<1, 0, false, false><<<grid, 128, 0, stream>>>(a);
    } }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_blend_group_stats(const BlendFwdArgs& a, unsigned long long* out32, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    blend_group_stats_kernel<<<grid, 128, 0, stream>>>(a, out32);
    GSR_CHECK_LAUNCH();
    return 0;
}

// out[0] = floats in [-x_max, -0] where exp1_exact != expf, out[1] = where the packed version differs
int gsr_launch_exp_check(float x_max, unsigned long long* out2, cudaStream_t stream) {
    GSR_CHECK(cudaMemsetAsync(out2, 0, 16, stream));
    uint32_t hi_bits;
    const float neg = -x_max;
    memcpy(&hi_bits, &neg, 4);
    exp_check_kernel<<<148 * 8, 256, 0, stream>>>(0x80000000u, hi_bits, out2);
    GSR_CHECK_LAUNCH();
    return 0;
}
