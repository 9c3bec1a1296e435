// Tile blending, forward and backward.
//
// Replaces cuda_rasterizer/forward.cu:261-374 (renderCUDA fwd) and
// backward.cu:399-557 (renderCUDA bwd).  One CTA per 16x16 tile, one thread per
// pixel; a warp owns an 8x4 pixel patch (compact patches diverge less than the
// reference's 16x2 rows).  FP32-pipe / shared-memory bound, not HBM bound.
//
// What is kept bit-identical to the reference: the exponent `power`, expf, alpha,
// the three reject tests and the T recurrence - every decision a pixel takes is
// the reference's decision (a flipped 1/255 test would move a pixel by ~4e-3).
// What is redesigned:
//   * each 256-entry batch of the tile's list is CULLED while it is staged: an
//     entry whose Gaussian cannot reach alpha >= 1/255 anywhere inside this tile
//     (closed-form minimum of the conic form over the tile rectangle, with a
//     rounding-safe margin) is dropped by a stable ballot compaction, so the
//     256 pixel threads never loop over it.  The reference's lists come from a
//     3-sigma circle's bounding square and are mostly such entries;
//   * surviving pixels reject with `power < cut` (cut = log(1/(255 opacity)) minus a
//     margin, precomputed per Gaussian) before paying for expf;
//   * colours ride in the staged record (no global load in the inner loop);
//   * the next batch's gathers are issued before the current batch is blended;
//   * backward: per-Gaussian gradients are reduced across the warp with a
//     transposing butterfly (14 shuffles for 9 values instead of 45), summed
//     across the CTA's warps in shared memory, and leave the SM as two 128-bit
//     vector reductions + one scalar per (tile, Gaussian) - the reference issues
//     9 scalar atomics per (pixel, Gaussian).
#include "geom_exact.cuh"
#include "kernels.cuh"

namespace {

constexpr int BLK = 256;
constexpr unsigned FULL = 0xffffffffu;

// Pixel of thread `tid` inside the tile: warp w -> patch (w&1, w>>1) of 8x4 pixels.
__device__ __forceinline__ void tile_pixel(int tid, int& lx, int& ly) {
    const int w = tid >> 5, l = tid & 31;
    lx = ((w & 1) << 3) + (l & 7);
    ly = ((w >> 1) << 2) + (l >> 3);
}

// True unless NO pixel centre in [x0,x1]x[y0,y1] can have power >= cut.
// power(d) = -0.5 (a dx^2 + c dy^2) - b dx dy is concave for a positive-definite
// conic; if the mean lies outside the rectangle the maximum sits on one of the
// (at most two) edges facing the mean, where it is a clamped 1-D parabola.
__device__ __forceinline__ bool tile_may_contribute(float mx, float my, float a, float b, float c, float cut,
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
    // Rounding of the reference's own power evaluation scales with its largest term.
    const float dxm = fmaxf(fabsf(dxl), fabsf(dxh)), dym = fmaxf(fabsf(dyl), fabsf(dyh));
    const float mag = 0.5f * (a * dxm * dxm + c * dym * dym) + fabsf(b) * dxm * dym;
    const float margin = 1e-2f + 4e-6f * mag;
    return !(-0.5f * qmin < cut - margin);
}

// Stable compaction of one flag per thread across the CTA.  Returns this
// thread's output slot (valid when keep) and the CTA total in `total`.
__device__ __forceinline__ int block_compact(bool keep, uint32_t* s_wcount, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(FULL, keep);
    if (lane == 0) s_wcount[warp] = __popc(m);
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BLK / 32; w++) {
        const int c = (int)s_wcount[w];
        base += (w < warp) ? c : 0;
        tot += c;
    }
    total = tot;
    return base + __popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK) blend_fwd_kernel(BlendFwdArgs a) {
    __shared__ float4 s_q0[BLK];     // x, y, conic.x, conic.y
    __shared__ float4 s_q1[BLK];     // conic.z, opacity, cut, (contributor number as bits)
    __shared__ float4 s_q2[BLK];     // r, g, b, -
    __shared__ uint32_t s_wcount[BLK / 32];

    const int tid = threadIdx.x;
    const int tile = blockIdx.y * a.grid_x + blockIdx.x;
    int lx, ly;
    tile_pixel(tid, lx, ly);
    const int px = blockIdx.x * GSR_TILE + lx, py = blockIdx.y * GSR_TILE + ly;
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const float tx0 = (float)(blockIdx.x * GSR_TILE), ty0 = (float)(blockIdx.y * GSR_TILE);
    const float tx1 = fminf(tx0 + 15.0f, (float)(a.W - 1)), ty1 = fminf(ty0 + 15.0f, (float)(a.H - 1));

    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);
    const int rounds = (todo + BLK - 1) / BLK;

    bool done = !inside;
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    uint32_t last_contributor = 0;

    // Software pipeline: gathers of batch i+1 are in flight while batch i is blended.
    float4 n0, n1, n2;
    bool nvalid = false;
    auto fetch = [&](int round) {
        const int pos = round * BLK + tid;
        nvalid = pos < todo;
        if (nvalid) {
            const uint32_t id = a.point_list[range.x + pos];
            const float4* r = a.recs + 3 * (size_t)id;
            n0 = __ldg(r); n1 = __ldg(r + 1); n2 = __ldg(r + 2);
        }
    };
    if (rounds > 0) fetch(0);

    for (int i = 0; i < rounds; i++) {
        // Whole CTA done?  (also the barrier that protects the staging buffers)
        if (__syncthreads_count(done) == BLK) break;
        const float4 q0 = n0, q1 = n1, q2 = n2;
        const bool keep = nvalid && tile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q2.y, tx0, ty0, tx1, ty1);
        const int pos = i * BLK + tid;
        if (i + 1 < rounds) fetch(i + 1);
        int n;
        const int slot = block_compact(keep, s_wcount, n);
        if (keep) {
            s_q0[slot] = q0;
            s_q1[slot] = make_float4(q1.x, q1.y, q2.y, __uint_as_float((uint32_t)pos + 1u));
            s_q2[slot] = make_float4(q1.z, q1.w, q2.x, 0.0f);
        }
        __syncthreads();
        for (int j = 0; !done && j < n; j++) {
            const float4 g0 = s_q0[j];
            const float4 g1 = s_q1[j];
            const float dx = g0.x - pxf, dy = g0.y - pyf;
            const float power = blend_power_exact(dx, dy, g0.z, g0.w, g1.x);
            if (power > 0.0f || power < g1.z) continue;
            const float alpha = fminf(0.99f, g1.y * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float4 col = s_q2[j];
            C0 = fmaf(T, alpha * col.x, C0);
            C1 = fmaf(T, alpha * col.y, C1);
            C2 = fmaf(T, alpha * col.z, C2);
            T = test_T;
            last_contributor = __float_as_uint(g1.w);
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * a.W + px;
        const size_t HW = (size_t)a.H * a.W;
        a.final_T[pix] = T;
        a.n_contrib[pix] = last_contributor;
        a.out_color[pix] = fmaf(T, a.bg[0], C0);
        a.out_color[HW + pix] = fmaf(T, a.bg[1], C1);
        a.out_color[2 * HW + pix] = fmaf(T, a.bg[2], C2);
    }
}

// ---------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------
// Sum v[0..8] over the warp.  On return lane 4k (k = 0..7) holds the total of
// v[k] in v[0]; every lane holds the total of v[8] in v[8].
__device__ __forceinline__ void warp_reduce9(float (&v)[9], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = hi ? v[i] : v[i + 4];
            const float keep = hi ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = hi ? v[i] : v[i + 2];
            const float keep = hi ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    v[0] += __shfl_xor_sync(FULL, v[0], 2);
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[8] += __shfl_xor_sync(FULL, v[8], o);
}

__global__ void __launch_bounds__(BLK) blend_bwd_kernel(BlendBwdArgs a) {
    __shared__ float4 s_q0[BLK];
    __shared__ float4 s_q1[BLK];     // conic.z, opacity, cut, list position as bits
    __shared__ float4 s_q2[BLK];     // r, g, b, -
    __shared__ uint32_t s_id[BLK];
    __shared__ float s_grad[BLK][9];
    __shared__ uint32_t s_touched[BLK];
    __shared__ uint32_t s_wcount[BLK / 32];
    __shared__ uint32_t s_wmax[BLK / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * a.grid_x + blockIdx.x;
    int lx, ly;
    tile_pixel(tid, lx, ly);
    const int px = blockIdx.x * GSR_TILE + lx, py = blockIdx.y * GSR_TILE + ly;
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const float tx0 = (float)(blockIdx.x * GSR_TILE), ty0 = (float)(blockIdx.y * GSR_TILE);
    const float tx1 = fminf(tx0 + 15.0f, (float)(a.W - 1)), ty1 = fminf(ty0 + 15.0f, (float)(a.H - 1));
    const uint2 range = a.ranges[tile];

    const size_t pix = (size_t)py * a.W + px;
    const size_t HW = (size_t)a.H * a.W;
    const float T_final = inside ? a.final_T[pix] : 0.0f;
    const uint32_t last_contributor = inside ? a.n_contrib[pix] : 0u;
    float dpx0 = 0.0f, dpx1 = 0.0f, dpx2 = 0.0f;
    if (inside) { dpx0 = a.dL_dpix[pix]; dpx1 = a.dL_dpix[HW + pix]; dpx2 = a.dL_dpix[2 * HW + pix]; }
    const float bg_dot = a.bg[0] * dpx0 + a.bg[1] * dpx1 + a.bg[2] * dpx2;
    const bool has_bg = (a.bg[0] != 0.0f) || (a.bg[1] != 0.0f) || (a.bg[2] != 0.0f);
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;

    // Nothing behind the tile's deepest contributor matters to any pixel.
    uint32_t m = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
    if (lane == 0) s_wmax[warp] = m;
    __syncthreads();
    uint32_t tile_last = 0;
#pragma unroll
    for (int w = 0; w < BLK / 32; w++) tile_last = max(tile_last, s_wmax[w]);

    float T = T_final;
    float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
    float last_alpha = 0.0f, lc0 = 0.0f, lc1 = 0.0f, lc2 = 0.0f;

    float4 n0, n1, n2;
    uint32_t nid = 0;
    bool nvalid = false;
    auto fetch = [&](int start) {
        const int pos = start - 1 - tid;       // back to front
        nvalid = pos >= 0;
        if (nvalid) {
            nid = a.point_list[range.x + pos];
            const float4* r = a.recs + 3 * (size_t)nid;
            n0 = __ldg(r); n1 = __ldg(r + 1); n2 = __ldg(r + 2);
        }
    };
    if (tile_last > 0) fetch((int)tile_last);

    for (int start = (int)tile_last; start > 0; start -= BLK) {
        const float4 q0 = n0, q1 = n1, q2 = n2;
        const uint32_t id = nid;
        const int pos = start - 1 - tid;
        const bool keep = nvalid && tile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q2.y, tx0, ty0, tx1, ty1);
        if (start - BLK > 0) fetch(start - BLK);
        int n;
        const int slot = block_compact(keep, s_wcount, n);   // contains a barrier: previous flush is complete
        if (keep) {
            s_q0[slot] = q0;
            s_q1[slot] = make_float4(q1.x, q1.y, q2.y, __uint_as_float((uint32_t)pos));
            s_q2[slot] = make_float4(q1.z, q1.w, q2.x, 0.0f);
            s_id[slot] = id;
        }
        for (int k = tid; k < n * 9; k += BLK) (&s_grad[0][0])[k] = 0.0f;
        if (tid < n) s_touched[tid] = 0;
        __syncthreads();

        for (int j = 0; j < n; j++) {
            const float4 g0 = s_q0[j];
            const float4 g1 = s_q1[j];
            const uint32_t pos_j = __float_as_uint(g1.w);
            float v[9];
            bool active = false;
            float dx = 0.f, dy = 0.f, power = 0.f;
            if (pos_j < last_contributor) {
                dx = g0.x - pxf; dy = g0.y - pyf;
                power = blend_power_exact(dx, dy, g0.z, g0.w, g1.x);
                active = !(power > 0.0f || power < g1.z);
            }
            float G = 0.f, alpha = 0.f;
            if (active) {
                G = expf(power);
                alpha = fminf(0.99f, g1.y * G);
                active = !(alpha < 1.0f / 255.0f);
            }
            if (!__any_sync(FULL, active)) continue;
            if (active) {
                const float4 col = s_q2[j];
                const float one_m_alpha = 1.0f - alpha;
                T = T / one_m_alpha;
                const float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.0f;
                acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0; lc0 = col.x;
                dL_dalpha += (col.x - acc0) * dpx0;
                acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1; lc1 = col.y;
                dL_dalpha += (col.y - acc1) * dpx1;
                acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2; lc2 = col.z;
                dL_dalpha += (col.z - acc2) * dpx2;
                dL_dalpha *= T;
                last_alpha = alpha;
                if (has_bg) dL_dalpha += (-T_final / one_m_alpha) * bg_dot;
                const float dL_dG = g1.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                v[0] = dL_dG * dG_ddelx * ddelx_dx;
                v[1] = dL_dG * dG_ddely * ddely_dy;
                v[2] = -0.5f * gdx * dx * dL_dG;
                v[3] = -0.5f * gdx * dy * dL_dG;
                v[4] = -0.5f * gdy * dy * dL_dG;
                v[5] = G * dL_dalpha;
                v[6] = dchannel_dcolor * dpx0;
                v[7] = dchannel_dcolor * dpx1;
                v[8] = dchannel_dcolor * dpx2;
            } else {
#pragma unroll
                for (int k = 0; k < 9; k++) v[k] = 0.0f;
            }
            warp_reduce9(v, lane);
            if ((lane & 3) == 0) atomicAdd(&s_grad[j][lane >> 2], v[0]);
            if (lane == 1) { atomicAdd(&s_grad[j][8], v[8]); s_touched[j] = 1; }
        }
        __syncthreads();
        // Flush the CTA's per-Gaussian sums: two 128-bit vector reductions + one scalar.
        if (tid < n && s_touched[tid]) {
            const float* g = s_grad[tid];
            float4* dst = a.grad_recs + 3 * (size_t)s_id[tid];
            atomicAdd(dst, make_float4(g[0], g[1], g[2], g[3]));
            atomicAdd(dst + 1, make_float4(g[4], g[5], g[6], g[7]));
            atomicAdd(reinterpret_cast<float*>(dst + 2), g[8]);
        }
        // next iteration's block_compact barrier orders this flush before re-staging
    }
}

}  // namespace

int gsr_launch_blend_fwd(const BlendFwdArgs& a, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    { GsrProfScope prof_("blend_fwd", stream);
    blend_fwd_kernel<<<grid, BLK, 0, stream>>>(a); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_blend_bwd(const BlendBwdArgs& a, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    { GsrProfScope prof_("blend_bwd", stream);
    blend_bwd_kernel<<<grid, BLK, 0, stream>>>(a); }
    GSR_CHECK_LAUNCH();
    return 0;
}
