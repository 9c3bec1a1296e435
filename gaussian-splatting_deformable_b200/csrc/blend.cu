// Tile blending, forward and backward - FIRST-GENERATION kernels (CTA-cooperative staging), kept as the
// selectable fallback/baseline of blend_v2.cu (GSR_BLEND_FWD_V=1 / GSR_BLEND_BWD_V=1) and as the home of the
// TMA-staged variant (GSR_BLEND_TMA=1) and of the workload counters (gsr_debug_blend_stats).
//
// Replaces cuda_rasterizer/forward.cu:261-374 (renderCUDA fwd) and
// backward.cu:399-557 (renderCUDA bwd).  One CTA per 16x16 tile; a warp owns PPT 8x4 pixel
// patches (compact patches diverge less than the reference's 16x2 rows).
//
// What is kept bit-identical to the reference: the exponent `power`, expf, alpha,
// the three reject tests and the T recurrence - every decision a pixel takes is
// the reference's decision (a flipped 1/255 test would move a pixel by ~4e-3).
// What is redesigned:
//   * each batch of the tile's list is CULLED while it is staged: an
//     entry whose Gaussian cannot reach alpha >= 1/255 anywhere inside this tile
//     (closed-form minimum of the conic form over the tile rectangle, with a
//     rounding-safe margin) is dropped by a stable ballot compaction, so the
//     pixel threads never loop over it.  The reference's lists come from a
//     3-sigma circle's bounding square and are mostly such entries;
//   * surviving pixels reject with `power < cut` (cut = log(1/(255 opacity)) minus a
//     margin, precomputed per Gaussian) before paying for expf;
//   * colours ride in the staged record (no global load in the inner loop);
//   * the next batch's gathers are issued before the current batch is blended;
//   * backward: per-Gaussian gradients are reduced across the warp with a
//     transposing butterfly (14 shuffles for 9 values instead of 45) and leave the SM as
//     nine native RED.ADD.F32 per (warp, Gaussian) - the reference issues
//     9 scalar atomics per (pixel, Gaussian).
#include "geom_exact.cuh"
#include "kernels.cuh"
#include <stdlib.h>

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
template <int NT = BLK>
__device__ __forceinline__ int block_compact(bool keep, uint32_t* s_wcount, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(FULL, keep);
    if (lane == 0) s_wcount[warp] = __popc(m);
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) {
        const int c = (int)s_wcount[w];
        base += (w < warp) ? c : 0;
        tot += c;
    }
    total = tot;
    return base + __popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------
// TMA (bulk async copy) staging.  Every list entry is one contiguous 48-byte splat
// record, i.e. one `cp.async.bulk` global->shared copy: the thread that owns the
// entry issues it (SASS: UBLKCP), completion is tracked by an mbarrier per stage
// (expect_tx = 48 x entries of the batch), and the batch lands in a 2-stage shared
// ring with no register staging while the previous batch is being blended.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_48(void* dst_smem, const void* src_gmem, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 48, [%2];" ::"r"(
                     smem_addr(dst_smem)), "l"(src_gmem), "r"(smem_addr(bar)) : "memory");
}

// ---------------------------------------------------------------------------
// Forward.  PPT pixels per thread: a CTA has 256/PPT threads; warp w owns the PPT
// consecutive 8x4 patches w*PPT .. w*PPT+PPT-1 (patch p at ((p&1)*8, (p>>1)*4)) and
// lane l owns the same in-patch pixel of each.  Staged records, loop control and the
// contributor bookkeeping are then paid once per PPT pixels.
// ---------------------------------------------------------------------------
template <int PPT>
__device__ __forceinline__ void patch_pixel(int tid, int slot, int& lx, int& ly) {
    const int w = tid >> 5, l = tid & 31;
    const int p = w * PPT + slot;
    lx = ((p & 1) << 3) + (l & 7);
    ly = ((p >> 1) << 2) + (l >> 3);
}

template <int PPT, bool TMA>
__global__ void __launch_bounds__(BLK / PPT) blend_fwd_kernel(BlendFwdArgs a) {
    constexpr int NT = BLK / PPT;
    __shared__ __align__(16) float4 s_raw[TMA ? 2 : 1][TMA ? NT * 3 : 1];   // TMA landing ring (2 stages x NT records)
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ float4 s_q0[NT];     // x, y, conic.x, conic.y
    __shared__ float4 s_q1[NT];     // conic.z, opacity, cut, (contributor number as bits)
    __shared__ float4 s_q2[NT];     // r, g, b, -
    __shared__ uint32_t s_wcount[BLK / 32];

    const int tid = threadIdx.x;
    const int tile = blockIdx.y * a.grid_x + blockIdx.x;
    float pxf[PPT], pyf[PPT], T[PPT], C0[PPT], C1[PPT], C2[PPT];
    uint32_t last_contributor[PPT];
    bool done[PPT], inside[PPT];
    int pxi[PPT], pyi[PPT];
#pragma unroll
    for (int s = 0; s < PPT; s++) {
        int lx, ly;
        patch_pixel<PPT>(tid, s, lx, ly);
        pxi[s] = blockIdx.x * GSR_TILE + lx; pyi[s] = blockIdx.y * GSR_TILE + ly;
        inside[s] = pxi[s] < a.W && pyi[s] < a.H;
        pxf[s] = (float)pxi[s]; pyf[s] = (float)pyi[s];
        done[s] = !inside[s];
        T[s] = 1.0f; C0[s] = C1[s] = C2[s] = 0.0f; last_contributor[s] = 0;
    }
    const float tx0 = (float)(blockIdx.x * GSR_TILE), ty0 = (float)(blockIdx.y * GSR_TILE);
    const float tx1 = fminf(tx0 + 15.0f, (float)(a.W - 1)), ty1 = fminf(ty0 + 15.0f, (float)(a.H - 1));

    const uint2 range = a.ranges[tile];
    const int todo = (int)(range.y - range.x);
    const int rounds = (todo + NT - 1) / NT;

    if (TMA) {
        if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
        __syncthreads();
    }
    // Software pipeline: the gathers of batch i+1 are in flight while batch i is blended
    // (TMA: into the shared ring; otherwise into registers).
    float4 n0, n1, n2;
    bool nvalid = false;
    int issued = 0, consumed = 0;
    auto fetch = [&](int round) {
        const int pos = round * NT + tid;
        nvalid = pos < todo;
        if (TMA) {
            const int st = round & 1;
            if (tid == 0) mbar_arrive_expect_tx(&s_bar[st], 48u * (uint32_t)min(NT, todo - round * NT));
            if (nvalid) {
                const uint32_t id = a.point_list[range.x + pos];
                tma_load_48(&s_raw[st][3 * tid], a.recs + 3 * (size_t)id, &s_bar[st]);
            }
            issued = round + 1;
        } else if (nvalid) {
            const uint32_t id = a.point_list[range.x + pos];
            const float4* r = a.recs + 3 * (size_t)id;
            n0 = __ldg(r); n1 = __ldg(r + 1); n2 = __ldg(r + 2);
        }
    };
    if (rounds > 0) fetch(0);

    for (int i = 0; i < rounds; i++) {
        bool all_done = true;
#pragma unroll
        for (int s = 0; s < PPT; s++) all_done = all_done && done[s];
        // Whole CTA done?  (also the barrier that protects the staging buffers)
        if (__syncthreads_count(all_done) == NT) break;
        const int pos = i * NT + tid;
        if (TMA) {
            const int st = i & 1;
            mbar_wait(&s_bar[st], (uint32_t)(i >> 1) & 1u);
            nvalid = pos < todo;
            if (nvalid) { n0 = s_raw[st][3 * tid]; n1 = s_raw[st][3 * tid + 1]; n2 = s_raw[st][3 * tid + 2]; }
            consumed = i + 1;
        }
        const float4 q0 = n0, q1 = n1, q2 = n2;
        const bool keep = nvalid && tile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q2.y, tx0, ty0, tx1, ty1);
        if (i + 1 < rounds) fetch(i + 1);
        int n;
        const int slot = block_compact<NT>(keep, s_wcount, n);
        if (keep) {
            s_q0[slot] = q0;
            s_q1[slot] = make_float4(q1.x, q1.y, q2.y, __uint_as_float((uint32_t)pos + 1u));
            s_q2[slot] = make_float4(q1.z, q1.w, q2.x, 0.0f);
        }
        __syncthreads();
        for (int j = 0; j < n; j++) {
            if (PPT == 1) { if (done[0]) break; }
            const float4 g0 = s_q0[j];
            const float4 g1 = s_q1[j];
#pragma unroll
            for (int s = 0; s < PPT; s++) {
                if (done[s]) continue;
                const float dx = g0.x - pxf[s], dy = g0.y - pyf[s];
                const float power = blend_power_exact(dx, dy, g0.z, g0.w, g1.x);
                if (power > 0.0f || power < g1.z) continue;
                const float alpha = fminf(0.99f, g1.y * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T[s] * (1.0f - alpha);
                if (test_T < 0.0001f) { done[s] = true; continue; }
                const float4 col = s_q2[j];
                C0[s] = fmaf(T[s], alpha * col.x, C0[s]);
                C1[s] = fmaf(T[s], alpha * col.y, C1[s]);
                C2[s] = fmaf(T[s], alpha * col.z, C2[s]);
                T[s] = test_T;
                last_contributor[s] = __float_as_uint(g1.w);
            }
        }
    }
    // A CTA must not retire with a bulk copy still landing in its shared memory.
    if (TMA && issued > consumed) mbar_wait(&s_bar[consumed & 1], (uint32_t)(consumed >> 1) & 1u);
    const size_t HW = (size_t)a.H * a.W;
#pragma unroll
    for (int s = 0; s < PPT; s++) {
        if (inside[s]) {
            const size_t pix = (size_t)pyi[s] * a.W + pxi[s];
            a.final_T[pix] = T[s];
            a.n_contrib[pix] = last_contributor[s];
            a.out_color[pix] = fmaf(T[s], a.bg[0], C0[s]);
            a.out_color[HW + pix] = fmaf(T[s], a.bg[1], C1[s]);
            a.out_color[2 * HW + pix] = fmaf(T[s], a.bg[2], C2[s]);
        }
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

// 1/x to ~1 ulp for x in [0.01, 1]: MUFU.RCP + one Newton step (the reference's IEEE
// division T/(1-alpha) costs ~10 instructions; gradients tolerate 1e-4 relative).
__device__ __forceinline__ float rcp_nr(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

template <int PPT, int MINB, bool TMA>
__global__ void __launch_bounds__(BLK / PPT, MINB) blend_bwd_kernel(BlendBwdArgs a) {
    constexpr int NT = BLK / PPT;
    __shared__ __align__(16) float4 s_raw[TMA ? 2 : 1][TMA ? NT * 3 : 1];   // TMA landing ring
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ float4 s_q0[NT];
    __shared__ float4 s_q1[NT];     // conic.z, opacity, cut, list position as bits
    __shared__ float4 s_q2[NT];     // r, g, b, Gaussian id as bits
    __shared__ uint32_t s_wcount[BLK / 32];
    __shared__ uint32_t s_wmax[BLK / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.y * a.grid_x + blockIdx.x;
    const float tx0 = (float)(blockIdx.x * GSR_TILE), ty0 = (float)(blockIdx.y * GSR_TILE);
    const float tx1 = fminf(tx0 + 15.0f, (float)(a.W - 1)), ty1 = fminf(ty0 + 15.0f, (float)(a.H - 1));
    const uint2 range = a.ranges[tile];
    const size_t HW = (size_t)a.H * a.W;
    const bool has_bg = (a.bg[0] != 0.0f) || (a.bg[1] != 0.0f) || (a.bg[2] != 0.0f);
    const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;

    float pxf[PPT], pyf[PPT], T[PPT], T_final[PPT], dpx0[PPT], dpx1[PPT], dpx2[PPT], bg_dot[PPT];
    float acc0[PPT], acc1[PPT], acc2[PPT], lc0[PPT], lc1[PPT], lc2[PPT], last_alpha[PPT];
    uint32_t last_contributor[PPT];
    uint32_t m = 0;
#pragma unroll
    for (int s = 0; s < PPT; s++) {
        int lx, ly;
        patch_pixel<PPT>(tid, s, lx, ly);
        const int px = blockIdx.x * GSR_TILE + lx, py = blockIdx.y * GSR_TILE + ly;
        const bool inside = px < a.W && py < a.H;
        pxf[s] = (float)px; pyf[s] = (float)py;
        const size_t pix = (size_t)py * a.W + px;
        T_final[s] = inside ? a.final_T[pix] : 0.0f;
        last_contributor[s] = inside ? a.n_contrib[pix] : 0u;
        dpx0[s] = dpx1[s] = dpx2[s] = 0.0f;
        if (inside) { dpx0[s] = a.dL_dpix[pix]; dpx1[s] = a.dL_dpix[HW + pix]; dpx2[s] = a.dL_dpix[2 * HW + pix]; }
        bg_dot[s] = a.bg[0] * dpx0[s] + a.bg[1] * dpx1[s] + a.bg[2] * dpx2[s];
        T[s] = T_final[s];
        acc0[s] = acc1[s] = acc2[s] = lc0[s] = lc1[s] = lc2[s] = last_alpha[s] = 0.0f;
        m = max(m, last_contributor[s]);
    }
    // Nothing behind the tile's deepest contributor matters to any pixel.
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
    if (lane == 0) s_wmax[warp] = m;
    __syncthreads();
    uint32_t tile_last = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) tile_last = max(tile_last, s_wmax[w]);

    if (TMA) {
        if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
        __syncthreads();
    }
    float4 n0, n1, n2;
    uint32_t nid = 0, cur_id = 0;
    bool nvalid = false;
    auto fetch = [&](int start, int round) {
        const int pos = start - 1 - tid;       // back to front
        nvalid = pos >= 0;
        if (TMA) {
            const int st = round & 1;
            if (tid == 0) mbar_arrive_expect_tx(&s_bar[st], 48u * (uint32_t)min(NT, start));
            if (nvalid) {
                nid = a.point_list[range.x + pos];
                tma_load_48(&s_raw[st][3 * tid], a.recs + 3 * (size_t)nid, &s_bar[st]);
            }
        } else if (nvalid) {
            nid = a.point_list[range.x + pos];
            const float4* r = a.recs + 3 * (size_t)nid;
            n0 = __ldg(r); n1 = __ldg(r + 1); n2 = __ldg(r + 2);
        }
    };
    if (tile_last > 0) fetch((int)tile_last, 0);
    float* const grad_base = reinterpret_cast<float*>(a.grad_recs);

    int round = 0;
    for (int start = (int)tile_last; start > 0; start -= NT, round++) {
        const int pos = start - 1 - tid;
        cur_id = nid;
        if (TMA) {
            const int st = round & 1;
            mbar_wait(&s_bar[st], (uint32_t)(round >> 1) & 1u);
            nvalid = pos >= 0;
            if (nvalid) { n0 = s_raw[st][3 * tid]; n1 = s_raw[st][3 * tid + 1]; n2 = s_raw[st][3 * tid + 2]; }
        }
        const float4 q0 = n0, q1 = n1, q2 = n2;
        const uint32_t id = cur_id;
        const bool keep = nvalid && tile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q2.y, tx0, ty0, tx1, ty1);
        if (start - NT > 0) fetch(start - NT, round + 1);
        int n;
        const int slot = block_compact<NT>(keep, s_wcount, n);   // barrier: previous batch fully consumed
        if (keep) {
            s_q0[slot] = q0;
            s_q1[slot] = make_float4(q1.x, q1.y, q2.y, __uint_as_float((uint32_t)pos));
            s_q2[slot] = make_float4(q1.z, q1.w, q2.x, __uint_as_float(id));
        }
        __syncthreads();

        for (int j = 0; j < n; j++) {
            const float4 g0 = s_q0[j];
            const float4 g1 = s_q1[j];
            const uint32_t pos_j = __float_as_uint(g1.w);
            bool act[PPT];
            float dx[PPT], dy[PPT], G[PPT], alpha[PPT];
            bool any_lane = false;
#pragma unroll
            for (int s = 0; s < PPT; s++) {
                act[s] = false;
                if (pos_j < last_contributor[s]) {
                    dx[s] = g0.x - pxf[s]; dy[s] = g0.y - pyf[s];
                    const float power = blend_power_exact(dx[s], dy[s], g0.z, g0.w, g1.x);
                    if (!(power > 0.0f || power < g1.z)) {
                        G[s] = expf(power);
                        alpha[s] = fminf(0.99f, g1.y * G[s]);
                        act[s] = !(alpha[s] < 1.0f / 255.0f);
                    }
                }
                any_lane = any_lane || act[s];
            }
            if (!__any_sync(FULL, any_lane)) continue;
            const float4 col = s_q2[j];
            float v[9];
#pragma unroll
            for (int k = 0; k < 9; k++) v[k] = 0.0f;
#pragma unroll
            for (int s = 0; s < PPT; s++) {
                if (act[s]) {
                    const float one_m_alpha = 1.0f - alpha[s];
                    const float inv = rcp_nr(one_m_alpha);
                    T[s] = T[s] * inv;
                    const float dchannel_dcolor = alpha[s] * T[s];
                    const float la = last_alpha[s], one_m_la = 1.0f - la;
                    acc0[s] = la * lc0[s] + one_m_la * acc0[s]; lc0[s] = col.x;
                    acc1[s] = la * lc1[s] + one_m_la * acc1[s]; lc1[s] = col.y;
                    acc2[s] = la * lc2[s] + one_m_la * acc2[s]; lc2[s] = col.z;
                    float dL_dalpha = ((col.x - acc0[s]) * dpx0[s] + (col.y - acc1[s]) * dpx1[s] +
                                       (col.z - acc2[s]) * dpx2[s]) * T[s];
                    last_alpha[s] = alpha[s];
                    if (has_bg) dL_dalpha += (-T_final[s] * inv) * bg_dot[s];
                    const float dL_dG = g1.y * dL_dalpha;
                    const float gdx = G[s] * dx[s], gdy = G[s] * dy[s];
                    const float dG_ddelx = -gdx * g0.z - gdy * g0.w;
                    const float dG_ddely = -gdy * g1.x - gdx * g0.w;
                    const float h = -0.5f * dL_dG;
                    v[0] += dL_dG * dG_ddelx;
                    v[1] += dL_dG * dG_ddely;
                    v[2] += h * gdx * dx[s];
                    v[3] += h * gdx * dy[s];
                    v[4] += h * gdy * dy[s];
                    v[5] += G[s] * dL_dalpha;
                    v[6] += dchannel_dcolor * dpx0[s];
                    v[7] += dchannel_dcolor * dpx1[s];
                    v[8] += dchannel_dcolor * dpx2[s];
                }
            }
            warp_reduce9(v, lane);
            // Native float reductions straight to the per-Gaussian gradient record.
            float* dst = grad_base + 12 * (size_t)__float_as_uint(col.w);
            if ((lane & 3) == 0) {
                const int k = lane >> 2;
                float val = v[0];
                if (k == 0) val *= ddelx_dx;
                if (k == 1) val *= ddely_dy;
                atomicAdd(dst + k, val);
            } else if (lane == 1) {
                atomicAdd(dst + 8, v[8]);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Debug: replay the forward loop and count what the blend kernels have to do.
// out[0] staged list entries   out[1] survivors of the tile cull
// out[2] (warp, survivor) iterations   out[3] ... with >= 1 lane passing the power/cut test
// out[4] ... with >= 1 lane blending   out[5] (pixel, entry) pairs evaluated
// out[6] (pixel, entry) pairs blended  out[7] list entries in all tile ranges
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLK) blend_stats_kernel(BlendFwdArgs a, unsigned long long* out) {
    __shared__ float4 s_q0[BLK];
    __shared__ float4 s_q1[BLK];
    __shared__ uint32_t s_wcount[BLK / 32];
    __shared__ unsigned long long s_cnt[8];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 8) s_cnt[tid] = 0;
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
    float T = 1.0f;
    unsigned long long c_staged = 0, c_surv = 0, c_wit = 0, c_wpow = 0, c_wblend = 0, c_peval = 0, c_pblend = 0;
    for (int i = 0; i < rounds; i++) {
        if (__syncthreads_count(done) == BLK) break;
        const int pos = i * BLK + tid;
        bool keep = false;
        float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0;
        if (pos < todo) {
            const uint32_t id = a.point_list[range.x + pos];
            const float4* r = a.recs + 3 * (size_t)id;
            q0 = r[0]; q1 = r[1]; q2 = r[2];
            keep = tile_may_contribute(q0.x, q0.y, q0.z, q0.w, q1.x, q2.y, tx0, ty0, tx1, ty1);
            c_staged++;
        }
        int n;
        const int slot = block_compact(keep, s_wcount, n);
        if (keep) { s_q0[slot] = q0; s_q1[slot] = make_float4(q1.x, q1.y, q2.y, 0.f); c_surv++; }
        __syncthreads();
        for (int j = 0; j < n; j++) {
            if (__all_sync(FULL, done)) break;
            if (lane == 0) c_wit++;
            const float4 g0 = s_q0[j];
            const float4 g1 = s_q1[j];
            bool pw = false, bl = false;
            if (!done) {
                c_peval++;
                const float dx = g0.x - pxf, dy = g0.y - pyf;
                const float power = blend_power_exact(dx, dy, g0.z, g0.w, g1.x);
                if (!(power > 0.0f || power < g1.z)) {
                    pw = true;
                    const float alpha = fminf(0.99f, g1.y * expf(power));
                    if (!(alpha < 1.0f / 255.0f)) {
                        const float test_T = T * (1.0f - alpha);
                        if (test_T < 0.0001f) done = true;
                        else { T = test_T; bl = true; c_pblend++; }
                    }
                }
            }
            const bool anyp = __any_sync(FULL, pw), anyb = __any_sync(FULL, bl);
            if (lane == 0) { c_wpow += anyp; c_wblend += anyb; }
        }
    }
    atomicAdd(&s_cnt[0], c_staged); atomicAdd(&s_cnt[1], c_surv); atomicAdd(&s_cnt[2], c_wit);
    atomicAdd(&s_cnt[3], c_wpow); atomicAdd(&s_cnt[4], c_wblend); atomicAdd(&s_cnt[5], c_peval);
    atomicAdd(&s_cnt[6], c_pblend);
    __syncthreads();
    if (tid < 7) atomicAdd(&out[tid], s_cnt[tid]);
    if (tid == 7) atomicAdd(&out[7], (unsigned long long)todo);
}

}  // namespace

int gsr_launch_blend_stats(const BlendFwdArgs& a, unsigned long long* out, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    blend_stats_kernel<<<grid, BLK, 0, stream>>>(a, out);
    GSR_CHECK_LAUNCH();
    return 0;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Variant selection: pixels per thread of the blend kernels (GSR_FWD_PPT / GSR_BWD_PPT =
// 1, 2 or 4) and TMA bulk-copy staging of the splat records (GSR_BLEND_TMA = 0/1).
int gsr_launch_blend_fwd(const BlendFwdArgs& a, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    static const int ver = env_int("GSR_BLEND_FWD_V", 2);   // 2: blend_v2.cu (warp regions + packed FP32x2), 1: this file
    if (ver == 2) return gsr_launch_blend_fwd_v2(a, stream);
    static const int ppt = env_int("GSR_FWD_PPT", 2);
    static const int tma = env_int("GSR_BLEND_TMA", 0);   // measured 4% slower than the register-staged gather at C2
    { GsrProfScope prof_("blend_fwd", stream);
    if (tma) {
        if (ppt == 4) blend_fwd_kernel<4, true><<<grid, BLK / 4, 0, stream>>>(a);
        else if (ppt == 2) blend_fwd_kernel<2, true><<<grid, BLK / 2, 0, stream>>>(a);
        else blend_fwd_kernel<1, true><<<grid, BLK, 0, stream>>>(a);
    } else {
        if (ppt == 4) blend_fwd_kernel<4, false><<<grid, BLK / 4, 0, stream>>>(a);
        else if (ppt == 2) blend_fwd_kernel<2, false><<<grid, BLK / 2, 0, stream>>>(a);
        else blend_fwd_kernel<1, false><<<grid, BLK, 0, stream>>>(a);
    } }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_blend_bwd_writes_moments() {
    static const int ver = env_int("GSR_BLEND_BWD_V", 2);
    return ver == 2 ? 1 : 0;
}

int gsr_launch_blend_bwd(const BlendBwdArgs& a, cudaStream_t stream) {
    dim3 grid(a.grid_x, a.grid_y, 1);
    static const int ver = env_int("GSR_BLEND_BWD_V", 2);   // 2: blend_v2.cu, 1: this file
    if (ver == 2) return gsr_launch_blend_bwd_v2(a, stream);
    static const int ppt = env_int("GSR_BWD_PPT", 4);
    static const int tma = env_int("GSR_BLEND_TMA", 0);   // measured 4% slower than the register-staged gather at C2
    { GsrProfScope prof_("blend_bwd", stream);
    if (tma) {
        if (ppt == 4) blend_bwd_kernel<4, 8, true><<<grid, BLK / 4, 0, stream>>>(a);
        else if (ppt == 2) blend_bwd_kernel<2, 5, true><<<grid, BLK / 2, 0, stream>>>(a);
        else blend_bwd_kernel<1, 4, true><<<grid, BLK, 0, stream>>>(a);
    } else {
        if (ppt == 4) blend_bwd_kernel<4, 8, false><<<grid, BLK / 4, 0, stream>>>(a);
        else if (ppt == 2) blend_bwd_kernel<2, 5, false><<<grid, BLK / 2, 0, stream>>>(a);
        else blend_bwd_kernel<1, 4, false><<<grid, BLK, 0, stream>>>(a);
    } }
    GSR_CHECK_LAUNCH();
    return 0;
}
