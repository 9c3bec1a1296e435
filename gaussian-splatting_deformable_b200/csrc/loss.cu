// Training loss of the reference, fused: (1 - lambda) * L1 + lambda * (1 - SSIM)  (train.py:323,529;
// utils/loss_utils.py:17-64), forward and backward in two kernels instead of ~40 torch launches
// (5 grouped conv2d, their elementwise glue, means, and the autograd mirror of all of it).
//
//   ssim_map = (2 mu1 mu2 + C1)(2 s12 + C2) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2))
//   mu = w * x,  s1 = w * x1^2 - mu1^2,  s2 = w * x2^2 - mu2^2,  s12 = w * x1 x2 - mu1 mu2
// with w the 11x11 Gaussian window (sigma 1.5) applied with ZERO padding, per channel.  The window is an
// outer product (loss_utils.py:27-31), so both kernels run it as two 11-tap passes through shared memory.
//
// Forward (ssim_l1_fwd_kernel): a CTA owns a 32x32 output tile of one channel, stages the 42x42 halo of both
// images, runs the horizontal pass of the five moments into shared memory and the vertical pass into registers -
// each thread filters 4 consecutive outputs from a 14-value sliding window in registers, which is what keeps the
// kernel off the shared-memory bandwidth limit - then evaluates the map and its three partial derivatives
//   dS/dmu1 (conv outputs P = w*x1^2, Q = w*x1x2 held fixed), dS/dP, dS/dQ
// which are written out (12 B/pixel/channel) for the backward; |x1 - x2| and the map are block-reduced and
// leave as two atomics; the last CTA turns the two sums into the loss.
// Backward (ssim_l1_bwd_kernel): dL/dx1 = gS (w * dS/dmu1 + 2 x1 (w * dS/dP) + x2 (w * dS/dQ)) + gL sign(x1 - x2),
// gS = -lambda / N, gL = (1 - lambda) / N, times the upstream gradient (a device scalar) - the adjoint of a
// zero-padded convolution with a symmetric window is the same convolution.
#include "f32x2.cuh"
#include "kernels.cuh"

namespace {

constexpr int TX = 32, TY = 32, HALO = 5, WIN = 11, NT = 256;
constexpr int SX = TX + 2 * HALO, SY = TY + 2 * HALO;       // 42 x 42 staged pixels
constexpr int SEG = 4;                                       // consecutive outputs per thread along the filtered axis:
                                                             // 14 shared loads feed 4 outputs (a sliding window in registers)

struct SsimWindow { float w[WIN]; };

__device__ __forceinline__ float block_sum(float v, float* s_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = v;
    __syncthreads();
    float t = 0.0f;
    if (warp == 0) {
        t = lane < NT / 32 ? s_part[lane] : 0.0f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;        // valid in thread 0
}

__global__ void __launch_bounds__(NT) ssim_l1_fwd_kernel(const float* __restrict__ img, const float* __restrict__ gt, int H, int W,
                                                         SsimWindow win, float lambda, float* __restrict__ dmaps,
                                                         float* __restrict__ sums /*[2] + counter + out[3]*/) {
    // (x1, x2) and the moment pairs (w*x1, w*x2), (w*x1^2, w*x2^2) live interleaved so that one LDS.64 feeds one
    // packed FFMA2 (f32x2.cuh); w*x1x2 stays scalar.  5 FMA + 3 MUL per tap become 3 packed + 2 scalar.
    __shared__ float2 s_x[SY][SX + 1];
    __shared__ float2 s_hm[SY][TX + 1];            // horizontal pass of (x1, x2)
    __shared__ float2 s_hs[SY][TX + 1];            //                    (x1^2, x2^2)
    __shared__ float s_hc[SY][TX + 1];             //                    x1 x2
    __shared__ float s_part[NT / 32];
    __shared__ bool s_last;
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const size_t plane = (size_t)H * W;
    const float* p1 = img + c * plane;
    const float* p2 = gt + c * plane;
    const int tid = threadIdx.x;
    for (int i = tid; i < SX * SY; i += NT) {
        const int sx = i % SX, sy = i / SX;
        const int gx = x0 + sx - HALO, gy = y0 + sy - HALO;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        s_x[sy][sx] = in ? make_float2(p1[(size_t)gy * W + gx], p2[(size_t)gy * W + gx]) : make_float2(0.0f, 0.0f);
    }
    __syncthreads();
    // horizontal pass: task = (row, segment of SEG output columns); consecutive lanes take consecutive ROWS
    // (row stride 43 float2 -> conflict-free 64-bit shared accesses)
    for (int i = tid; i < SY * (TX / SEG); i += NT) {
        const int sy = i % SY, seg = i / SY;
        const int ox = seg * SEG;
        f32x2 u[SEG + WIN - 1];
#pragma unroll
        for (int k = 0; k < SEG + WIN - 1; k++) { const float2 t = s_x[sy][ox + k]; u[k] = pk(t.x, t.y); }
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            f32x2 m = pk1(0.0f), q = pk1(0.0f);
            float ab = 0.0f;
#pragma unroll
            for (int k = 0; k < WIN; k++) {
                const f32x2 w2 = pk1(win.w[k]);
                const f32x2 uv = u[o + k];
                float a1, a2;
                upk(uv, a1, a2);
                m = fma2(w2, uv, m);
                q = fma2(w2, mul2(uv, uv), q);
                ab = fmaf(win.w[k], a1 * a2, ab);
            }
            float m1, m2, q1, q2;
            upk(m, m1, m2); upk(q, q1, q2);
            s_hm[sy][ox + o] = make_float2(m1, m2);
            s_hs[sy][ox + o] = make_float2(q1, q2);
            s_hc[sy][ox + o] = ab;
        }
    }
    __syncthreads();
    // vertical pass: thread = (column, segment of SEG output rows)
    const int tx = tid % TX, oy0 = (tid / TX) * SEG;
    float acc[5][SEG];
    {
        f32x2 cm[SEG + WIN - 1], cs[SEG + WIN - 1];
        float cc[SEG + WIN - 1];
#pragma unroll
        for (int k = 0; k < SEG + WIN - 1; k++) {
            const float2 a = s_hm[oy0 + k][tx], b = s_hs[oy0 + k][tx];
            cm[k] = pk(a.x, a.y); cs[k] = pk(b.x, b.y); cc[k] = s_hc[oy0 + k][tx];
        }
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            f32x2 m = pk1(0.0f), q = pk1(0.0f);
            float ab = 0.0f;
#pragma unroll
            for (int k = 0; k < WIN; k++) {
                const f32x2 w2 = pk1(win.w[k]);
                m = fma2(w2, cm[o + k], m);
                q = fma2(w2, cs[o + k], q);
                ab = fmaf(win.w[k], cc[o + k], ab);
            }
            upk(m, acc[0][o], acc[1][o]);
            upk(q, acc[2][o], acc[3][o]);
            acc[4][o] = ab;
        }
    }
    float l1_acc = 0.0f, ssim_acc = 0.0f;
#pragma unroll
    for (int o = 0; o < SEG; o++) {
        const int oy = oy0 + o;
        const int gx = x0 + tx, gy = y0 + oy;
        if (gx < W && gy < H) {
            const float mu1 = acc[0][o], mu2 = acc[1][o], e11 = acc[2][o], e22 = acc[3][o], e12 = acc[4][o];
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu12;
            const float A1 = 2.0f * mu12 + C1, A2 = 2.0f * s12 + C2;
            const float B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
            const float inv = 1.0f / (B1 * B2);
            const float S = A1 * A2 * inv;
            // partial derivatives of the map w.r.t. mu1, P = w*x1^2, Q = w*x1x2 (mu2, w*x2^2 do not depend on x1)
            const float dS_dmu1 = 2.0f * mu2 * (A2 - A1) * inv - S * 2.0f * mu1 * (1.0f / B1 - 1.0f / B2);
            const float dS_dP = -S / B2;
            const float dS_dQ = 2.0f * A1 * inv;
            const size_t o_ = c * plane + (size_t)gy * W + gx;
            const size_t CHW = 3 * plane;
            dmaps[o_] = dS_dmu1; dmaps[CHW + o_] = dS_dP; dmaps[2 * CHW + o_] = dS_dQ;
            ssim_acc += S;
            const float2 xc = s_x[oy + HALO][tx + HALO];
            l1_acc += fabsf(xc.x - xc.y);
        }
    }
    const float bl1 = block_sum(l1_acc, s_part);
    const float bss = block_sum(ssim_acc, s_part);
    if (tid == 0) {
        atomicAdd(&sums[0], bl1);
        atomicAdd(&sums[1], bss);
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(&sums[2]), 1u) + 1u;
        s_last = done == total;
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        const float n = (float)(3.0 * (double)plane);
        const float l1 = atomicAdd(&sums[0], 0.0f) / n, ss = atomicAdd(&sums[1], 0.0f) / n;
        sums[3] = (1.0f - lambda) * l1 + lambda * (1.0f - ss);       // train.py:529
        sums[4] = l1;
        sums[5] = ss;
    }
}

__global__ void __launch_bounds__(NT) ssim_l1_bwd_kernel(const float* __restrict__ img, const float* __restrict__ gt, int H, int W,
                                                         SsimWindow win, float lambda, const float* __restrict__ dmaps,
                                                         const float* __restrict__ upstream, float* __restrict__ dL_dimg) {
    __shared__ float2 s_d01[SY][SX + 1];           // (dS/dmu1, dS/dP) interleaved: one LDS.64 per packed FFMA2
    __shared__ float s_d2[SY][SX + 1];             // dS/dQ
    __shared__ float2 s_h01[SY][TX + 1];
    __shared__ float s_h2[SY][TX + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const size_t plane = (size_t)H * W, CHW = 3 * plane;
    const int tid = threadIdx.x;
    for (int i = tid; i < SX * SY; i += NT) {
        const int sx = i % SX, sy = i / SX;
        const int gx = x0 + sx - HALO, gy = y0 + sy - HALO;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const size_t o = c * plane + (size_t)gy * W + gx;
        s_d01[sy][sx] = in ? make_float2(dmaps[o], dmaps[CHW + o]) : make_float2(0.0f, 0.0f);
        s_d2[sy][sx] = in ? dmaps[2 * CHW + o] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < SY * (TX / SEG); i += NT) {
        const int sy = i % SY, seg = i / SY;
        const int ox = seg * SEG;
        f32x2 u[SEG + WIN - 1];
        float v[SEG + WIN - 1];
#pragma unroll
        for (int k = 0; k < SEG + WIN - 1; k++) { const float2 t = s_d01[sy][ox + k]; u[k] = pk(t.x, t.y); v[k] = s_d2[sy][ox + k]; }
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            f32x2 a = pk1(0.0f);
            float d = 0.0f;
#pragma unroll
            for (int k = 0; k < WIN; k++) { a = fma2(pk1(win.w[k]), u[o + k], a); d = fmaf(win.w[k], v[o + k], d); }
            float a0, a1;
            upk(a, a0, a1);
            s_h01[sy][ox + o] = make_float2(a0, a1);
            s_h2[sy][ox + o] = d;
        }
    }
    __syncthreads();
    const float up = upstream ? upstream[0] : 1.0f;
    const float n = (float)(3.0 * (double)plane);
    const float gS = -lambda / n * up, gL = (1.0f - lambda) / n * up;
    const int tx = tid % TX, oy0 = (tid / TX) * SEG;
    float acc[3][SEG];
    {
        f32x2 c01[SEG + WIN - 1];
        float c2[SEG + WIN - 1];
#pragma unroll
        for (int k = 0; k < SEG + WIN - 1; k++) { const float2 t = s_h01[oy0 + k][tx]; c01[k] = pk(t.x, t.y); c2[k] = s_h2[oy0 + k][tx]; }
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            f32x2 a = pk1(0.0f);
            float d = 0.0f;
#pragma unroll
            for (int k = 0; k < WIN; k++) { a = fma2(pk1(win.w[k]), c01[o + k], a); d = fmaf(win.w[k], c2[o + k], d); }
            upk(a, acc[0][o], acc[1][o]);
            acc[2][o] = d;
        }
    }
#pragma unroll
    for (int o = 0; o < SEG; o++) {
        const int gx = x0 + tx, gy = y0 + oy0 + o;
        if (gx < W && gy < H) {
            const size_t o_ = c * plane + (size_t)gy * W + gx;
            const float x1 = img[o_], x2 = gt[o_];
            const float diff = x1 - x2;
            const float sgn = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);      // torch.abs backward: sign, 0 at 0
            dL_dimg[o_] = gS * (acc[0][o] + 2.0f * x1 * acc[1][o] + x2 * acc[2][o]) + gL * sgn;
        }
    }
}

}  // namespace

// sums: device float[8], zeroed here: [0] sum |x1-x2|, [1] sum ssim_map, [2] CTA counter, [3] loss, [4] L1 mean, [5] SSIM mean
int gsr_launch_ssim_l1_fwd(const float* img, const float* gt, int H, int W, const float* window11, float lambda,
                           float* dmaps, float* sums, cudaStream_t stream) {
    SsimWindow win;
    for (int k = 0; k < WIN; k++) win.w[k] = window11[k];
    GSR_CHECK(cudaMemsetAsync(sums, 0, 8 * sizeof(float), stream));
    dim3 grid(gsr_div_up(W, TX), gsr_div_up(H, TY), 3);
    { GsrProfScope prof_("ssim_l1_fwd", stream);
    ssim_l1_fwd_kernel<<<grid, NT, 0, stream>>>(img, gt, H, W, win, lambda, dmaps, sums); }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_ssim_l1_bwd(const float* img, const float* gt, int H, int W, const float* window11, float lambda,
                           const float* dmaps, const float* upstream, float* dL_dimg, cudaStream_t stream) {
    SsimWindow win;
    for (int k = 0; k < WIN; k++) win.w[k] = window11[k];
    dim3 grid(gsr_div_up(W, TX), gsr_div_up(H, TY), 3);
    { GsrProfScope prof_("ssim_l1_bwd", stream);
    ssim_l1_bwd_kernel<<<grid, NT, 0, stream>>>(img, gt, H, W, win, lambda, dmaps, upstream, dL_dimg); }
    GSR_CHECK_LAUNCH();
    return 0;
}
