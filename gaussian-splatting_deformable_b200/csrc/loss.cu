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
    __shared__ float s_x1[SY][SX + 1];
    __shared__ float s_x2[SY][SX + 1];
    __shared__ float s_h[5][SY][TX + 1];           // horizontal pass of x1, x2, x1^2, x2^2, x1 x2
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
        s_x1[sy][sx] = in ? p1[(size_t)gy * W + gx] : 0.0f;
        s_x2[sy][sx] = in ? p2[(size_t)gy * W + gx] : 0.0f;
    }
    __syncthreads();
    // horizontal pass: task = (row, segment of SEG output columns)
    for (int i = tid; i < SY * (TX / SEG); i += NT) {
        const int seg = i % (TX / SEG), sy = i / (TX / SEG);
        const int ox = seg * SEG;
        float u[SEG + WIN - 1], v[SEG + WIN - 1];
#pragma unroll
        for (int k = 0; k < SEG + WIN - 1; k++) { u[k] = s_x1[sy][ox + k]; v[k] = s_x2[sy][ox + k]; }
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            float a = 0, b = 0, aa = 0, bb = 0, ab = 0;
#pragma unroll
            for (int k = 0; k < WIN; k++) {
                const float w = win.w[k], uu = u[o + k], vv = v[o + k];
                a = fmaf(w, uu, a); b = fmaf(w, vv, b);
                aa = fmaf(w, uu * uu, aa); bb = fmaf(w, vv * vv, bb); ab = fmaf(w, uu * vv, ab);
            }
            s_h[0][sy][ox + o] = a; s_h[1][sy][ox + o] = b; s_h[2][sy][ox + o] = aa; s_h[3][sy][ox + o] = bb;
            s_h[4][sy][ox + o] = ab;
        }
    }
    __syncthreads();
    // vertical pass: thread = (column, segment of SEG output rows)
    const int tx = tid % TX, oy0 = (tid / TX) * SEG;
    float acc[5][SEG];
#pragma unroll
    for (int q = 0; q < 5; q++) {
        float col[SEG + WIN - 1];
#pragma unroll
        for (int k = 0; k < SEG + WIN - 1; k++) col[k] = s_h[q][oy0 + k][tx];
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            float a = 0;
#pragma unroll
            for (int k = 0; k < WIN; k++) a = fmaf(win.w[k], col[o + k], a);
            acc[q][o] = a;
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
            l1_acc += fabsf(s_x1[oy + HALO][tx + HALO] - s_x2[oy + HALO][tx + HALO]);
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
    __shared__ float s_d[3][SY][SX + 1];
    __shared__ float s_h[3][SY][TX + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const size_t plane = (size_t)H * W, CHW = 3 * plane;
    const int tid = threadIdx.x;
    for (int i = tid; i < SX * SY; i += NT) {
        const int sx = i % SX, sy = i / SX;
        const int gx = x0 + sx - HALO, gy = y0 + sy - HALO;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const size_t o = c * plane + (size_t)gy * W + gx;
#pragma unroll
        for (int m = 0; m < 3; m++) s_d[m][sy][sx] = in ? dmaps[m * CHW + o] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < SY * (TX / SEG); i += NT) {
        const int seg = i % (TX / SEG), sy = i / (TX / SEG);
        const int ox = seg * SEG;
#pragma unroll
        for (int m = 0; m < 3; m++) {
            float u[SEG + WIN - 1];
#pragma unroll
            for (int k = 0; k < SEG + WIN - 1; k++) u[k] = s_d[m][sy][ox + k];
#pragma unroll
            for (int o = 0; o < SEG; o++) {
                float a = 0;
#pragma unroll
                for (int k = 0; k < WIN; k++) a = fmaf(win.w[k], u[o + k], a);
                s_h[m][sy][ox + o] = a;
            }
        }
    }
    __syncthreads();
    const float up = upstream ? upstream[0] : 1.0f;
    const float n = (float)(3.0 * (double)plane);
    const float gS = -lambda / n * up, gL = (1.0f - lambda) / n * up;
    const int tx = tid % TX, oy0 = (tid / TX) * SEG;
    float acc[3][SEG];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        float col[SEG + WIN - 1];
#pragma unroll
        for (int k = 0; k < SEG + WIN - 1; k++) col[k] = s_h[m][oy0 + k][tx];
#pragma unroll
        for (int o = 0; o < SEG; o++) {
            float a = 0;
#pragma unroll
            for (int k = 0; k < WIN; k++) a = fmaf(win.w[k], col[o + k], a);
            acc[m][o] = a;
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
