// Microbenchmark: issue cost of packed FP32 (FFMA2/FMUL2/FADD2, sm_100) vs scalar FFMA,
// alone and mixed with integer instructions.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__device__ __forceinline__ float fma1(float a,float b,float c){float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b){unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;}

template <int MODE>  // 0: 16 FFMA  1: 8 FFMA2  2: 16 FFMA + 16 IADD  3: 8 FFMA2 + 16 IADD  4: 16 IADD
__global__ void __launch_bounds__(256) bench(float* out, int iters, long long* clk) {
    float a[16]; u64 p[8]; unsigned q[16];
    const float s = 1.0000001f, t = 1e-9f;
    for (int i = 0; i < 16; i++) { a[i] = threadIdx.x * 1e-3f + i; q[i] = threadIdx.x + i; }
    for (int i = 0; i < 8; i++) p[i] = ((u64)__float_as_uint(a[2*i]) << 32) | __float_as_uint(a[2*i+1]);
    const u64 s2 = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s), t2 = ((u64)__float_as_uint(t) << 32) | __float_as_uint(t);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fma1(a[i], s, t);
        }
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], s2, t2);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) q[i] = iadd(q[i], 3u);
        }
    }
    long long t1 = clock64();
    float r = 0; for (int i = 0; i < 16; i++) r += a[i] + (float)q[i];
    for (int i = 0; i < 8; i++) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * 256 + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
// numeric check: does mul.rn.f32x2 followed by add.rn.f32x2 round twice (IEEE) or once (fused)?
__global__ void fuse_check(float* out) {
    float x = 1.0f + 1.1920929e-7f, y = 1.0f + 1.1920929e-7f, z = -1.0f;
    u64 X = ((u64)__float_as_uint(x) << 32) | __float_as_uint(x), Y = X, Z = ((u64)__float_as_uint(z) << 32) | __float_as_uint(z);
    X += out[8] > 1e30f; 
    u64 m, r;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(X), "l"(Y));
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(m), "l"(Z));
    out[0] = __uint_as_float((unsigned)r);
    float xs = x + (out[8] > 1e30f ? 1.f : 0.f);
    out[1] = __fadd_rn(__fmul_rn(xs, y), z);
    out[2] = __fmaf_rn(xs, y, z);
}
template <int MODE> void run(const char* name, float* d, long long* dclk) {
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<blocks, 256>>>(d, iters, dclk);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, 256>>>(d, iters, dclk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk; cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
    // each SM sub-partition runs 8 CTAs x 8 warps / 4 = 16 warps
    printf("%-22s %.3f ms   CTA0 clocks/iter %.2f  (16 warps per scheduler -> clocks per warp-iter per scheduler %.2f)\n", name, ms, (double)clk / iters, (double)clk / iters / 16.0);
}
int main() {
    float* d; long long* dclk; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&dclk, 8);
    cudaMemset(d, 0, 64);
    run<0>("16 FFMA", d, dclk); run<1>("8 FFMA2", d, dclk); run<2>("16 FFMA + 16 IADD", d, dclk);
    run<3>("8 FFMA2 + 16 IADD", d, dclk); run<4>("16 IADD", d, dclk);
    fuse_check<<<1, 1>>>(d); float h[3]; cudaMemcpy(h, d, 12, cudaMemcpyDeviceToHost);
    printf("mul2+add2 = %.9g   scalar mul,add = %.9g   fma = %.9g  -> %s\n", h[0], h[1], h[2], h[0] == h[1] ? "NOT fused (IEEE)" : "FUSED by ptxas");
    return 0;
}
