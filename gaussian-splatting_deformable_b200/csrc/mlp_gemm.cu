// Tensor-core GEMM of the deformation network: C = epilogue(A . B^T) with fp32-grade accuracy on
// tcgen05 (5th-generation tensor cores, accumulators in TMEM), operands staged by TMA.
//
// The reference evaluates DirectTemporalNeRF (scene/gaussian_model.py:242-316: 84 -> 8 x 256 ReLU, skip after layer
// 4, heads 3/3/4/48) with torch fp32 linears, i.e. cuBLAS SGEMM on the FP32 pipe.  Tensor cores take TF32 (10-bit
// mantissa) at best, which alone misses the fp32 result by ~1e-3; so every operand travels as two planes
//      x = hi + lo,   hi = x rounded to nearest TF32 value (13 low mantissa bits zero),   lo = x - hi (exact)
// and every product is   A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo   (three kind::tf32 MMAs into one fp32 TMEM
// accumulator; the dropped lo.lo term and the TF32 rounding of lo are ~2^-22 relative).  Measured against an fp64
// evaluation the result is as close as cuBLAS SGEMM's own (tests/test_gpu_mlp.py).
//
// One persistent CTA per SM, warp-specialised (320 threads):
//   warp 0       TMA producer: per k-block of 16 fp32 (one 64-byte swizzle row) loads the A tile [128 x 16] and the B
//                tile [BN x 16] into a 4-stage shared ring (cp.async.bulk.tensor, SWIZZLE_64B, mbarrier tx).  An operand
//                given as ONE fp32 plane (activations: P-sized, they should cross HBM once) lands in the stage's hi slot;
//                an operand given as pre-split (hi, lo) planes (weights: tiny, L2-resident) lands in both slots
//   warps 2,3,8,9  converters: split a single-plane tile in shared memory, in place (hi over the raw values, lo into the
//                stage's lo slot; elementwise, so the hardware swizzle pattern is irrelevant), then
//                fence.proxy.async and hand the stage to the MMA warp
//   warp 1       MMA issuer: one elected lane issues 2 k-steps x 3 tcgen05.mma (M 128, N BN, K 8) per k-block and
//                commits the stage back to the producer; after the last k-block commits the accumulator to the epilogue
//   warp 2       also the TMEM allocator (512 columns = two accumulator stages: tile i's epilogue overlaps tile i+1's MMAs)
//   warps 4-7    epilogue: tcgen05.ld (lane = row), bias / ReLU / ReLU-mask, 256-bit stores (one full DRAM sector per
//                thread-instruction) of the fp32 result (or of (hi, lo) planes), transposed copy for the weight-gradient
//                GEMM, atomic adds for split-K, column sums for bias gradients
// The same kernel runs all three GEMMs of a linear layer: forward (A = activations, B = W), input gradient
// (A = dZ, B = W^T) and weight gradient (A = dZ^T, B = X^T, split over the point dimension, atomic epilogue).
#include "kernels.cuh"
#include <cuda.h>
#include <string.h>

namespace {

constexpr int BM = 128;            // rows per tile = TMEM lanes
constexpr int BK = 16;             // fp32 per k-block = one 64-byte swizzle row
constexpr int UMMA_K = 8;          // K of one kind::tf32 MMA
constexpr int STAGES = 4;
constexpr int NUM_THREADS = 320;
constexpr int NUM_CONVERTERS = 128;
constexpr uint32_t TMEM_COLS = 512;

struct GemmParams {
    int M;                   // rows of A / C
    int N;                   // valid output columns (<= BN)
    int kblocks[2];          // k-blocks of A segment 0 and 1 (B's K runs over both, segment 0 first)
    int m_tiles, k_splits;   // tiles = m_tiles * k_splits; a tile covers k-blocks [ks * kb_per_split, ...)
    int kb_per_split;
    int mode;                // GSR_GEMM_* epilogue mode
    int convA, convB;        // 1: the operand arrives as one fp32 plane and is split in shared memory
    int mn_major;            // 1: A is [K x M] and B is [K x N] row-major (the reduction runs over ROWS: weight gradients)
    const float* bias;       // [N] or null
    const float* mask_src;   // [M x ld_mask] ReLU mask source (v *= mask_src > 0) or null
    int ld_mask;
    float* out_hi; float* out_lo; int ld_out;              // row-major planes [M x ld_out] (or the fp32 / atomic target in out_hi)
    float* outT_hi; float* outT_lo; long long ld_outT;     // transposed planes [N x ld_outT] or null
    float* colsum;           // [N] += column sums of the stored values (bias gradient) or null
    uint32_t* error_flag;    // set when a barrier wait times out (never in a healthy run)
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end as a reported error, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++) {
        if (spins > (1u << 22) || *abort_flag) { *abort_flag = 1; return false; }
    }
    return true;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major tile [rows x 16 fp32] written by TMA with SWIZZLE_64B: rows are 64
// bytes apart, groups of 8 rows (one 512-byte swizzle atom) are SBO = 512 bytes apart; LBO is unused for swizzled
// K-major layouts (1); descriptor version 1 (sm_100); layout type 4 = SWIZZLE_64B.  (cute/arch/mma_sm100_desc.hpp:
// canonical K-major layout B64 = Swizzle<2,4,3> o ((8,m),(T,2)) : ((4T,SBO),(1,T)), T = 4 fp32)
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// MN-major tile (the reduction index K runs over the ROWS of a row-major global tensor): TMA boxes of [16 rows x 32
// fp32], one box per 32 M/N-elements.  For 32-bit (TF32) operands the tensor core transposes only the layout
// SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem
// layout"), TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: Swizzle<2,5,2> = 32-byte chunks of a 128-byte row XORed with the
// row index mod 4.  Canonical form (mma_sm100_desc.hpp / mma_traits_sm100.hpp, 16-byte units) ((8,n),(4,k)) : ((1,LBO),(8,SBO)):
// 128 contiguous bytes along M/N, 4 K-rows 128 bytes apart (one 512-byte swizzle atom), the next 4 K-rows SBO = 512 bytes
// on, the next 32 M/N-elements LBO = 2048 bytes (one box) on.  Layout type 1.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// kind::tf32 instruction descriptor: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9, 10-12 = 2), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, bool mn_major = false) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// hi = x rounded TO NEAREST to TF32 precision (11 significant bits; ties away from zero): then |lo| = |x - hi| <= 2^-12 |x|
// carries at most 12 significant bits, of which the tensor core keeps 11, i.e. the split loses <= 2^-23 |x| - fp32's own
// half-ulp (truncating instead would leave |lo| <= 2^-11 |x| with 13 bits and lose 2^-21 |x|: measured 4x the error).
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

template <int BN>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 4;     // 8 KB
    static constexpr int B_BYTES = BN * BK * 4;     // 16 KB at BN = 256
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // slots: A hi | A lo | B hi | B lo
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = RING_BYTES;                      // full[STAGES], conv[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
    static constexpr int MISC_OFF = BAR_OFF + 8 * (3 * STAGES + 4);  // tmem base, abort flag
    static constexpr int COLSUM_OFF = MISC_OFF + 16;                 // float[BN]
    static constexpr int TOTAL = COLSUM_OFF + BN * 4 + 1024;         // + slack for the 1024-byte alignment of the ring
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mlp_gemm_kernel(const __grid_constant__ CUtensorMap mapA0_hi, const __grid_constant__ CUtensorMap mapA0_lo,
                const __grid_constant__ CUtensorMap mapA1_hi, const __grid_constant__ CUtensorMap mapA1_lo,
                const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo, GemmParams p) {
    using L = SmemLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* conv_bar = full_bar + STAGES;
    uint64_t* empty_bar = conv_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(smem + L::MISC_OFF);
    volatile uint32_t* abort_flag = reinterpret_cast<volatile uint32_t*>(smem + L::MISC_OFF + 4);
    float* s_colsum = reinterpret_cast<float*>(smem + L::COLSUM_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.k_splits;
    const int kb_total = p.kblocks[0] + p.kblocks[1];

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&conv_bar[s], NUM_CONVERTERS); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
        *abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < BN; i += NUM_THREADS) s_colsum[i] = 0.0f;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m_blk = t % p.m_tiles, ks = t / p.m_tiles;
                const int kb0 = ks * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; kb++) {
                    if (!mbar_wait(&empty_bar[stage], phase ^ 1, abort_flag)) break;
                    uint8_t* st = smem + stage * L::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], (p.convA ? 1 : 2) * L::A_BYTES + (p.convB ? 1 : 2) * L::B_BYTES);
                    if (p.mn_major) {
                        // one [16 x 32] box per 32 output rows / columns; coordinates (column, row = reduction index)
#pragma unroll
                        for (int j = 0; j < BM / 32; j++) tma_load_2d(st + j * 2048, &mapA0_hi, m_blk * BM + 32 * j, kb * BK, &full_bar[stage]);
#pragma unroll
                        for (int j = 0; j < BN / 32; j++) tma_load_2d(st + 2 * L::A_BYTES + j * 2048, &mapB_hi, 32 * j, kb * BK, &full_bar[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    const bool seg1 = kb >= p.kblocks[0];
                    const int ka = (seg1 ? kb - p.kblocks[0] : kb) * BK;
                    tma_load_2d(st, seg1 ? &mapA1_hi : &mapA0_hi, ka, m_blk * BM, &full_bar[stage]);
                    if (!p.convA) tma_load_2d(st + L::A_BYTES, seg1 ? &mapA1_lo : &mapA0_lo, ka, m_blk * BM, &full_bar[stage]);
                    tma_load_2d(st + 2 * L::A_BYTES, &mapB_hi, kb * BK, 0, &full_bar[stage]);
                    if (!p.convB) tma_load_2d(st + 2 * L::A_BYTES + L::B_BYTES, &mapB_lo, kb * BK, 0, &full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (*abort_flag) break;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(BM, BN, p.mn_major != 0);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int ks = t / p.m_tiles;
                const int kb0 = ks * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
                if (!mbar_wait(&acc_empty[acc], acc_phase ^ 1, abort_flag)) break;      // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; kb++) {
                    if (!mbar_wait(&conv_bar[stage], phase, abort_flag)) break;       // loaded AND split
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * L::STAGE_BYTES);
                    const bool mn = p.mn_major != 0;
                    const uint64_t a_hi = mn ? smem_desc_mn_sw128(st) : smem_desc_sw64(st);
                    const uint64_t a_lo = mn ? smem_desc_mn_sw128(st + L::A_BYTES) : smem_desc_sw64(st + L::A_BYTES);
                    const uint64_t b_hi = mn ? smem_desc_mn_sw128(st + 2 * L::A_BYTES) : smem_desc_sw64(st + 2 * L::A_BYTES);
                    const uint64_t b_lo = mn ? smem_desc_mn_sw128(st + 2 * L::A_BYTES + L::B_BYTES) : smem_desc_sw64(st + 2 * L::A_BYTES + L::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++) {
                        // K-major: +32 bytes along K inside the swizzle row; MN-major: the next 8 K-rows = +1024 bytes
                        const uint64_t adv = mn ? (uint64_t)((k * 1024) >> 4) : (uint64_t)((k * UMMA_K * 4) >> 4);
                        // small terms first: they accumulate in fp32 either way, order only matters for rounding
                        umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                        umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
                    }
                    umma_commit(&empty_bar[stage]);                                  // frees the stage when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[acc]);                                          // accumulator complete -> epilogue
                if (*abort_flag) break;
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp == 2 || warp == 3 || warp >= 8) {
        // ===================== converters: single-plane operand tiles -> (hi, lo) in shared memory =====================
        const int ct = (warp < 4 ? warp - 2 : warp - 6) * 32 + lane;        // 0 .. 127
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int ks = t / p.m_tiles;
            const int kb0 = ks * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
            bool ok = true;
            for (int kb = kb0; kb < kb1 && ok; kb++) {
                if (!mbar_wait(&full_bar[stage], phase, abort_flag)) { ok = false; break; }
                uint8_t* st = smem + stage * L::STAGE_BYTES;
                if (p.convA) {
                    float4* hi = reinterpret_cast<float4*>(st);
                    float4* lo = reinterpret_cast<float4*>(st + L::A_BYTES);
#pragma unroll
                    for (int i = ct; i < L::A_BYTES / 16; i += NUM_CONVERTERS) {
                        const float4 v = hi[i];
                        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                        hi[i] = h;
                        lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
                }
                if (p.convB) {
                    float4* hi = reinterpret_cast<float4*>(st + 2 * L::A_BYTES);
                    float4* lo = reinterpret_cast<float4*>(st + 2 * L::A_BYTES + L::B_BYTES);
#pragma unroll
                    for (int i = ct; i < L::B_BYTES / 16; i += NUM_CONVERTERS) {
                        const float4 v = hi[i];
                        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                        hi[i] = h;
                        lo[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                    }
                }
                fence_proxy_async_smem();            // generic-proxy writes -> visible to the tensor core's async-proxy reads
                mbar_arrive(&conv_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (!ok || *abort_flag) break;
        }
    } else if (warp >= 4) {
        // ===================== epilogue (warps 4-7: TMEM lanes 32 (warp % 4) ...) =====================
        const int wq = warp & 3;
        const int row_in_tile = wq * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int m_blk = t % p.m_tiles;
            const long long row = (long long)m_blk * BM + row_in_tile;
            const bool row_ok = row < p.M;
            if (!mbar_wait(&acc_full[acc], acc_phase, abort_flag)) break;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (c0 >= p.N) break;
                uint32_t raw[32];
                tmem_ld32(taddr + (uint32_t)c0, raw);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = __uint_as_float(raw[j]);
                if (p.mode == GSR_GEMM_ATOMIC) {
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (c0 + j < p.N) atomicAdd(p.out_hi + row * p.ld_out + c0 + j, v[j]);
                    }
                    continue;
                }
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] += (c0 + j < p.N) ? __ldg(p.bias + c0 + j) : 0.0f;
                }
                if (p.mode == GSR_GEMM_RELU_SPLIT) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.0f);
                }
                if (p.mask_src && row_ok) {
                    const float4* ms = reinterpret_cast<const float4*>(p.mask_src + row * p.ld_mask + c0);
#pragma unroll
                    for (int j4 = 0; j4 < 8; j4++) {
                        const float4 m4 = __ldg(ms + j4);
                        v[4 * j4 + 0] = m4.x > 0.0f ? v[4 * j4 + 0] : 0.0f;
                        v[4 * j4 + 1] = m4.y > 0.0f ? v[4 * j4 + 1] : 0.0f;
                        v[4 * j4 + 2] = m4.z > 0.0f ? v[4 * j4 + 2] : 0.0f;
                        v[4 * j4 + 3] = m4.w > 0.0f ? v[4 * j4 + 3] : 0.0f;
                    }
                }
                if (!row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = 0.0f;
                }
                // Row-major result for the next GEMM's A operand: one fp32 plane (split later, in the consumer's shared
                // memory) or, on request, (hi, lo) planes.  A lane owns 32 consecutive floats of its row = four full
                // 32-byte sectors: 256-bit stores.
                const bool full = (p.N - c0 >= 32) && ((p.ld_out & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 31) == 0) &&
                                  (!p.out_lo || (reinterpret_cast<uintptr_t>(p.out_lo) & 31) == 0);
                if (row_ok && p.out_hi) {
                    float* oh = p.out_hi + row * p.ld_out + c0;
                    if (!p.out_lo) {
                        if (full) {
#pragma unroll
                            for (int j8 = 0; j8 < 4; j8++) st256(oh + 8 * j8, v + 8 * j8);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (c0 + j < p.N) oh[j] = v[j];
                        }
                    } else {
                        float* ol = p.out_lo + row * p.ld_out + c0;
#pragma unroll
                        for (int j8 = 0; j8 < 4; j8++) {
                            float h[8], l[8];
#pragma unroll
                            for (int e = 0; e < 8; e++) { h[e] = tf32_hi(v[8 * j8 + e]); l[e] = v[8 * j8 + e] - h[e]; }
                            if (full) { st256(oh + 8 * j8, h); st256(ol + 8 * j8, l); }
                            else {
#pragma unroll
                                for (int e = 0; e < 8; e++)
                                    if (c0 + 8 * j8 + e < p.N) { oh[8 * j8 + e] = h[e]; ol[8 * j8 + e] = l[e]; }
                            }
                        }
                    }
                }
                // Transposed copy for the weight-gradient GEMM: for a fixed column the 32 lanes hold 32 consecutive
                // rows, so every store instruction writes one contiguous 128-byte line.
                if (row_ok && p.outT_hi) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (c0 + j < p.N) {
                            if (p.outT_lo) {
                                const float h = tf32_hi(v[j]);
                                p.outT_hi[(long long)(c0 + j) * p.ld_outT + row] = h;
                                p.outT_lo[(long long)(c0 + j) * p.ld_outT + row] = v[j] - h;
                            } else {
                                p.outT_hi[(long long)(c0 + j) * p.ld_outT + row] = v[j];
                            }
                        }
                    }
                }
                if (p.colsum) {
                    // column sums over the warp's 32 rows: transposing butterfly (31 shuffles for 32 columns),
                    // lane j ends up with the sum of column c0 + j
#pragma unroll
                    for (int w = 16; w >= 1; w >>= 1) {
#pragma unroll
                        for (int j = 0; j < w; j++) {
                            const bool upper = (lane & w) != 0;
                            const float send = upper ? v[j] : v[j + w];
                            const float keep = upper ? v[j + w] : v[j];
                            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, w);
                        }
                    }
                    // after the butterfly lane l holds column bitrev-free index: column (c0 + l) by construction
                    atomicAdd(&s_colsum[c0 + lane], v[0]);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.colsum) {
        for (int i = threadIdx.x; i < p.N; i += NUM_THREADS) {
            const float s = s_colsum[i];
            if (s != 0.0f) atomicAdd(p.colsum + i, s);
        }
    }
    if (threadIdx.x == 0 && *abort_flag && p.error_flag) atomicOr(p.error_flag, 1u);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side: tensor maps ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}
// [rows x cols] fp32, row stride ld elements; box = [box_rows x 16 cols], 64-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return gsr_set_error_msg(-5, "mlp_gemm: cuTensorMapEncodeTiled is not available from the driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3))
        return gsr_set_error_msg(-2, "mlp_gemm: operand planes must be 16-byte aligned with a row stride that is a multiple of 4 floats");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gsr_set_error_msg(-5, "mlp_gemm: cuTensorMapEncodeTiled failed");
    return 0;
}

// MN-major operand: [K rows x cols] fp32 row-major; box = [16 rows x 32 cols], 128-byte swizzle with 32-byte atoms
int make_map_mn(CUtensorMap* m, const float* base, long long rows, long long cols, long long ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return gsr_set_error_msg(-5, "mlp_gemm: cuTensorMapEncodeTiled is not available from the driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3))
        return gsr_set_error_msg(-2, "mlp_gemm: operand planes must be 16-byte aligned with a row stride that is a multiple of 4 floats");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)BK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gsr_set_error_msg(-5, "mlp_gemm: cuTensorMapEncodeTiled failed");
    return 0;
}

template <int BN>
int launch_mn(const GsrGemmArgs& g, cudaStream_t stream);

template <int BN>
int launch(const GsrGemmArgs& g, cudaStream_t stream) {
    using L = SmemLayout<BN>;
    if (g.mn_major) return launch_mn<BN>(g, stream);
    CUtensorMap a0h, a0l, a1h, a1l, bh, bl;
    const int kb0 = gsr_div_up(g.K0, BK), kb1 = g.A1_hi ? gsr_div_up(g.K1, BK) : 0;
    const bool convA = g.A0_lo == nullptr, convB = g.B_lo == nullptr;
    if (g.A1_hi && ((g.A1_lo == nullptr) != convA)) return gsr_set_error_msg(-2, "mlp_gemm: both A segments must be single planes or both pre-split");
    if (int rc = make_map(&a0h, g.A0_hi, g.M, g.K0, g.ldA0, BM)) return rc;
    if (convA) a0l = a0h; else if (int rc = make_map(&a0l, g.A0_lo, g.M, g.K0, g.ldA0, BM)) return rc;
    if (g.A1_hi) {
        if (int rc = make_map(&a1h, g.A1_hi, g.M, g.K1, g.ldA1, BM)) return rc;
        if (convA) a1l = a1h; else if (int rc = make_map(&a1l, g.A1_lo, g.M, g.K1, g.ldA1, BM)) return rc;
    } else { a1h = a0h; a1l = a0l; }
    // B [N rows x K cols]: K = K0, or 32 ceil(K0 / 32) + K1 with two segments.  The map's extent is the TRUE K (reads
    // beyond it are zero-filled by TMA, never fetched: the caller's padding may hold anything) and rows beyond N read
    // as zero, so the MMA always runs at the full BN.
    const long long bcols = g.A1_hi ? (long long)kb0 * BK + g.K1 : (long long)g.K0;
    if (g.ldB < bcols) return gsr_set_error_msg(-2, "mlp_gemm: ldB is smaller than K");
    if (int rc = make_map(&bh, g.B_hi, g.N, bcols, g.ldB, BN)) return rc;
    if (convB) bl = bh; else if (int rc = make_map(&bl, g.B_lo, g.N, bcols, g.ldB, BN)) return rc;
    GemmParams p{};
    p.M = g.M; p.N = g.N; p.kblocks[0] = kb0; p.kblocks[1] = kb1;
    p.m_tiles = gsr_div_up(g.M, BM);
    const int kb_total = kb0 + kb1;
    int splits = g.k_splits > 0 ? g.k_splits : 1;
    if (splits > kb_total) splits = kb_total;
    p.kb_per_split = gsr_div_up(kb_total, splits);
    p.k_splits = gsr_div_up(kb_total, p.kb_per_split);
    p.convA = convA ? 1 : 0; p.convB = convB ? 1 : 0;
    p.mode = g.mode; p.bias = g.bias; p.mask_src = g.mask_src; p.ld_mask = g.ld_mask;
    p.out_hi = g.out_hi; p.out_lo = g.out_lo; p.ld_out = g.ld_out;
    p.outT_hi = g.outT_hi; p.outT_lo = g.outT_lo; p.ld_outT = g.ld_outT;
    p.colsum = g.colsum; p.error_flag = g.error_flag;
    static bool attr_done = false;
    if (!attr_done) {
        GSR_CHECK(cudaFuncSetAttribute(mlp_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_done = true;
    }
    const int tiles = p.m_tiles * p.k_splits;
    const int grid = tiles < 148 ? tiles : 148;
    { GsrProfScope prof_(g.prof_name ? g.prof_name : "mlp_gemm", stream);
    mlp_gemm_kernel<BN><<<grid, NUM_THREADS, L::TOTAL, stream>>>(a0h, a0l, a1h, a1l, bh, bl, p); }
    GSR_CHECK_LAUNCH();
    return 0;
}

// Weight-gradient form: C[M x N] += A^T B with A [K x M], B [K x N] row-major single planes (K = points), split over K.
template <int BN>
int launch_mn(const GsrGemmArgs& g, cudaStream_t stream) {
    using L = SmemLayout<BN>;
    if (g.A0_lo || g.B_lo || g.A1_hi) return gsr_set_error_msg(-2, "mlp_gemm: MN-major operands are single fp32 planes, one K segment");
    if (g.mode != GSR_GEMM_ATOMIC) return gsr_set_error_msg(-2, "mlp_gemm: MN-major operands are for the atomic (split-K) mode");
    CUtensorMap ah, bh;
    if (int rc = make_map_mn(&ah, g.A0_hi, g.K0, g.M, g.ldA0)) return rc;
    if (int rc = make_map_mn(&bh, g.B_hi, g.K0, g.N, g.ldB)) return rc;
    GemmParams p{};
    p.M = g.M; p.N = g.N; p.kblocks[0] = gsr_div_up(g.K0, BK); p.kblocks[1] = 0;
    p.m_tiles = gsr_div_up(g.M, BM);
    const int kb_total = p.kblocks[0];
    int splits = g.k_splits > 0 ? g.k_splits : 1;
    if (splits > kb_total) splits = kb_total;
    p.kb_per_split = gsr_div_up(kb_total, splits);
    p.k_splits = gsr_div_up(kb_total, p.kb_per_split);
    p.convA = 1; p.convB = 1; p.mn_major = 1;
    p.mode = g.mode; p.out_hi = g.out_hi; p.ld_out = g.ld_out; p.error_flag = g.error_flag;
    static bool attr_done = false;
    if (!attr_done) {
        GSR_CHECK(cudaFuncSetAttribute(mlp_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_done = true;
    }
    const int tiles = p.m_tiles * p.k_splits;
    const int grid = tiles < 148 ? tiles : 148;
    { GsrProfScope prof_(g.prof_name ? g.prof_name : "mlp_gemm_dw", stream);
    mlp_gemm_kernel<BN><<<grid, NUM_THREADS, L::TOTAL, stream>>>(ah, ah, ah, ah, bh, bh, p); }
    GSR_CHECK_LAUNCH();
    return 0;
}

}  // namespace

int gsr_launch_mlp_gemm(const GsrGemmArgs& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.K0 <= 0) return 0;
    if (g.N > 256) return gsr_set_error_msg(-2, "mlp_gemm: N must be <= 256");
    if (!g.A0_hi || !g.B_hi || !g.out_hi) return gsr_set_error_msg(-1, "mlp_gemm: NULL operand");
    if (g.outT_lo && !g.outT_hi) return gsr_set_error_msg(-1, "mlp_gemm: transposed lo plane without hi plane");
    if (g.mask_src && ((g.ld_mask & 3) || (reinterpret_cast<uintptr_t>(g.mask_src) & 15)))
        return gsr_set_error_msg(-2, "mlp_gemm: mask source must be 16-byte aligned");
    if (g.N <= 64) return launch<64>(g, stream);
    if (g.N <= 128) return launch<128>(g, stream);
    return launch<256>(g, stream);
}

// ---- elementwise helpers of the network ------------------------------------------------------------------------
namespace {

// x -> (hi, lo) planes, optionally transposed ([rows x cols] -> [cols x ldT])
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, long long n, float* __restrict__ hi, float* __restrict__ lo) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float v = x[i], h = tf32_hi(v);
    hi[i] = h; lo[i] = v - h;
}
// W [rows x cols] (row stride ld_in) -> planes of W^T [cols x ldT] (weights only: tiny)
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, int rows, int cols, int ld_in,
                                                             float* __restrict__ hi, float* __restrict__ lo, int ldT) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i % cols;
    const float v = x[(size_t)r * ld_in + c], h = tf32_hi(v);
    hi[(size_t)c * ldT + r] = h; lo[(size_t)c * ldT + r] = v - h;
}

// x [rows x cols] (row stride ld_in) -> row-major planes [rows x ld_out] and/or transposed planes [cols x ldT], plus column
// sums: the operand preparation of a gradient that arrives from autograd (dL/d heads).  32 x 32 tiles through shared memory,
// so that both the reads and the transposed writes are 128-byte lines.
__global__ void __launch_bounds__(256) prepare_kernel(const float* __restrict__ x, int rows, int cols, long long ld_in,
                                                     float* __restrict__ hi, float* __restrict__ lo, long long ld_out,
                                                     float* __restrict__ hiT, float* __restrict__ loT, long long ldT,
                                                     float* __restrict__ colsum) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const long long r = r0 + ty + 8 * k;
        const int c = c0 + tx;
        const float v = (r < rows && c < cols) ? x[r * ld_in + c] : 0.0f;
        tile[ty + 8 * k][tx] = v;
        if (hi && r < rows && c < cols) {
            if (lo) { const float h = tf32_hi(v); hi[r * ld_out + c] = h; lo[r * ld_out + c] = v - h; }
            else hi[r * ld_out + c] = v;
        }
    }
    __syncthreads();
    if (hiT) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c0 + ty + 8 * k;
            const long long r = r0 + tx;
            if (c < cols && r < rows) {
                const float v = tile[tx][ty + 8 * k];
                if (loT) { const float h = tf32_hi(v); hiT[c * ldT + r] = h; loT[c * ldT + r] = v - h; }
                else hiT[c * ldT + r] = v;
            }
        }
    }
    if (colsum && ty == 0) {
        float sacc = 0.0f;
#pragma unroll
        for (int r = 0; r < 32; r++) sacc += tile[r][tx];
        if (c0 + tx < cols && sacc != 0.0f) atomicAdd(colsum + c0 + tx, sacc);
    }
}

// Positional embedding of gaussian_model.py:33-81 with multires 10: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^9 x), cos(2^9 x)]
// = 63 values per point, stored as hi | lo planes of 64 columns (column 63 = 0) and, for the weight gradients, transposed.
__global__ void __launch_bounds__(256) embed_kernel(const float* __restrict__ xyz, int P, float* __restrict__ e_hi, float* __restrict__ e_lo,
                                                   float* __restrict__ eT_hi, float* __restrict__ eT_lo, long long ldT) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    float x[3] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]};
    float e[64];
    e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
    for (int f = 0; f < 10; f++) {
        const float fr = (float)(1 << f);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            // torch: sin(x * freq), cos(x * freq) in fp32 (freq is an exact power of two)
            e[3 + 6 * f + d] = sinf(x[d] * fr);
            e[3 + 6 * f + 3 + d] = cosf(x[d] * fr);
        }
    }
    e[63] = 0.0f;
    float4* oh = reinterpret_cast<float4*>(e_hi + 64 * (size_t)i);
    float4* ol = e_lo ? reinterpret_cast<float4*>(e_lo + 64 * (size_t)i) : nullptr;
#pragma unroll
    for (int j4 = 0; j4 < 16; j4++) {
        if (ol) {
            float h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { h[k] = tf32_hi(e[4 * j4 + k]); l[k] = e[4 * j4 + k] - h[k]; }
            oh[j4] = make_float4(h[0], h[1], h[2], h[3]);
            ol[j4] = make_float4(l[0], l[1], l[2], l[3]);
        } else {
            oh[j4] = make_float4(e[4 * j4], e[4 * j4 + 1], e[4 * j4 + 2], e[4 * j4 + 3]);
        }
    }
    if (eT_hi) {
#pragma unroll
        for (int j = 0; j < 64; j++) {
            if (eT_lo) { const float h = tf32_hi(e[j]); eT_hi[(size_t)j * ldT + i] = h; eT_lo[(size_t)j * ldT + i] = e[j] - h; }
            else eT_hi[(size_t)j * ldT + i] = e[j];
        }
    }
}

// d(embedding)/dx: dx[d] = de[d] + sum_f 2^f (cos(2^f x_d) de_sin[f][d] - sin(2^f x_d) de_cos[f][d])
__global__ void __launch_bounds__(256) embed_bwd_kernel(const float* __restrict__ xyz, int P, const float* __restrict__ de /*[P x 64]*/,
                                                       float* __restrict__ dxyz, int accumulate) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float4* d4 = reinterpret_cast<const float4*>(de + 64 * (size_t)i);
    float d[64];
#pragma unroll
    for (int j4 = 0; j4 < 16; j4++) { const float4 q = d4[j4]; d[4 * j4] = q.x; d[4 * j4 + 1] = q.y; d[4 * j4 + 2] = q.z; d[4 * j4 + 3] = q.w; }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float x = xyz[3 * (size_t)i + k];
        float g = d[k];
#pragma unroll
        for (int f = 0; f < 10; f++) {
            const float fr = (float)(1 << f);
            float s, c;
            sincosf(x * fr, &s, &c);
            g += fr * (c * d[3 + 6 * f + k] - s * d[3 + 6 * f + 3 + k]);
        }
        float* o = dxyz + 3 * (size_t)i + k;
        *o = accumulate ? *o + g : g;
    }
}

}  // namespace

int gsr_launch_mlp_split(const float* x, long long n, float* hi, float* lo, cudaStream_t stream) {
    if (n <= 0) return 0;
    { GsrProfScope prof_("mlp_split", stream);
    split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, n, hi, lo); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_mlp_split_transpose(const float* x, int rows, int cols, int ld_in, float* hi, float* lo, int ldT, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    { GsrProfScope prof_("mlp_split_transpose", stream);
    split_transpose_kernel<<<gsr_div_up((long long)rows * cols, 256), 256, 0, stream>>>(x, rows, cols, ld_in, hi, lo, ldT); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_mlp_prepare(const float* x, int rows, int cols, long long ld_in, float* hi, float* lo, long long ld_out,
                           float* hiT, float* loT, long long ldT, float* colsum, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    const dim3 grid(gsr_div_up(rows, 32), gsr_div_up(cols, 32), 1);
    { GsrProfScope prof_("mlp_prepare", stream);
    prepare_kernel<<<grid, 256, 0, stream>>>(x, rows, cols, ld_in, hi, lo, ld_out, hiT, loT, ldT, colsum); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_mlp_embed(const float* xyz, int P, float* e_hi, float* e_lo, float* eT_hi, float* eT_lo, long long ldT, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("mlp_embed", stream);
    embed_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(xyz, P, e_hi, e_lo, eT_hi, eT_lo, ldT); }
    GSR_CHECK_LAUNCH();
    return 0;
}
int gsr_launch_mlp_embed_bwd(const float* xyz, int P, const float* de, float* dxyz, int accumulate, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("mlp_embed_bwd", stream);
    embed_bwd_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(xyz, P, de, dxyz, accumulate); }
    GSR_CHECK_LAUNCH();
    return 0;
}
