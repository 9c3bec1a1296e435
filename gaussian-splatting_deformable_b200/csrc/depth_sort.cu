// Depth order of the Gaussians: a two-level bucket sort of the P (depth bits, Gaussian id) pairs.
//
// Replaces the depth half of cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:303-308).  The reference's stable
// sort of (tile | depth) keys orders the duplicates of one tile by (depth bits, Gaussian index); the Gaussian index is
// unique, so "ascending (depth bits, index)" is a TOTAL order and any algorithm that realises it gives the same
// permutation - stability is not needed anywhere, which is what makes the cheap scheme below legal:
//
//   preprocess_fwd        min / max of the depth bits of the Gaussians that emit duplicates (one atomic pair per CTA)
//   ds_hist                bucket = (bits - min) >> shift: 2^nb equal slices of the occupied key range; counts by
//                          global atomics (the table is L2-resident)
//   ds_scan     (1 CTA)    scans the counts in place (-> bucket cursors) and cuts the bucketed array into segments of
//                          ~1024 entries at bucket boundaries
//   ds_scatter             pair -> its bucket, position by one returning atomic on the bucket cursor (any order)
//   ds_local_sort          one CTA per segment: pairs into registers (their tile rects requested at once), split again
//                          into 2048 sub-buckets of the segment's own key range (~0.5 pairs each) in shared memory, rank
//                          inside the sub-bucket by comparing (bits, id) with its few mates, write the id - and the
//                          (id, tile rect) record the tile counting sort consumes (what gather_rects did) - to its
//                          final position.
//
// P = 1 M (B200, CUDA events, which add ~5 us per launch): 11 + 17 + 15 + 23 us for the four launches, every array L2-resident - against 4 onesweep digit
// passes (each a ~23 us latency chain of ranking, decoupled look-back and reorder however few keys there are) + a
// histogram scan + gather_rects = 110 us.  Global atomics: a counter per 128-byte line - packed counters serialise in
// the L2 atomic unit (600 k atomics on a packed 32 KB table: 30 us; one per line: 12 us).
// Skew: a segment that does not fit shared memory (one bucket > 2048 pairs: a key range populated > 25x its average
// density) or holds a sub-bucket of > 64 pairs (many equal depths) is sorted by its CTA with a plain stable LSD radix
// sort on (id, varying key bits) in global memory - slow but exact, and only that CTA pays.
#include "kernels.cuh"

namespace {

constexpr int DS_T = 1024;                 // nominal pairs per segment
constexpr int DS_CAP = 3072;               // pairs a CTA sorts in shared memory (T + the largest admissible bucket)
#ifndef GSR_DS_STRIDE
#define GSR_DS_STRIDE 32                  // u32 words between bucket counters: one counter per 128-byte line (the L2 atomic
#endif                                     // unit serialises the atomics of a line; packed counters: 30 us per 600 k atomics)
constexpr int DS_STRIDE = GSR_DS_STRIDE;
constexpr int DS_NSB_BITS = 11;             // measured at P = 1 M: 1024 -> 29 us, 2048 -> 23 us, 4096 -> 29 us for the local pass
constexpr int DS_NSB = 1 << DS_NSB_BITS;   // sub-buckets per segment
constexpr int DS_MAX_SUB = 64;             // largest sub-bucket ranked by all-pairs comparison
constexpr int DS_THREADS = 256;
constexpr int DS_HIST_THREADS = 1024;
constexpr unsigned FULLM = 0xffffffffu;

constexpr int DS_EPT = DS_CAP / DS_THREADS; // pairs per thread (registers), largest segment
constexpr int DS_EPT_SMALL = 6;             // ... of the usual segment (<= 1536 pairs): no register spills on that path
struct LocalSmem {
    uint4 f[DS_CAP];            // 48 KB: (key, id, rect lo, rect hi) grouped by sub-bucket
    uint32_t off[DS_NSB];       //  8 KB: sub-bucket counts -> starts
    uint32_t red[64];
};

__device__ __forceinline__ int key_bits(uint32_t range) { return 32 - __clz(range); }   // 0 for range == 0

// ---- 1. bucket histogram ------------------------------------------------------------------------------
__global__ void __launch_bounds__(DS_HIST_THREADS) ds_hist_kernel(const uint32_t* __restrict__ keys, uint32_t P, uint32_t* state,
                                                                   int nb_bits) {
    const uint32_t* const ctrl = state;
    uint32_t* const hist = state + GSR_DS_CTRL_WORDS;
    const int tid = threadIdx.x;
    const uint32_t kmin = ~ctrl[GSR_DS_NOT_KMIN], kmax = ctrl[GSR_DS_KMAX];
    const uint32_t range = kmax >= kmin ? kmax - kmin : 0u;          // no emitting Gaussian: kmin = ~0, kmax = 0
    const int bits = key_bits(range);
    const int shift = bits > nb_bits ? bits - nb_bits : 0;
    // four keys per thread and round: the loads of a round are in flight together
    for (uint32_t i0 = blockIdx.x * (DS_HIST_THREADS * 4) + tid; i0 < P; i0 += gridDim.x * (DS_HIST_THREADS * 4)) {
        uint32_t k[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t i = i0 + j * DS_HIST_THREADS;
            k[j] = i < P ? __ldg(keys + i) : 0xffffffffu;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (k[j] != 0xffffffffu) atomicAdd(&hist[(size_t)((k[j] - kmin) >> shift) * DS_STRIDE], 1u);
    }
}

// ---- 1b. one CTA: exclusive scan of the counts in place (-> scatter cursors) and the segment cuts ----------------
// (as the last-finishing CTA of ds_hist it took the same time: 8192 counters on 8192 lines are ~10 us for one SM either way)
__global__ void __launch_bounds__(DS_HIST_THREADS) ds_scan_kernel(uint32_t* state, int nb_bits, uint32_t nseg) {
    uint32_t* const ctrl = state;
    uint32_t* const hist = state + GSR_DS_CTRL_WORDS;
    const uint32_t NB = 1u << nb_bits;
    uint32_t* const seg_lo = hist + (size_t)NB * DS_STRIDE;
    __shared__ uint32_t s_warp[32], s_last[32];
    __shared__ uint32_t s_misc[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t kmin = ~ctrl[GSR_DS_NOT_KMIN], kmax = ctrl[GSR_DS_KMAX];
    const uint32_t range = kmax >= kmin ? kmax - kmin : 0u;
    const int bits = key_bits(range);
    const int shift = bits > nb_bits ? bits - nb_bits : 0;
    // ---- exclusive scan of the NB counts, in place (they become the scatter cursors), 8 buckets per thread and round ----
    constexpr int PER = 8;
    uint32_t carry = 0;                                               // base of the round's first bucket
    uint32_t carry_prev_cnt = 0;                                      // count of the bucket just before the round
    for (uint32_t rb = 0; rb < NB; rb += DS_HIST_THREADS * PER) {
        const uint32_t b0 = rb + tid * PER;
        uint32_t cnt[PER];
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            cnt[j] = (b0 + j < NB) ? __ldcg(hist + (size_t)(b0 + j) * DS_STRIDE) : 0u;
            sum += cnt[j];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULLM, incl, o);
            if (lane >= o) incl += t;
        }
        // the last count of the thread before this one (bucket b0 - 1)
        uint32_t cnt_prev = __shfl_up_sync(FULLM, cnt[PER - 1], 1);
        if (lane == 31) { s_warp[warp] = incl; s_last[warp] = cnt[PER - 1]; }
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 32; w++) {
            const uint32_t t = s_warp[w];
            wbase += (w < warp) ? t : 0u;
            total += t;
        }
        if (lane == 0) cnt_prev = warp ? s_last[warp - 1] : carry_prev_cnt;
        uint32_t run = carry + wbase + incl - sum;                    // base of bucket b0
        uint32_t prev_base = run - cnt_prev;                          // base of bucket b0 - 1
#pragma unroll
        for (int j = 0; j < PER; j++) {
            if (b0 + j < NB) {
                const uint32_t b = b0 + j;
                // segment c starts at the first bucket whose base is >= c T: the multiples of T in (base[b-1], base[b]]
                if (b > 0)
                    for (uint32_t c = prev_base / DS_T + 1; c <= run / DS_T; c++) seg_lo[c] = run;
                hist[(size_t)b * DS_STRIDE] = run;
                prev_base = run;
                run += cnt[j];
            }
        }
        if (b0 < NB && b0 + PER >= NB) s_misc[1] = prev_base;        // base of the last bucket
        const uint32_t last_cnt = s_last[31];
        __syncthreads();
        carry += total;
        carry_prev_cnt = last_cnt;
    }
    const uint32_t total = carry;
    const uint32_t last_base = s_misc[1];
    for (uint32_t c = last_base / DS_T + 1 + tid; c <= nseg; c += DS_HIST_THREADS) seg_lo[c] = total;
    if (tid == 0) {
        seg_lo[0] = 0;
        ctrl[GSR_DS_N_EMIT] = total;
        ctrl[GSR_DS_SHIFT] = (uint32_t)shift;
        ctrl[GSR_DS_KMIN] = kmin;
    }
}

// ---- 2. pairs into their buckets (any order inside a bucket) ---------------------------------------------
__global__ void __launch_bounds__(256) ds_scatter_kernel(const uint32_t* __restrict__ keys, uint32_t P, uint32_t* state,
                                                          uint2* __restrict__ pairs) {
    const uint32_t* const ctrl = state;
    uint32_t* const cursor = state + GSR_DS_CTRL_WORDS;
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= P) return;
    const uint32_t k = __ldg(keys + i);
    if (k == 0xffffffffu) return;
    const uint32_t b = (k - ctrl[GSR_DS_KMIN]) >> ctrl[GSR_DS_SHIFT];
    const uint32_t pos = atomicAdd(&cursor[(size_t)b * DS_STRIDE], 1u);
    pairs[pos] = make_uint2(k, i);
}

// One stable 8-bit LSD pass of a segment by ONE CTA (the slow path): src[0..n) -> dst[0..n).
// pass < id_passes: digit of the id; else digit of the key.
__device__ void cta_lsd_pass(const uint2* src, uint2* dst, uint32_t n, int which, int dshift, uint32_t* s_hist /*[256]*/,
                             uint32_t (*s_wh)[256] /*[8][256], zero on entry and on exit*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_hist[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += DS_THREADS) {
        const uint2 e = src[i];
        atomicAdd(&s_hist[((which ? e.x : e.y) >> dshift) & 255u], 1u);
    }
    __syncthreads();
    {   // exclusive scan of the 256 counts (thread d owns digit d)
        __shared__ uint32_t s_wt[8];
        const uint32_t v = s_hist[tid];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULLM, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wt[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) base += (w < warp) ? s_wt[w] : 0u;
        s_hist[tid] = base + incl - v;
    }
    __syncthreads();
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t c0 = 0; c0 < n; c0 += DS_THREADS) {
        const uint32_t i = c0 + tid;
        const bool valid = i < n;
        uint2 e = make_uint2(0, 0);
        if (valid) e = src[i];
        const uint32_t d = valid ? (((which ? e.x : e.y) >> dshift) & 255u) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(FULLM, d);
        const bool leader = valid && lane == __ffs(peers) - 1;
        if (leader) s_wh[warp][d] = (uint32_t)__popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) off += (w < warp) ? s_wh[w][d] : 0u;
            dst[s_hist[d] + off + (uint32_t)__popc(peers & lt)] = e;
        }
        __syncthreads();
        if (leader) { atomicAdd(&s_hist[d], (uint32_t)__popc(peers)); s_wh[warp][d] = 0; }
        __syncthreads();
    }
}

// Fast path of a segment that fits shared memory, pairs held in registers (EPT per thread).  Returns false (nothing
// written) when a sub-bucket is too large for all-pairs ranking; *bits_out = key bits the LSD path has to sort on.
template <int EPT>
__device__ __forceinline__ bool local_sort_fast(LocalSmem& s, const uint2* __restrict__ A, uint32_t n, uint32_t lo,
                                                const uint2* __restrict__ rects, uint32_t* __restrict__ order,
                                                uint4* __restrict__ srec, int* bits_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t kmn = 0xffffffffu, kmx = 0u;
    uint2 e[EPT], rc[EPT];
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        if (k * DS_THREADS >= (int)n) break;                           // CTA-uniform
        const uint32_t i = tid + k * DS_THREADS;
        if (i < n) { e[k] = A[i]; kmn = min(kmn, e[k].x); kmx = max(kmx, e[k].x); }
    }
    // the tile rect of every pair (a dependent random gather) is requested now and consumed after the sub-bucket scan
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        if (k * DS_THREADS >= (int)n) break;
        rc[k] = (rects && tid + k * DS_THREADS < n) ? __ldg(rects + e[k].y) : make_uint2(0u, 0u);
    }
    kmn = __reduce_min_sync(FULLM, kmn); kmx = __reduce_max_sync(FULLM, kmx);
    if (lane == 0) { s.red[warp] = kmn; s.red[8 + warp] = kmx; }
    for (int i = tid; i < DS_NSB; i += DS_THREADS) s.off[i] = 0;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; w++) { kmn = min(kmn, s.red[w]); kmx = max(kmx, s.red[8 + w]); }
    const int bits = key_bits(kmx - kmn);
    *bits_out = key_bits(kmx ^ kmn);                                   // highest bit in which two keys of the segment can differ
    const int s2 = bits > DS_NSB_BITS ? bits - DS_NSB_BITS : 0;
    uint32_t r[EPT];                                                   // arrival rank inside the sub-bucket
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        if (k * DS_THREADS >= (int)n) break;
        if (tid + k * DS_THREADS < n) r[k] = atomicAdd(&s.off[(e[k].x - kmn) >> s2], 1u);
    }
    __syncthreads();
    // exclusive scan of the DS_NSB counts (8 per thread) + the largest count
    constexpr int PER = DS_NSB / DS_THREADS;
    uint32_t c[PER], sum = 0, mx = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) { c[j] = s.off[tid * PER + j]; sum += c[j]; mx = max(mx, c[j]); }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULLM, incl, o);
        if (lane >= o) incl += t;
    }
    mx = __reduce_max_sync(FULLM, mx);
    if (lane == 31) s.red[16 + warp] = incl;
    if (lane == 0) s.red[24 + warp] = mx;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { base += (w < warp) ? s.red[16 + w] : 0u; mx = max(mx, s.red[24 + w]); }
    if (mx > (uint32_t)DS_MAX_SUB) return false;                       // CTA-uniform
    uint32_t run = base + incl - sum;
#pragma unroll
    for (int j = 0; j < PER; j++) { s.off[tid * PER + j] = run; run += c[j]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        if (k * DS_THREADS >= (int)n) break;
        if (tid + k * DS_THREADS < n) s.f[s.off[(e[k].x - kmn) >> s2] + r[k]] = make_uint4(e[k].x, e[k].y, rc[k].x, rc[k].y);
    }
    __syncthreads();
    // rank among the few mates of the sub-bucket by (bits, id); the final position is known then
    for (uint32_t p = tid; p < n; p += DS_THREADS) {
        const uint4 x = s.f[p];
        const uint32_t sb = (x.x - kmn) >> s2;
        const uint32_t st = s.off[sb], en = sb + 1 < (uint32_t)DS_NSB ? s.off[sb + 1] : n;
        const unsigned long long xk = ((unsigned long long)x.x << 32) | x.y;
        uint32_t rank = 0;
        for (uint32_t q = st; q < en; q++) {
            const uint2 m = *reinterpret_cast<const uint2*>(&s.f[q]);
            rank += ((((unsigned long long)m.x << 32) | m.y) < xk) ? 1u : 0u;
        }
        const uint32_t dst = lo + st + rank;
        order[dst] = x.y;
        if (srec) srec[dst] = make_uint4(x.y, x.z, x.w, 0u);
    }
    return true;
}

// ---- 3. one CTA per segment: final order of its pairs ---------------------------------------------------------
__global__ void __launch_bounds__(DS_THREADS, 4) ds_local_sort_kernel(uint2* pairs_a, uint2* pairs_b, uint32_t* state, int nb_bits,
                                                                       uint32_t id_bits, const uint2* __restrict__ rects,
                                                                       uint32_t* __restrict__ order, uint4* __restrict__ srec) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LocalSmem& s = *reinterpret_cast<LocalSmem*>(smem_raw);
    const uint32_t* const seg_lo = state + GSR_DS_CTRL_WORDS + ((size_t)DS_STRIDE << nb_bits);
    const uint32_t lo = seg_lo[blockIdx.x], hi = seg_lo[blockIdx.x + 1];
    if (hi <= lo) return;
    const uint32_t n = hi - lo;
    const int tid = threadIdx.x;
    const uint2* const A = pairs_a + lo;
    int bits = 32;
    if (n <= (uint32_t)DS_EPT_SMALL * DS_THREADS) {                     // the usual segment: T + one bucket
        if (local_sort_fast<DS_EPT_SMALL>(s, A, n, lo, rects, order, srec, &bits)) return;
    } else if (n <= (uint32_t)DS_CAP) {
        if (local_sort_fast<DS_EPT>(s, A, n, lo, rects, order, srec, &bits)) return;
    } else {
        uint32_t kmn = 0xffffffffu, kmx = 0u;
        for (uint32_t i = tid; i < n; i += DS_THREADS) {
            const uint32_t k = A[i].x;
            kmn = min(kmn, k); kmx = max(kmx, k);
        }
        kmn = __reduce_min_sync(FULLM, kmn); kmx = __reduce_max_sync(FULLM, kmx);
        if ((tid & 31) == 0) { s.red[tid >> 5] = kmn; s.red[8 + (tid >> 5)] = kmx; }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < 8; w++) { kmn = min(kmn, s.red[w]); kmx = max(kmx, s.red[8 + w]); }
        bits = key_bits(kmx ^ kmn);       // NOT the range: 0x..ff and 0x..100 are 1 apart and differ in bit 8
    }
    // ---- slow path: stable LSD radix sort of the segment in global memory, by this CTA alone ----
    __syncthreads();
    uint32_t* const s_hist = s.off;                                   // [256]
    uint32_t (*s_wh)[256] = reinterpret_cast<uint32_t (*)[256]>(s.f); // [8][256] = 8 KB of the pair array
    for (int i = tid; i < 8 * 256; i += DS_THREADS) (&s_wh[0][0])[i] = 0;
    if (tid == 0) atomicAdd(&state[GSR_DS_SLOW_SEGMENTS], 1u);
    __syncthreads();
    uint2* src = pairs_a + lo;
    uint2* dst = pairs_b + lo;
    const int id_passes = ((int)id_bits + 7) / 8, key_passes = (bits + 7) / 8;
    for (int p = 0; p < id_passes + key_passes; p++) {
        const int which = p >= id_passes;
        cta_lsd_pass(src, dst, n, which, 8 * (which ? p - id_passes : p), s_hist, s_wh);
        __threadfence_block();
        __syncthreads();
        uint2* t = src; src = dst; dst = t;
    }
    for (uint32_t i = tid; i < n; i += DS_THREADS) {
        const uint32_t id = src[i].y;
        const uint2 rc = rects ? __ldg(rects + id) : make_uint2(0u, 0u);
        order[lo + i] = id;
        if (srec) srec[lo + i] = make_uint4(id, rc.x, rc.y, 0u);
    }
}

// stand-alone use (gsr_debug_depth_order): min / max of the valid keys
__global__ void __launch_bounds__(256) ds_minmax_kernel(const uint32_t* __restrict__ keys, uint32_t P, uint32_t* state) {
    uint32_t kmn = 0xffffffffu, kmx = 0u;
    bool any = false;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < P; i += gridDim.x * 256u) {
        const uint32_t k = keys[i];
        if (k != 0xffffffffu) { kmn = min(kmn, k); kmx = max(kmx, k); any = true; }
    }
    kmn = __reduce_min_sync(FULLM, kmn); kmx = __reduce_max_sync(FULLM, kmx);
    any = __any_sync(FULLM, any);
    if ((threadIdx.x & 31) == 0 && any) { atomicMax(&state[GSR_DS_NOT_KMIN], ~kmn); atomicMax(&state[GSR_DS_KMAX], kmx); }
}

}  // namespace

int gsr_depth_sort_bucket_bits(uint32_t P) {
    int lg = 0;
    while ((1ull << lg) < (unsigned long long)P) lg++;
    int nb = lg - 7;
    if (nb < 6) nb = 6;
    if (nb > 15) nb = 15;
    return nb;
}
static inline uint32_t ds_num_segments(uint32_t P) { return P / DS_T + 1; }

size_t gsr_depth_sort_state_bytes(uint32_t P) {
    return sizeof(uint32_t) * ((size_t)GSR_DS_CTRL_WORDS + ((size_t)DS_STRIDE << gsr_depth_sort_bucket_bits(P)) + ds_num_segments(P) + 2);
}

int gsr_launch_depth_sort(uint32_t P, const uint32_t* keys, uint32_t* state, uint2* pairs_a, uint2* pairs_b, const uint2* rects,
                          uint32_t* order, uint4* srec, bool minmax_ready, cudaStream_t stream) {
    if (P == 0) return 0;
    if (P >= (1u << 30)) return gsr_set_error_msg(-2, "depth sort: P must be < 2^30");
    static bool attr_set = false;
    if (!attr_set) {
        GSR_CHECK(cudaFuncSetAttribute(ds_local_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LocalSmem)));
        attr_set = true;
    }
    const int nb_bits = gsr_depth_sort_bucket_bits(P);
    const uint32_t nseg = ds_num_segments(P);
    if (!minmax_ready) {
        int blocks = gsr_div_up(P, 256 * 8);
        if (blocks > 148 * 4) blocks = 148 * 4;
        ds_minmax_kernel<<<blocks, 256, 0, stream>>>(keys, P, state);
        GSR_CHECK_LAUNCH();
    }
    int hist_blocks = gsr_div_up(P, DS_HIST_THREADS * 4);
    if (hist_blocks > 148 * 2) hist_blocks = 148 * 2;
    { GsrProfScope prof_("depth_sort_hist", stream);
    ds_hist_kernel<<<hist_blocks, DS_HIST_THREADS, 0, stream>>>(keys, P, state, nb_bits); }
    GSR_CHECK_LAUNCH();
    { GsrProfScope prof_("depth_sort_scan", stream);
    ds_scan_kernel<<<1, DS_HIST_THREADS, 0, stream>>>(state, nb_bits, nseg); }
    GSR_CHECK_LAUNCH();
    { GsrProfScope prof_("depth_sort_scatter", stream);
    ds_scatter_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(keys, P, state, pairs_a); }
    GSR_CHECK_LAUNCH();
    uint32_t id_bits = 1;
    while (id_bits < 32 && (1ull << id_bits) < (unsigned long long)P) id_bits++;
    { GsrProfScope prof_("depth_sort_local", stream);
    ds_local_sort_kernel<<<nseg, DS_THREADS, sizeof(LocalSmem), stream>>>(pairs_a, pairs_b, state, nb_bits, id_bits, rects, order, srec); }
    GSR_CHECK_LAUNCH();
    return 0;
}
