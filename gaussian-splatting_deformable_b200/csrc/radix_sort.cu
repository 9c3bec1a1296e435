// Onesweep LSD radix sort of (u64 key, u32 value) pairs on a bit range.
//
// Replaces cub::DeviceRadixSort::SortPairs<uint64_t,uint32_t> as called at
// rasterizer_impl.cu:303-308 (stable, ascending, bits [0, 32+tile_bits)).  The
// result of a stable sort is fully specified, so the output is bit-identical to
// the reference's whatever the algorithm.
//
// Structure (Adinets & Merrill "Onesweep", one read + one write of the pairs per
// digit pass):
//   1. one histogram kernel reads the keys ONCE and builds the digit histograms
//      of every pass (shared-memory atomics, then 256*passes global atomics/block);
//   2. a tiny kernel turns them into exclusive digit offsets;
//   3. per pass, one kernel: each CTA takes a tile of 256 thr x 16 items through a
//      ticket counter (so a CTA only ever waits on CTAs that already started -
//      no deadlock whatever the hardware scheduling order), ranks its items
//      stably with warp match-any, publishes its per-digit counts, resolves its
//      global offsets by decoupled look-back over the predecessor tiles, reorders
//      the tile in shared memory and writes digit-contiguous (coalesced) runs.
// HBM-bound: 8 B/key for the histograms + 24 B/pair/pass.
#include "kernels.cuh"

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = GSR_SORT_RADIX;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int ITEMS = 16;
constexpr int TILE_ITEMS = SORT_THREADS * ITEMS;   // 4096
constexpr int MAX_PASSES = GSR_SORT_MAX_PASSES;
#ifndef GSR_SORT_LOOKBACK
#define GSR_SORT_LOOKBACK 8
#endif
constexpr uint32_t FLAG_PARTIAL = 1u << 30;
constexpr uint32_t FLAG_INCLUSIVE = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

typedef GsrSortPlan SortPlan;
inline SortPlan make_plan(int begin_bit, int end_bit) { return gsr_make_sort_plan(begin_bit, end_bit); }

// ---- 1. histograms of all passes in one read of the keys --------------------
template <typename KeyT>
__global__ void __launch_bounds__(256) histogram_kernel(const KeyT* __restrict__ keys, uint32_t n, SortPlan plan,
                                                        uint32_t* __restrict__ g_hist /*[passes][RADIX]*/) {
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < plan.passes * RADIX; i += 256) s_hist[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * 256u;
    const int lane = threadIdx.x & 31;
    // Warp-uniform trip count so match-any can aggregate equal digits: tile-id and
    // exponent digits take a handful of values, plain atomics would serialise 32-way.
    for (uint32_t i0 = blockIdx.x * 256u; i0 < n; i0 += stride) {
        const uint32_t i = i0 + threadIdx.x;
        const bool valid = i < n;
        const KeyT k = valid ? keys[i] : (KeyT)0;
#pragma unroll
        for (int p = 0; p < MAX_PASSES; p++) {
            if (p < plan.passes) {
                const uint32_t d = valid ? ((uint32_t)(k >> plan.shift[p]) & plan.mask[p]) : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_hist[p * RADIX + d], (uint32_t)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * RADIX; i += 256) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&g_hist[i], c);
    }
}

// ---- 2. exclusive scan of each pass' 256 bins -------------------------------
__global__ void __launch_bounds__(RADIX) scan_hist_kernel(uint32_t* g_hist) {
    __shared__ uint32_t wt[RADIX / 32];
    uint32_t* h = g_hist + blockIdx.x * RADIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = h[threadIdx.x];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wt[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; w++) base += (w < warp) ? wt[w] : 0u;
    h[threadIdx.x] = base + incl - v;
}

// ---- 3. one onesweep digit pass ---------------------------------------------
template <typename KeyT>
struct __align__(16) SortSmem {
    KeyT keys[TILE_ITEMS];                   // 32 KB (u64) / 16 KB (u32)
    uint32_t vals[TILE_ITEMS];               // 16 KB
    uint32_t warp_hist[SORT_WARPS][RADIX];   //  8 KB  per-warp digit counts -> per-warp digit offsets
    uint32_t digit_base[RADIX];              // global position of the tile's first item of each digit, minus its local start
    uint32_t local_start[RADIX];
    uint32_t warp_tot[SORT_WARPS];
    uint32_t tile;
};

template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS, sizeof(KeyT) == 4 ? 5 : 3)
onesweep_kernel(const KeyT* __restrict__ keys_in, KeyT* __restrict__ keys_out,
                const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out, uint32_t n,
                const uint32_t* __restrict__ g_offsets /*[RADIX] exclusive*/, volatile uint32_t* status /*[tiles][RADIX]*/,
                uint32_t* ticket, uint32_t* err_flag, int shift, uint32_t mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem<KeyT>& s = *reinterpret_cast<SortSmem<KeyT>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s.tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s.tile;
    const uint32_t tile_base = tile * (uint32_t)TILE_ITEMS;
    const uint32_t warp_base = tile_base + warp * (32u * ITEMS);

    // Warp-striped load: item j of lane l is element warp_base + j*32 + l, so the
    // (j, lane) order IS the input order - which stability requires.
    // (values are loaded only after the look-back, when the key registers are free:
    //  fewer live registers -> more resident CTAs to hide the latency of this chain)
    KeyT key[ITEMS];
    uint32_t rank[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const uint32_t i = warp_base + j * 32 + lane;
        key[j] = (i < n) ? keys_in[i] : (KeyT)~(KeyT)0;
    }
    // Stable ranking inside the warp: match-any groups lanes with equal digits; the
    // group's first lane bumps the warp's digit counter with ONE shared-memory atomic
    // that returns the count of earlier items (no load/store/__syncwarp round trip, so
    // the 16 items' match/atomic/shuffle chains overlap instead of serialising).
    // A warp's shared-memory instructions execute in program order, so successive
    // items see the counter values in item order - which is what stability needs.
    uint32_t* wh = s.warp_hist[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const uint32_t i = warp_base + j * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & mask) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) prev = atomicAdd(&wh[d], (uint32_t)__popc(peers));
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[j] = prev + __popc(peers & lt_mask);
    }
    __syncthreads();

    // Thread d owns digit d: offsets of each warp inside the digit, tile count,
    // publication and decoupled look-back.
    {
        const int d = tid;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const uint32_t c = s.warp_hist[w][d];
            s.warp_hist[w][d] = sum;
            sum += c;
        }
        volatile uint32_t* my = status + (size_t)tile * RADIX + d;
        *my = (tile == 0 ? FLAG_INCLUSIVE : FLAG_PARTIAL) | sum;
        uint32_t excl = 0;
        if (tile > 0) {
            // Decoupled look-back, LOOKBACK predecessors per round trip: fetch a window of status words at once and
            // consume the ready prefix.  (Measured, round 2: a window of 32 instead of 8 makes a pass 20 % SLOWER at
            // n = 1 M - the pass is bound by the per-CTA serial work at 1-2 resident CTAs per SM, not by this chain.)
            constexpr int LOOKBACK = GSR_SORT_LOOKBACK;
            int t = (int)tile - 1;
            uint32_t polls = 0;
            bool found = false;
            while (!found) {
                uint32_t st[LOOKBACK];
#pragma unroll
                for (int b = 0; b < LOOKBACK; b++)
                    st[b] = (t - b >= 0) ? status[(size_t)(t - b) * RADIX + d] : (2u << 30) /* FLAG_INCLUSIVE, value 0 */;
                int used = 0;
#pragma unroll
                for (int b = 0; b < LOOKBACK; b++) {
                    const uint32_t fl = st[b] & FLAG_MASK;
                    if (!found && used == b && fl != 0) {      // ready and contiguous with what was consumed
                        excl += st[b] & VALUE_MASK;
                        used = b + 1;
                        if (fl == FLAG_INCLUSIVE) found = true;
                    }
                }
                t -= used;
                if (used == 0) {
                    // Tickets make this wait finite; the poll bound only turns a would-be
                    // hang (e.g. a corrupted workspace) into a reported error.
                    if (++polls > (1u << 24)) { atomicOr(err_flag, 1u); break; }
                }
            }
            *my = FLAG_INCLUSIVE | (excl + sum);
        }
        // block-wide exclusive scan of the tile's digit counts
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t2 = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t2;
        }
        if (lane == 31) s.warp_tot[warp] = incl;
        __syncthreads();
        uint32_t wb = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) wb += (w < warp) ? s.warp_tot[w] : 0u;
        const uint32_t lstart = wb + incl - sum;
        s.local_start[d] = lstart;
        s.digit_base[d] = g_offsets[d] + excl - lstart;
    }
    __syncthreads();

    // Reorder the tile by digit in shared memory: keys first (rank[] becomes the slot) ...
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
        const uint32_t i = warp_base + j * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[j] >> shift) & mask;
            const uint32_t pos = s.local_start[d] + s.warp_hist[warp][d] + rank[j];
            s.keys[pos] = key[j];
            rank[j] = pos;
        }
    }
    // ... then the values, loaded now that the key registers are dead.
    {
        uint32_t val[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const uint32_t i = warp_base + j * 32 + lane;
            val[j] = (i < n) ? vals_in[i] : 0u;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const uint32_t i = warp_base + j * 32 + lane;
            if (i < n) s.vals[rank[j]] = val[j];
        }
    }
    __syncthreads();
    // ... and write digit-contiguous runs.
    const uint32_t count = min((uint32_t)TILE_ITEMS, n - tile_base);
    for (uint32_t i = tid; i < count; i += SORT_THREADS) {
        const KeyT k = s.keys[i];
        const uint32_t d = (uint32_t)(k >> shift) & mask;
        const uint32_t pos = s.digit_base[d] + i;
        keys_out[pos] = k;
        vals_out[pos] = s.vals[i];
    }
}

}  // namespace

GsrSortPlan gsr_make_sort_plan(int begin_bit, int end_bit) {
    GsrSortPlan p{};
    const int bits = end_bit - begin_bit;
    if (bits <= 0) return p;
    int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    if (passes > MAX_PASSES) passes = MAX_PASSES;
    int bit = begin_bit;
    for (int i = 0; i < passes; i++) {
        const int left = end_bit - bit;
        int nb = (left + (passes - i) - 1) / (passes - i);     // even split: 13 bits -> 7 + 6
        if (nb > RADIX_BITS) nb = RADIX_BITS;
        p.shift[i] = bit;
        p.mask[i] = (1u << nb) - 1u;
        bit += nb;
    }
    p.passes = passes;
    return p;
}

static inline uint32_t sort_num_tiles(uint32_t n) { return (n + TILE_ITEMS - 1) / TILE_ITEMS; }

// temp layout: [hist: MAX_PASSES*RADIX u32][tickets: MAX_PASSES u32 (padded to 64; word 63 = error flag)][status: passes*tiles*RADIX u32]
size_t gsr_sort_temp_bytes(uint32_t n, int begin_bit, int end_bit) {
    const SortPlan p = make_plan(begin_bit, end_bit);
    return sizeof(uint32_t) * ((size_t)MAX_PASSES * RADIX + 64 + (size_t)p.passes * sort_num_tiles(n) * RADIX);
}

template <typename KeyT>
static int sort_pairs_impl(KeyT* keys_a, KeyT* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t n, int begin_bit,
                           int end_bit, void* temp, size_t temp_bytes, int* result_in_b, cudaStream_t stream,
                           int site = 0, bool hist_ready = false) {
    // profiler labels: which call site of the pipeline this sort serves
    static const char* const kHist[3] = {"sort_histogram", "depth_sort_histogram", "tile_sort_histogram"};
    static const char* const kPass[3] = {"sort_onesweep_pass", "depth_sort_onesweep_pass", "tile_sort_onesweep_pass"};
    *result_in_b = 0;
    if (n == 0 || end_bit <= begin_bit) return 0;
    if (n >= (1u << 30)) return gsr_set_error_msg(-2, "radix sort: n must be < 2^30");
    if (end_bit > (int)sizeof(KeyT) * 8) return gsr_set_error_msg(-2, "radix sort: end_bit exceeds the key width");
    const SortPlan plan = make_plan(begin_bit, end_bit);
    const size_t need = gsr_sort_temp_bytes(n, begin_bit, end_bit);
    if (temp_bytes < need) return gsr_set_error_msg(-3, "radix sort: temp buffer too small");
    static bool attr_set = false;
    if (!attr_set) {
        GSR_CHECK(cudaFuncSetAttribute(onesweep_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(SortSmem<KeyT>)));
        attr_set = true;
    }
    uint32_t* hist = reinterpret_cast<uint32_t*>(temp);
    uint32_t* tickets = hist + MAX_PASSES * RADIX;
    uint32_t* status = tickets + 64;
    const uint32_t tiles = sort_num_tiles(n);
    if (!hist_ready) {
        GSR_CHECK(cudaMemsetAsync(temp, 0, need, stream));
        int hist_blocks = (int)((n + 256u * 16u - 1) / (256u * 16u));
        if (hist_blocks > 148 * 8) hist_blocks = 148 * 8;
        { GsrProfScope prof_(kHist[site], stream);
        histogram_kernel<KeyT><<<hist_blocks, 256, 0, stream>>>(keys_a, n, plan, hist); }
        GSR_CHECK_LAUNCH();
    }
    { GsrProfScope prof_("sort_scan_hist", stream);
    scan_hist_kernel<<<plan.passes, RADIX, 0, stream>>>(hist); }
    GSR_CHECK_LAUNCH();
    KeyT* kin = keys_a; KeyT* kout = keys_b;
    uint32_t* vin = vals_a; uint32_t* vout = vals_b;
    for (int p = 0; p < plan.passes; p++) {
        { GsrProfScope prof_(kPass[site], stream);
        onesweep_kernel<KeyT><<<tiles, SORT_THREADS, sizeof(SortSmem<KeyT>), stream>>>(
            kin, kout, vin, vout, n, hist + p * RADIX, status + (size_t)p * tiles * RADIX, tickets + p,
            tickets + 63, plan.shift[p], plan.mask[p]); }
        GSR_CHECK_LAUNCH();
        KeyT* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *result_in_b = (plan.passes & 1) ? 1 : 0;
    return 0;
}

int gsr_launch_sort_pairs(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t n,
                          int begin_bit, int end_bit, void* temp, size_t temp_bytes, int* result_in_b,
                          cudaStream_t stream) {
    return sort_pairs_impl<uint64_t>(keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes,
                                     result_in_b, stream);
}

int gsr_launch_sort_pairs32(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t n,
                            int begin_bit, int end_bit, void* temp, size_t temp_bytes, int* result_in_b,
                            cudaStream_t stream, int site, bool hist_ready) {
    return sort_pairs_impl<uint32_t>(keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes,
                                     result_in_b, stream, site, hist_ready);
}
