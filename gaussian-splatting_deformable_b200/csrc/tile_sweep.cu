// Tile binning as ONE stable counting sort over the depth-ordered Gaussians.
//
// Replaces, like binning.cu + radix_sort.cu, rasterizer_impl.cu:70-138,277-318 of the
// reference (duplicateWithKeys, cub::DeviceRadixSort::SortPairs on (tile | depth) keys,
// identifyTileRanges) - and produces the identical point_list / ranges - but never
// materialises or sorts the R (tile, Gaussian) duplicates:
//
//   the P Gaussians are already in depth order (depth_sort.cu), so the
//   reference's final order is "for every tile, the Gaussians that cover it, in depth
//   order" = a STABLE partition of the duplicate stream by tile id.  A stable partition
//   is a counting sort:
//     up-sweep    count[chunk][tile]   (chunk = G consecutive Gaussians of the depth order)
//     scan        start[chunk][tile] = base[tile] + sum_{c' < chunk} count[c'][tile]   (tile_column_scan_kernel;
//                 its last CTA also scans the tile totals)
//                 (base = exclusive scan of the tile totals = the reference's `ranges`)
//     down-sweep  every duplicate goes straight to point_list[start + rank-in-chunk]
//   HBM traffic: 16 B x P (sorted rect records, read twice) + the chunk x tile matrix
//   (3 passes) + 4 B x R written once.  The radix path moved 8 B x R + 2 x 16 B x R + 4 B x R.
//
// Work decomposition: one WARP per (chunk, stripe of 4 tile rows) with a warp-PRIVATE counter
// array in shared memory; see tile_sweep_kernel.
#include "kernels.cuh"
#include <stdlib.h>

namespace {

constexpr unsigned FULL = 0xffffffffu;

// srec[i] = (Gaussian id, rect lo, rect hi, 0) of the i-th Gaussian in depth order, written by the depth sort
// (depth_sort.cu): one coalesced LDG.128 per lane in the sweeps instead of an id load + dependent gather.
// *n_emit = number of Gaussians that emit at least one duplicate = entries of srec.

// One warp = (chunk of the depth order, stripe of GSR_SWEEP_ROWS tile rows).  Lane l maps to
// cell (row l / CW, column l % CW) of a Gaussian's clipped rect, CW = 32 / GSR_SWEEP_ROWS columns
// at a time: the cells of ONE Gaussian are distinct tiles, so a plain returning shared atomic per
// lane ranks them, and Gaussians are taken one after the other in program (= depth) order - which
// is all stability needs.  No scan, no search, no match, no CTA barrier.
template <bool SCATTER>
__global__ void __launch_bounds__(32 * GSR_SWEEP_WARPS) tile_sweep_kernel(const uint32_t* __restrict__ n_emit_p,
                                                                         const uint4* __restrict__ srec, GsrTileBinPlan pl,
                                                                         int grid_x, int grid_y,
                                                                         uint32_t* __restrict__ matrix,
                                                                         const uint32_t* __restrict__ tile_base,
                                                                         uint32_t* __restrict__ point_list,
                                                                         const uint32_t* __restrict__ overflow) {
    constexpr int ROWS = GSR_SWEEP_ROWS, CW = 32 / ROWS;
    extern __shared__ uint32_t s_cnt_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stripe = blockIdx.y * GSR_SWEEP_WARPS + warp;
    if (SCATTER && overflow && *overflow) return;           // capacity mode: the list does not fit
    if (stripe >= pl.stripes) return;                       // warps are independent: no CTA barrier below
    uint32_t* cnt = s_cnt_all + warp * pl.stripe_tiles;
    const int row0 = stripe * ROWS;
    const int row1 = min(grid_y, row0 + ROWS);
    const int tile0 = row0 * grid_x, ntile = (row1 - row0) * grid_x;
    const int chunk = blockIdx.x;
    const uint32_t n_emit = *n_emit_p;
    const uint32_t G = (((n_emit + pl.chunks - 1) / pl.chunks) + 31u) & ~31u;   // Gaussians per chunk
    const uint32_t g_begin = chunk * G, g_end = min(n_emit, g_begin + G);
    uint32_t* mrow = matrix + (size_t)chunk * pl.num_tiles + tile0;
    if (g_begin >= g_end) {
        if (!SCATTER) for (int i = lane; i < ntile; i += 32) mrow[i] = 0u;
        return;
    }
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (g_begin + lane < g_end) cur = srec[g_begin + lane];
    if (SCATTER) { for (int i = lane; i < ntile; i += 32) cnt[i] = tile_base[tile0 + i] + mrow[i]; }
    else { for (int i = lane; i < ntile; i += 32) cnt[i] = 0u; }
    __syncwarp();

    const int lrow = lane / CW, lcol = lane % CW;
    for (uint32_t b = g_begin; b < g_end; b += 32) {
        const uint4 rec = cur;
        if (b + 32 + lane < g_end) cur = srec[b + 32 + lane];      // prefetch the next batch
        else cur = make_uint4(0, 0, 0, 0);
        const int x0 = rec.y & 0xffffu, y0 = rec.y >> 16, x1 = rec.z & 0xffffu, y1 = rec.z >> 16;
        const int cy0 = max(y0, row0), cy1 = min(y1, row1);
        const bool has = (x1 > x0) && (cy1 > cy0);
        // per-lane description of the clipped rect: first stripe-local cell, width, height
        const uint32_t first = (uint32_t)((cy0 - row0) * grid_x + x0);
        const uint32_t wh = (uint32_t)(x1 - x0) | ((uint32_t)(cy1 - cy0) << 16);
        uint32_t m = __ballot_sync(FULL, has);
        while (m) {
            const int g = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t gfirst = __shfl_sync(FULL, first, g);
            const uint32_t gwh = __shfl_sync(FULL, wh, g);
            const uint32_t gid = __shfl_sync(FULL, rec.x, g);
            const int gw = (int)(gwh & 0xffffu), gh = (int)(gwh >> 16);
            if (lrow < gh) {
                const uint32_t rbase = gfirst + (uint32_t)(lrow * grid_x);
                for (int c = lcol; c < gw; c += CW) {
                    if (SCATTER) {
                        const uint32_t pos = atomicAdd(&cnt[rbase + c], 1u);
                        point_list[pos] = gid;
                    } else {
                        atomicAdd(&cnt[rbase + c], 1u);
                    }
                }
            }
        }
    }
    if (!SCATTER) {
        __syncwarp();
        for (int i = lane; i < ntile; i += 32) mrow[i] = cnt[i];
    }
}

// Up-sweep, order-free variant: counting does not care in which order the Gaussians of a chunk
// are visited, so one THREAD takes a Gaussian and bumps a CTA-wide counter array covering the whole
// tile grid (4 B x tiles of shared memory; 32 KB at 1080p, 127 KB at 4K).  The warp-per-stripe sweep
// above spends 17 warps re-scanning every chunk with 7 of 32 lanes busy; this does the same
// counting in a tenth of the instructions.  The rank-producing down-sweep keeps the ordered sweep.
__global__ void __launch_bounds__(1024) tile_count_kernel(const uint32_t* __restrict__ n_emit_p,
                                                          const uint4* __restrict__ srec, GsrTileBinPlan pl, int grid_x,
                                                          uint32_t* __restrict__ matrix) {
    extern __shared__ uint32_t s_cnt_all[];
    const int chunk = blockIdx.x;
    for (int i = threadIdx.x; i < pl.num_tiles; i += blockDim.x) s_cnt_all[i] = 0u;
    __syncthreads();
    const uint32_t n_emit = *n_emit_p;
    const uint32_t G = (((n_emit + pl.chunks - 1) / pl.chunks) + 31u) & ~31u;   // same chunking as the sweep
    const uint32_t g_begin = chunk * G, g_end = min(n_emit, g_begin + G);
    for (uint32_t g = g_begin + threadIdx.x; g < g_end; g += blockDim.x) {
        const uint4 rec = srec[g];
        const uint32_t x0 = rec.y & 0xffffu, y0 = rec.y >> 16, x1 = rec.z & 0xffffu, y1 = rec.z >> 16;
        for (uint32_t y = y0; y < y1; y++) {
            uint32_t* row = s_cnt_all + y * (uint32_t)grid_x;
            for (uint32_t x = x0; x < x1; x++) atomicAdd(row + x, 1u);
        }
    }
    __syncthreads();
    uint32_t* mrow = matrix + (size_t)chunk * pl.num_tiles;
    for (int i = threadIdx.x; i < pl.num_tiles; i += blockDim.x) mrow[i] = s_cnt_all[i];
}

// Down-sweep, second generation.  CTA = (chunk, group of GSR_SWEEP_WARPS stripes).  The chunk is taken
// in sub-batches of GSR_SCATTER_SUB Gaussians: (1) all warps stage the sub-batch's rect records in
// shared memory and, with one ballot per stripe, write a bitmap "Gaussian i touches stripe s";
// (2) warp s walks the set bits of ITS bitmap in order (= depth order) and ranks/scatters the cells of
// each Gaussian with one returning shared atomic per lane, as in tile_sweep_kernel.  A stripe warp
// thus only ever visits the Gaussians that reach its stripe (1 in 9 at 1080p) instead of scanning
// the whole chunk, and the rect arrives by one broadcast LDS.128 instead of three shuffles.
#define GSR_SCATTER_SUB 1024
__global__ void __launch_bounds__(32 * GSR_SCATTER_MAX_WARPS) tile_scatter_kernel(const uint32_t* __restrict__ n_emit_p,
                                                                           const uint4* __restrict__ srec, GsrTileBinPlan pl,
                                                                           int grid_x, int grid_y,
                                                                           const uint32_t* __restrict__ matrix,
                                                                           const uint32_t* __restrict__ tile_base,
                                                                           uint32_t* __restrict__ point_list,
                                                                           const uint32_t* __restrict__ overflow) {
    constexpr int ROWS = GSR_SWEEP_ROWS, CW = 32 / ROWS;
    extern __shared__ uint32_t s_cnt_all[];                                   // [warps][stripe_tiles]
    __shared__ uint4 s_rec[GSR_SCATTER_SUB];
    if (overflow && *overflow) return;                                        // capacity mode: the list does not fit
    __shared__ uint32_t s_bits[GSR_SCATTER_MAX_WARPS][GSR_SCATTER_SUB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NWARPS = pl.scatter_warps;                                       // = blockDim.x / 32
    const int stripe = blockIdx.y * NWARPS + warp;
    const bool live = stripe < pl.stripes;
    uint32_t* cnt = s_cnt_all + warp * pl.stripe_tiles;
    const int row0 = stripe * ROWS;
    const int row1 = min(grid_y, row0 + ROWS);
    const int tile0 = row0 * grid_x, ntile = live ? (row1 - row0) * grid_x : 0;
    const int chunk = blockIdx.x;
    const uint32_t n_emit = *n_emit_p;
    const uint32_t G = (((n_emit + pl.chunks - 1) / pl.chunks) + 31u) & ~31u;
    const uint32_t g_begin = chunk * G, g_end = min(n_emit, g_begin + G);
    if (g_begin >= g_end) return;
    const uint32_t* mrow = matrix + (size_t)chunk * pl.num_tiles + tile0;
    for (int i = lane; i < ntile; i += 32) cnt[i] = tile_base[tile0 + i] + mrow[i];
    const int grow0 = blockIdx.y * NWARPS * ROWS;                             // first tile row of this CTA's stripes
    const int lrow = lane / CW, lcol = lane % CW;
    uint32_t* const cnt_lane = cnt + (lrow - row0) * grid_x + lcol;

    for (uint32_t sb = g_begin; sb < g_end; sb += GSR_SCATTER_SUB) {
        const uint32_t sn = min((uint32_t)GSR_SCATTER_SUB, g_end - sb);
        __syncthreads();                                                      // previous sub-batch fully consumed
        // (1) stage + bitmaps: warp w takes Gaussians w*32 + lane, + 256, ...
        for (uint32_t i0 = warp * 32; i0 < sn; i0 += 32 * NWARPS) {
            const uint32_t i = i0 + lane;
            uint4 rec = make_uint4(0, 0, 0, 0);
            if (i < sn) rec = srec[sb + i];
            s_rec[i0 + lane] = rec;
            const int y0 = rec.y >> 16, y1 = rec.z >> 16;
            const bool any = (rec.z & 0xffffu) > (rec.y & 0xffffu);
            for (int s = 0; s < NWARPS; s++) {
                const int r0 = grow0 + s * ROWS;
                const unsigned m = __ballot_sync(FULL, any && y0 < r0 + ROWS && y1 > r0);
                if (lane == 0) s_bits[s][i0 >> 5] = m;
            }
        }
        __syncthreads();
        if (!live) continue;
        // (2) ordered walk over this stripe's Gaussians
        const int nwords = (int)((sn + 31) >> 5);
        for (int k = 0; k < nwords; k++) {
            uint32_t m = s_bits[warp][k];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const uint4 rec = s_rec[32 * k + b];
                const int x0 = rec.y & 0xffffu, y0 = rec.y >> 16, x1 = rec.z & 0xffffu, y1 = rec.z >> 16;
                const int cy0 = max(y0, row0), cy1 = min(y1, row1);
                const int gw = x1 - x0, gh = cy1 - cy0;
                const bool rowok = lrow < gh;
                uint32_t* cell = cnt_lane + cy0 * grid_x + x0;          // this lane's cell of the first 4 x 8 window
                if (rowok && lcol < gw) {
                    const uint32_t pos = atomicAdd(cell, 1u);
                    point_list[pos] = rec.x;
                }
                if (gw > CW) {                                           // wider than the window: rare, warp-uniform
#pragma unroll 1
                    for (int c = CW; c < gw; c += CW) {
                        if (rowok && lcol + c < gw) {
                            const uint32_t pos = atomicAdd(cell + c, 1u);
                            point_list[pos] = rec.x;
                        }
                    }
                }
            }
        }
    }
}

// Column scan of the chunk x tile count matrix: in place, counts -> exclusive prefix over the
// chunks; totals[tile] = column sum.  CTA = 32 tiles x 8 row segments: every thread sums its
// segment, the 8 partial sums are scanned in shared memory, then the segment is rewritten.
__global__ void __launch_bounds__(256) tile_column_scan_kernel(int chunks, int num_tiles, uint32_t* __restrict__ matrix,
                                                               uint32_t* __restrict__ totals, uint32_t* __restrict__ ticket,
                                                               uint32_t* __restrict__ base, uint2* __restrict__ ranges,
                                                               uint32_t capacity, uint32_t* __restrict__ overflow) {
    __shared__ uint32_t s_part[8][33];
    __shared__ bool s_last;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + tx;
    const int seg = (chunks + 7) / 8;
    const int r0 = ty * seg, r1 = min(chunks, r0 + seg);
    uint32_t sum = 0;
    if (t < num_tiles) {
        int c = r0;
        for (; c + 8 <= r1; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = matrix[(size_t)(c + k) * num_tiles + t];
#pragma unroll
            for (int k = 0; k < 8; k++) sum += v[k];
        }
        for (; c < r1; c++) sum += matrix[(size_t)c * num_tiles + t];
    }
    s_part[ty][tx] = sum;
    __syncthreads();
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) run += (k < ty) ? s_part[k][tx] : 0u;
    if (t < num_tiles) {
        int c = r0;
        for (; c + 8 <= r1; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = matrix[(size_t)(c + k) * num_tiles + t];
#pragma unroll
            for (int k = 0; k < 8; k++) { matrix[(size_t)(c + k) * num_tiles + t] = run; run += v[k]; }
        }
        for (; c < r1; c++) {
            const uint32_t v = matrix[(size_t)c * num_tiles + t];
            matrix[(size_t)c * num_tiles + t] = run;
            run += v;
        }
        if (ty == 7) totals[t] = run;
    }
    // The CTA that finishes last scans the tile totals: base[tile] and the reference's `ranges`
    // (identifyTileRanges + its memset, rasterizer_impl.cu:116-138,310) without another launch.
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) + 1u == gridDim.x);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0u;                     // ready for the next call
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    // 2048 tiles per round: a thread owns 8 consecutive tiles (two 16-byte L2 loads), block scan with a carry
    for (int c0 = 0; c0 < num_tiles; c0 += 2048) {
        const int i0 = c0 + threadIdx.x * 8;
        uint32_t v[8];
        if (i0 + 8 <= num_tiles && (num_tiles & 3) == 0) {
            const uint4 a = __ldcg(reinterpret_cast<const uint4*>(totals + i0));
            const uint4 b = __ldcg(reinterpret_cast<const uint4*>(totals + i0) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (i0 + k < num_tiles) ? __ldcg(totals + i0 + k) : 0u;
        }
        uint32_t sum8 = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum8 += v[k];
        uint32_t incl = sum8;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t wbase = s_carry, all = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { wbase += (w < warp) ? warp_tot[w] : 0u; all += warp_tot[w]; }
        uint32_t ex = wbase + incl - sum8;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (i0 + k < num_tiles) {
                base[i0 + k] = ex;
                ranges[i0 + k] = v[k] ? make_uint2(ex, ex + v[k]) : make_uint2(0u, 0u);
            }
            ex += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += all;
        __syncthreads();
    }
    // Sync-free forward (capacity mode): the list was sized from an estimate.  If this view emits more duplicates than
    // the workspace holds, raise the flag (the scatter returns at once, the host reads it with the view's counters) and
    // empty every tile, so that nothing downstream reads a list entry that was never written.
    if (capacity != 0u && s_carry > capacity) {
        if (threadIdx.x == 0) *overflow = 1u;
        for (int i = threadIdx.x; i < num_tiles; i += 256) ranges[i] = make_uint2(0u, 0u);
    }
}

// Parity/debug: the sorted tile ids the reference's keys carry, rebuilt from the ranges.
__global__ void __launch_bounds__(256) expand_tile_ids_kernel(int num_tiles, const uint2* __restrict__ ranges,
                                                              uint32_t* __restrict__ tile_ids) {
    const int t = blockIdx.x;
    if (t >= num_tiles) return;
    const uint2 r = ranges[t];
    for (uint32_t i = r.x + threadIdx.x; i < r.y; i += 256) tile_ids[i] = (uint32_t)t;
}

int env_int2(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

// Plan for a grid_x x grid_y tile grid; feasible == 0 when a stripe of tile rows does not fit a
// warp's counter array (image wider than 16 x GSR_SWEEP_MAX_STRIPE_TILES / GSR_SWEEP_ROWS pixels).
GsrTileBinPlan gsr_make_tile_bin_plan(int grid_x, int grid_y) {
    GsrTileBinPlan pl{};
    pl.num_tiles = grid_x * grid_y;
    static const int forced = env_int2("GSR_BINNING_RADIX", 0);
    if (forced || grid_x <= 0 || grid_y <= 0 || grid_x * GSR_SWEEP_ROWS > GSR_SWEEP_MAX_STRIPE_TILES) { pl.feasible = 0; return pl; }
    pl.feasible = 1;
    static const int c_env = env_int2("GSR_SWEEP_CHUNKS", 0);
    // 512 chunks: alone the three kernels are 7 % faster with 768 (more CTAs), but with several views in flight the smaller
    // chunk x tile matrix wins (C2 1.094 vs 1.102 ms/view, C3 0.591 vs 0.605; one view at a time at C4: 5.74 vs 5.72)
    pl.chunks = (c_env > 0 && c_env <= GSR_SWEEP_MAX_CHUNKS) ? c_env : 512;
    pl.stripes = (grid_y + GSR_SWEEP_ROWS - 1) / GSR_SWEEP_ROWS;
    pl.groups = (pl.stripes + GSR_SWEEP_WARPS - 1) / GSR_SWEEP_WARPS;
    pl.stripe_tiles = GSR_SWEEP_ROWS * grid_x;
    // tile_scatter: every CTA of a chunk stages the chunk's records, and the stripes are spread EVENLY over the CTAs of a
    // chunk (1080p, 17 stripes: 8 + 8 + 1 warps - the third CTA with one live warp - 83 us; 6 + 6 + 5: 77 us;
    // 9 + 8: 74 us; one CTA of 17 warps: 87 us).  At most 12 warps per CTA unless GSR_SCATTER_WARPS says otherwise.
    static const int w_env = env_int2("GSR_SCATTER_WARPS", 12);
    int wmax = (160 * 1024) / (pl.stripe_tiles * (int)sizeof(uint32_t));
    if (wmax > GSR_SCATTER_MAX_WARPS) wmax = GSR_SCATTER_MAX_WARPS;
    if (w_env > 0 && w_env < wmax) wmax = w_env;
    if (wmax < 1) wmax = 1;
    pl.scatter_groups = (pl.stripes + wmax - 1) / wmax;
    pl.scatter_warps = (pl.stripes + pl.scatter_groups - 1) / pl.scatter_groups;
    return pl;
}

size_t gsr_tile_matrix_bytes(int grid_x, int grid_y) {
    const GsrTileBinPlan pl = gsr_make_tile_bin_plan(grid_x, grid_y);
    return pl.feasible ? (size_t)pl.chunks * (size_t)pl.num_tiles * sizeof(uint32_t) : 0;
}

int gsr_launch_tile_binning(int P, const uint32_t* n_emit, const uint4* srec,
                            const GsrTileBinPlan& pl, int grid_x, int grid_y, uint32_t* matrix, uint32_t* totals,
                            uint32_t* tile_base, uint2* ranges, uint32_t* point_list, uint32_t* scan_ticket, cudaStream_t stream,
                            uint32_t capacity, uint32_t* overflow) {
    if (P <= 0) return 0;
    if (!pl.feasible) return gsr_set_error_msg(-2, "tile sweep: plan not feasible");
    const size_t smem = (size_t)GSR_SWEEP_WARPS * pl.stripe_tiles * sizeof(uint32_t);
    static bool attr_done = false;
    if (!attr_done) {
        GSR_CHECK(cudaFuncSetAttribute(tile_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GSR_SWEEP_WARPS * GSR_SWEEP_MAX_STRIPE_TILES * (int)sizeof(uint32_t)));
        GSR_CHECK(cudaFuncSetAttribute(tile_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GSR_SWEEP_WARPS * GSR_SWEEP_MAX_STRIPE_TILES * (int)sizeof(uint32_t)));
        attr_done = true;
    }
    const dim3 grid(pl.chunks, pl.groups, 1);
    static const int count_v = env_int2("GSR_SWEEP_COUNT_V", 2);
    const size_t cnt_smem = (size_t)pl.num_tiles * sizeof(uint32_t);
    if (count_v == 2 && cnt_smem <= 200 * 1024) {
        static bool cattr_done = false;
        if (!cattr_done) {
            GSR_CHECK(cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            cattr_done = true;
        }
        GsrProfScope prof_("tile_count", stream);
        tile_count_kernel<<<pl.chunks, cnt_smem <= 48 * 1024 ? 256 : 1024, cnt_smem, stream>>>(n_emit, srec, pl, grid_x, matrix);
    } else {
        GsrProfScope prof_("tile_sweep_count", stream);
        tile_sweep_kernel<false><<<grid, 32 * GSR_SWEEP_WARPS, smem, stream>>>(n_emit, srec, pl, grid_x, grid_y, matrix, tile_base, point_list, nullptr);
    }
    GSR_CHECK_LAUNCH();
    { GsrProfScope prof_("tile_column_scan", stream);
    tile_column_scan_kernel<<<gsr_div_up(pl.num_tiles, 32), 256, 0, stream>>>(pl.chunks, pl.num_tiles, matrix, totals, scan_ticket,
                                                                             tile_base, ranges, capacity, overflow); }
    GSR_CHECK_LAUNCH();
    static const int scatter_v = env_int2("GSR_SWEEP_SCATTER_V", 2);
    if (scatter_v == 2) {
        static bool sattr_done = false;
        if (!sattr_done) {
            GSR_CHECK(cudaFuncSetAttribute(tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024));
            sattr_done = true;
        }
        GsrProfScope prof_("tile_scatter", stream);
        const dim3 sgrid(pl.chunks, pl.scatter_groups, 1);
        const size_t ssmem = (size_t)pl.scatter_warps * pl.stripe_tiles * sizeof(uint32_t);
        tile_scatter_kernel<<<sgrid, 32 * pl.scatter_warps, ssmem, stream>>>(n_emit, srec, pl, grid_x, grid_y, matrix, tile_base, point_list, overflow);
    } else {
        GsrProfScope prof_("tile_sweep_scatter", stream);
        tile_sweep_kernel<true><<<grid, 32 * GSR_SWEEP_WARPS, smem, stream>>>(n_emit, srec, pl, grid_x, grid_y, matrix, tile_base, point_list, overflow);
    }
    GSR_CHECK_LAUNCH();
    return 0;
}

int gsr_launch_expand_tile_ids(int num_tiles, const uint2* ranges, uint32_t* tile_ids, cudaStream_t stream) {
    if (num_tiles <= 0) return 0;
    { GsrProfScope prof_("expand_tile_ids", stream);
    expand_tile_ids_kernel<<<num_tiles, 256, 0, stream>>>(num_tiles, ranges, tile_ids); }
    GSR_CHECK_LAUNCH();
    return 0;
}
