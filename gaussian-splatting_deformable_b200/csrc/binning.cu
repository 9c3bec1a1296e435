// Binning: offsets scan, tile-key duplication, tile ranges.
//
// Replaces rasterizer_impl.cu:70-138,277-318 of the reference:
//   cub::DeviceScan::InclusiveSum      -> block sums + one single-block scan + an
//                                         in-block scan fused into the duplication kernel
//   duplicateWithKeys                  -> block-cooperative, load-balanced emit
//   cudaMemset + identifyTileRanges    -> same semantics
// All integer/bit work: the sorted list, its keys and the tile ranges are bit-exact
// with the reference (see "Binning order" below for why the cheaper sort is the same
// permutation).
#include "geom_exact.cuh"
#include "kernels.cuh"

// ---------------------------------------------------------------------------
// Exclusive scan of <= a few 10^4 block sums by ONE block (1024 threads, chunked
// with a running carry).  Trivial traffic (4 B per 256 Gaussians).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(uint32_t* sums, int n, uint32_t* total) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < n) ? sums[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t warp_excl = warp ? warp_tot[warp - 1] : 0u;
        if (i < n) sums[i] = carry + warp_excl + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

int gsr_launch_scan_block_sums(uint32_t* block_sums, int num_blocks, uint32_t* d_total, cudaStream_t stream) {
    { GsrProfScope prof_("scan_block_sums", stream);
    scan_block_sums_kernel<<<1, 1024, 0, stream>>>(block_sums, num_blocks, d_total); }
    GSR_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------
// Binning order.  The reference sorts R (tile | depth) 64-bit keys (45 significant
// bits -> 6 digit passes over 12-byte pairs).  All duplicates of one Gaussian share
// the depth digits, so the same permutation is obtained much cheaper:
//   1. order the P Gaussians by (depth bits, index) once (depth_sort.cu;
//      Gaussians that emit nothing carry key 0xffffffff and sink to the end),
//   2. emit the duplicates walking the Gaussians in that order,
//   3. stable-sort the duplicates by TILE ID only (tile_bits <= 16 -> 2 passes over
//      8-byte pairs).
// Stability makes the result identical to the reference's: inside a tile entries end
// up ordered by (depth bits, Gaussian index), ties included.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sorted_block_sums_kernel(int P, const uint32_t* __restrict__ n_emit, const uint32_t* __restrict__ order,
                                                                const uint32_t* __restrict__ tiles_touched,
                                                                uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t wsum[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    uint32_t s = (i < P && (uint32_t)i < *n_emit) ? tiles_touched[order[i]] : 0u;   // order[] holds the emitting Gaussians only
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) t += wsum[k];
        block_sums[blockIdx.x] = t;
    }
}

int gsr_launch_sorted_block_sums(int P, const uint32_t* n_emit, const uint32_t* order, const uint32_t* tiles_touched,
                                 uint32_t* block_sums, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("sorted_block_sums", stream);
    sorted_block_sums_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, n_emit, order, tiles_touched, block_sums); }
    GSR_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------
// duplicateWithKeys, block-cooperative.  A block owns 256 consecutive Gaussians of the
// depth order.  It scans their tile counts in shared memory, then ALL 256 threads walk
// the block's output range: entry e belongs to the Gaussian found by binary search in
// the scanned counts, at rect cell k = e - start (y-major then x, as the reference).
// Every thread emits the same number of entries (+-1) whatever the splat sizes, and
// consecutive threads write consecutive words (fully coalesced), where the reference
// has one thread serially emitting a whole rect.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) duplicate_kernel(int P, const uint32_t* __restrict__ n_emit, const uint32_t* __restrict__ order,
                                                        const int* __restrict__ radii,
                                                        const uint32_t* __restrict__ tiles_touched,
                                                        const float4* __restrict__ recs,
                                                        const uint32_t* __restrict__ block_offsets,
                                                        uint32_t* __restrict__ tile_ids, uint32_t* __restrict__ vals,
                                                        int grid_x, int grid_y, GsrSortPlan plan,
                                                        uint32_t* __restrict__ tile_hist) {
    __shared__ uint32_t s_hist[2 * GSR_SORT_RADIX];   // tile-id digit histograms (tile sort has <= 2 passes up to 65536 tiles)
    __shared__ uint32_t s_start[257];   // exclusive scan of tile counts (+ total)
    __shared__ uint32_t s_rect[256];    // rmin.x | rmin.y << 16
    __shared__ uint32_t s_wide[256];    // rect width in tiles
    __shared__ uint32_t s_gid[256];
    __shared__ uint32_t warp_tot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * 256 + tid;
    for (int k = tid; k < 2 * GSR_SORT_RADIX; k += 256) s_hist[k] = 0;
    uint32_t cnt = 0;
    if (i < P && (uint32_t)i < *n_emit) {          // order[] holds the emitting Gaussians only
        const uint32_t g = order[i];
        cnt = tiles_touched[g];
        if (cnt > 0) {
            const float4 r0 = recs[3 * (size_t)g];
            uint2 rmin, rmax;
            tile_rect_exact(make_float2(r0.x, r0.y), radii[g], grid_x, grid_y, rmin, rmax);
            s_rect[tid] = rmin.x | (rmin.y << 16);
            s_wide[tid] = rmax.x - rmin.x;
            s_gid[tid] = g;
        }
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) wbase += (w < warp) ? warp_tot[w] : 0u;
    const uint32_t excl = wbase + incl - cnt;
    s_start[tid] = excl;
    if (tid == 255) s_start[256] = excl + cnt;
    const uint32_t base = block_offsets[blockIdx.x];
    __syncthreads();
    const uint32_t total = s_start[256];
    // warp-uniform trip count: the histogram below uses warp votes
    for (uint32_t e0 = 0; e0 < total; e0 += 256) {
        const uint32_t e = e0 + tid;
        const bool valid = e < total;
        uint32_t tile = 0;
        if (valid) {
            // largest g with s_start[g] <= e  (empty Gaussians share a start; the last one is the owner)
            int lo = 0, hi = 256;
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int mid = (lo + hi) >> 1;
                if (s_start[mid] <= e) lo = mid; else hi = mid;
            }
            const uint32_t k = e - s_start[lo];
            const uint32_t w = s_wide[lo];
            const uint32_t rx = s_rect[lo] & 0xffffu, ry = s_rect[lo] >> 16;
            const uint32_t y = ry + k / w, x = rx + k % w;
            tile = y * (uint32_t)grid_x + x;
            tile_ids[base + e] = tile;
            vals[base + e] = s_gid[lo];
        }
        // Digit histograms for the tile sort, fused here (the sort never re-reads the ids).
        // Low digit: neighbouring entries are neighbouring tiles -> distinct bins, plain atomics.
        // High digit: mostly one value per warp -> aggregate with match-any.
        if (valid) atomicAdd(&s_hist[(tile >> plan.shift[0]) & plan.mask[0]], 1u);
        if (plan.passes > 1) {
            const uint32_t d = valid ? ((tile >> plan.shift[1]) & plan.mask[1]) : 0xffffffffu;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_hist[GSR_SORT_RADIX + d], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int k = tid; k < 2 * GSR_SORT_RADIX; k += 256) {
        const uint32_t c = s_hist[k];
        if (c) atomicAdd(&tile_hist[k], c);
    }
}

int gsr_launch_duplicate(int P, const uint32_t* n_emit, const uint32_t* order, const int* radii, const uint32_t* tiles_touched,
                         const float4* recs, const uint32_t* block_offsets, uint32_t* tile_ids, uint32_t* vals,
                         int grid_x, int grid_y, GsrSortPlan tile_plan, uint32_t* tile_hist, cudaStream_t stream) {
    if (P <= 0) return 0;
    if (tile_plan.passes > 2) return gsr_set_error_msg(-2, "more than 65536 tiles are not supported");
    { GsrProfScope prof_("duplicate_with_keys", stream);
    duplicate_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, n_emit, order, radii, tiles_touched, recs, block_offsets,
                                                             tile_ids, vals, grid_x, grid_y, tile_plan, tile_hist); }
    GSR_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------
// identifyTileRanges: ranges[tile] = [first, last+1) in the sorted list; tiles
// without entries keep (0,0) from the memset, exactly as the reference.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_ranges_kernel(uint32_t L, const uint32_t* __restrict__ tile_ids,
                                                          uint2* __restrict__ ranges) {
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = tile_ids[idx];
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = tile_ids[idx - 1];
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

int gsr_launch_tile_ranges(uint32_t R, const uint32_t* sorted_tile_ids, uint2* ranges, int num_tiles,
                           cudaStream_t stream) {
    GSR_CHECK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
    if (R == 0) return 0;
    { GsrProfScope prof_("tile_ranges", stream);
    tile_ranges_kernel<<<gsr_div_up(R, 256), 256, 0, stream>>>(R, sorted_tile_ids, ranges); }
    GSR_CHECK_LAUNCH();
    return 0;
}

__global__ void __launch_bounds__(256) materialize_keys_kernel(uint32_t L, const uint32_t* __restrict__ tile_ids,
                                                               const uint32_t* __restrict__ point_list,
                                                               const float* __restrict__ depths,
                                                               uint64_t* __restrict__ keys64) {
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= L) return;
    keys64[idx] = ((uint64_t)tile_ids[idx] << 32) | (uint64_t)__float_as_uint(depths[point_list[idx]]);
}

int gsr_launch_materialize_keys(uint32_t R, const uint32_t* sorted_tile_ids, const uint32_t* point_list,
                                const float* depths, uint64_t* keys64, cudaStream_t stream) {
    if (R == 0) return 0;
    { GsrProfScope prof_("materialize_keys", stream);
    materialize_keys_kernel<<<gsr_div_up(R, 256), 256, 0, stream>>>(R, sorted_tile_ids, point_list, depths, keys64); }
    GSR_CHECK_LAUNCH();
    return 0;
}
