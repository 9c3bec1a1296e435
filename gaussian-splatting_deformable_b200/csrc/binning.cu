// Binning: offsets scan, (tile|depth) key duplication, tile ranges.
//
// Replaces rasterizer_impl.cu:70-138,277-318 of the reference:
//   cub::DeviceScan::InclusiveSum      -> block sums (fused in preprocess) + one
//                                         single-block scan + an in-block scan
//                                         fused into the duplication kernel
//   duplicateWithKeys                  -> block-cooperative, load-balanced emit
//   cudaMemset + identifyTileRanges    -> same semantics
// All integer/bit work: results are bit-exact with the reference by construction
// (same key layout, same emission order: Gaussian index major, then y, then x).
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
// duplicateWithKeys, block-cooperative.  A block owns 256 consecutive Gaussians.
// It scans their tile counts in shared memory, then ALL 256 threads walk the
// block's output range: entry e belongs to the Gaussian found by binary search
// in the scanned counts, at rect cell k = e - start.  Every thread emits the same
// number of entries (+-1) whatever the splat sizes, and consecutive threads write
// consecutive 8-byte keys / 4-byte values (fully coalesced), where the reference
// has one thread serially emitting a whole rect.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) duplicate_kernel(int P, const int* __restrict__ radii,
                                                        const float* __restrict__ depths,
                                                        const uint32_t* __restrict__ tiles_touched,
                                                        const float4* __restrict__ recs,
                                                        const uint32_t* __restrict__ block_offsets,
                                                        uint32_t* __restrict__ point_offsets,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                        int grid_x, int grid_y) {
    __shared__ uint32_t s_start[257];   // exclusive scan of tile counts (+ total)
    __shared__ uint32_t s_rect[256];    // rmin.x | rmin.y << 12 | width << 24 (grid dims < 4096, width < 256)
    __shared__ uint32_t s_wide[256];    // width for very wide rects (>= 256 tiles across)
    __shared__ uint32_t s_depth[256];
    __shared__ uint32_t warp_tot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int idx = blockIdx.x * 256 + tid;
    uint32_t cnt = 0;
    if (idx < P) {
        cnt = tiles_touched[idx];
        if (cnt > 0) {
            const float4 r0 = recs[3 * (size_t)idx];
            uint2 rmin, rmax;
            tile_rect_exact(make_float2(r0.x, r0.y), radii[idx], grid_x, grid_y, rmin, rmax);
            s_rect[tid] = rmin.x | (rmin.y << 16);
            s_wide[tid] = rmax.x - rmin.x;
            s_depth[tid] = __float_as_uint(depths[idx]);
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
    if (point_offsets && idx < P) point_offsets[idx] = base + excl + cnt;  // inclusive, as the reference's scan
    __syncthreads();
    const uint32_t total = s_start[256];
    for (uint32_t e = tid; e < total; e += 256) {
        // largest g with s_start[g] <= e  (empty Gaussians share a start; pick the last one = the owner)
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
        uint64_t key = (uint64_t)(y * (uint32_t)grid_x + x);
        key = (key << 32) | (uint64_t)s_depth[lo];
        keys[base + e] = key;
        vals[base + e] = (uint32_t)(blockIdx.x * 256 + lo);
    }
}

int gsr_launch_duplicate(int P, const int* radii, const float* depths, const uint32_t* tiles_touched,
                         const float4* recs, const uint32_t* block_offsets, uint32_t* point_offsets,
                         uint64_t* keys, uint32_t* vals, int grid_x, int grid_y, cudaStream_t stream) {
    if (P <= 0) return 0;
    { GsrProfScope prof_("duplicate_with_keys", stream);
    duplicate_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, radii, depths, tiles_touched, recs, block_offsets,
                                                             point_offsets, keys, vals, grid_x, grid_y); }
    GSR_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------
// identifyTileRanges: ranges[tile] = [first, last+1) in the sorted list; tiles
// without entries keep (0,0) from the memset, exactly as the reference.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_ranges_kernel(uint32_t L, const uint64_t* __restrict__ keys,
                                                          uint2* __restrict__ ranges) {
    const uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

int gsr_launch_tile_ranges(uint32_t R, const uint64_t* sorted_keys, uint2* ranges, int num_tiles,
                           cudaStream_t stream) {
    GSR_CHECK(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)num_tiles, stream));
    if (R == 0) return 0;
    { GsrProfScope prof_("tile_ranges", stream);
    tile_ranges_kernel<<<gsr_div_up(R, 256), 256, 0, stream>>>(R, sorted_keys, ranges); }
    GSR_CHECK_LAUNCH();
    return 0;
}
