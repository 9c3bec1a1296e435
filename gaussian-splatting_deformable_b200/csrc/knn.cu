// distCUDA2 replacement: mean squared distance to the 3 nearest other points.
//
// Replaces submodules/simple-knn (simple_knn.cu:185-221, spatial.cu:15-26): Morton
// sort + 1024-point boxes + pruned exhaustive scan.  Same EXACT answer (the
// reference's search is exact too), different machinery: a uniform grid with
// ~2 points per cell, points bucketed with the library's own onesweep sort, and a
// per-point expanding-shell search that stops as soon as the third-best distance
// is inside the searched cube.  No host synchronisation, no allocation (the
// reference cudaMalloc/cudaFree's and copies the bbox to the host twice per call).
#include "kernels.cuh"
#include <float.h>

namespace {

struct KnnGrid { float minv[3]; float inv_cell[3]; float cell[3]; int G; };

__device__ __forceinline__ unsigned f2ord(float f) {     // order-preserving float -> uint
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void knn_init_kernel(unsigned* bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;          // min
    else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;              // max
}

__global__ void __launch_bounds__(256) knn_bbox_kernel(int P, const float* __restrict__ pts, unsigned* bbox) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < P; i += gridDim.x * 256) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)i + k];
            mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&bbox[k], f2ord(mn[k])); atomicMax(&bbox[3 + k], f2ord(mx[k])); }
    }
}

__global__ void knn_grid_kernel(const unsigned* bbox, int G, KnnGrid* grid) {
    if (threadIdx.x == 0) {
        grid->G = G;
        for (int k = 0; k < 3; k++) {
            const float lo = ord2f(bbox[k]), hi = ord2f(bbox[3 + k]);
            float ext = hi - lo;
            if (!(ext > 0.0f)) ext = 1.0f;
            const float c = ext / (float)G;
            grid->minv[k] = lo; grid->cell[k] = c; grid->inv_cell[k] = 1.0f / c;
        }
    }
}

__device__ __forceinline__ int cell_of(float v, float lo, float inv, int G) {
    int c = (int)floorf((v - lo) * inv);
    return min(max(c, 0), G - 1);
}

__global__ void __launch_bounds__(256) knn_keys_kernel(int P, const float* __restrict__ pts, const KnnGrid* grid,
                                                       uint64_t* keys, uint32_t* vals) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const KnnGrid g = *grid;
    const int cx = cell_of(pts[3 * (size_t)i], g.minv[0], g.inv_cell[0], g.G);
    const int cy = cell_of(pts[3 * (size_t)i + 1], g.minv[1], g.inv_cell[1], g.G);
    const int cz = cell_of(pts[3 * (size_t)i + 2], g.minv[2], g.inv_cell[2], g.G);
    keys[i] = (uint64_t)(((uint32_t)cz * g.G + cy) * g.G + cx);
    vals[i] = i;
}

// cell_start[c] = first sorted slot of cell c; cell_start[cells] = P.
__global__ void __launch_bounds__(256) knn_cells_kernel(int P, const uint64_t* __restrict__ keys,
                                                        const uint32_t* __restrict__ vals,
                                                        const float* __restrict__ pts, uint32_t* cell_start,
                                                        uint32_t cells, float4* sorted_pts) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const uint32_t c = (uint32_t)keys[i];
    const uint32_t prev = (i == 0) ? 0u : (uint32_t)keys[i - 1] + 1u;   // first cell not yet started
    if (i == 0 || c + 1u != prev) {
        for (uint32_t k = (i == 0 ? 0u : prev); k <= c; k++) cell_start[k] = i;
    }
    if (i == P - 1) for (uint32_t k = c + 1; k <= cells; k++) cell_start[k] = P;
    const uint32_t id = vals[i];
    sorted_pts[i] = make_float4(pts[3 * (size_t)id], pts[3 * (size_t)id + 1], pts[3 * (size_t)id + 2],
                                __uint_as_float(id));
}

__device__ __forceinline__ void k_best3(float d, float (&best)[3]) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (best[j] > d) { const float t = best[j]; best[j] = d; d = t; }
    }
}

__global__ void __launch_bounds__(128) knn_query_kernel(int P, const float4* __restrict__ sorted_pts,
                                                        const uint32_t* __restrict__ cell_start, const KnnGrid* grid,
                                                        float* __restrict__ out) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= P) return;
    const KnnGrid g = *grid;
    const float4 me = sorted_pts[i];
    const int G = g.G;
    const int cx = cell_of(me.x, g.minv[0], g.inv_cell[0], G);
    const int cy = cell_of(me.y, g.minv[1], g.inv_cell[1], G);
    const int cz = cell_of(me.z, g.minv[2], g.inv_cell[2], G);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int r = 0; r < G; r++) {
        const int x0 = max(cx - r, 0), x1 = min(cx + r, G - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, G - 1);
        const int z0 = max(cz - r, 0), z1 = min(cz + r, G - 1);
        for (int z = z0; z <= z1; z++) {
            const bool zface = (z == cz - r) || (z == cz + r);
            for (int y = y0; y <= y1; y++) {
                const bool yface = (y == cy - r) || (y == cy + r);
                const bool whole_row = zface || yface;
                // rows on a z/y face of the shell are scanned whole (one contiguous
                // run of cells); interior rows only contribute their two end cells.
                const int nseg = whole_row ? 1 : 2;
                for (int sgm = 0; sgm < nseg; sgm++) {
                    int xa, xb;
                    if (whole_row) { xa = x0; xb = x1; }
                    else if (sgm == 0) { xa = cx - r; xb = cx - r; if (xa < 0) continue; }
                    else { xa = cx + r; xb = cx + r; if (xb > G - 1 || r == 0) continue; }
                    const uint32_t row = ((uint32_t)z * G + y) * G;
                    const uint32_t s = cell_start[row + xa], e = cell_start[row + xb + 1];
                    for (uint32_t k = s; k < e; k++) {
                        if ((int)k == i) continue;
                        const float4 o = sorted_pts[k];
                        const float dx = o.x - me.x, dy = o.y - me.y, dz = o.z - me.z;
                        k_best3(dx * dx + dy * dy + dz * dz, best);
                    }
                }
            }
        }
        // Everything within `reach` of the point has been examined.
        float reach = FLT_MAX;
        const float lo[3] = {g.minv[0] + (cx - r) * g.cell[0], g.minv[1] + (cy - r) * g.cell[1], g.minv[2] + (cz - r) * g.cell[2]};
        const float hi[3] = {g.minv[0] + (cx + r + 1) * g.cell[0], g.minv[1] + (cy + r + 1) * g.cell[1], g.minv[2] + (cz + r + 1) * g.cell[2]};
        const float pv[3] = {me.x, me.y, me.z};
        const int cc[3] = {cx, cy, cz};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (cc[k] - r > 0) reach = fminf(reach, pv[k] - lo[k]);          // a face with cells beyond it
            if (cc[k] + r < G - 1) reach = fminf(reach, hi[k] - pv[k]);
        }
        if (reach == FLT_MAX) break;                                          // whole grid searched
        reach = fmaxf(reach, 0.0f) * 0.9999f;                                 // guard the cell-boundary rounding
        if (best[2] <= reach * reach) break;
    }
    out[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

inline int knn_grid_dim(int P) {
    int G = (int)floor(cbrt((double)P / 2.0));
    if (G < 1) G = 1;
    if (G > 512) G = 512;
    return G;
}
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// temp: bbox[8 u32] | grid | keys_a | keys_b | vals_a | vals_b | cell_start | sorted_pts | sort temp
size_t gsr_knn_temp_bytes(int P) {
    const int G = knn_grid_dim(P);
    const size_t cells = (size_t)G * G * G;
    size_t b = 256 + 256;
    b += 2 * align256(sizeof(uint64_t) * (size_t)P) + 2 * align256(sizeof(uint32_t) * (size_t)P);
    b += align256(sizeof(uint32_t) * (cells + 1)) + align256(sizeof(float4) * (size_t)P);
    b += align256(gsr_sort_temp_bytes((uint32_t)P, 0, 32));
    return b;
}

int gsr_launch_knn_dist2(int P, const float* points, float* mean_dist2, void* temp, size_t temp_bytes,
                         cudaStream_t stream) {
    if (P <= 0) return 0;
    if (temp_bytes < gsr_knn_temp_bytes(P)) return gsr_set_error_msg(-3, "knn: temp buffer too small");
    const int G = knn_grid_dim(P);
    const uint32_t cells = (uint32_t)G * G * G;
    char* p = reinterpret_cast<char*>(temp);
    unsigned* bbox = reinterpret_cast<unsigned*>(p); p += 256;
    KnnGrid* grid = reinterpret_cast<KnnGrid*>(p); p += 256;
    uint64_t* keys_a = reinterpret_cast<uint64_t*>(p); p += align256(sizeof(uint64_t) * (size_t)P);
    uint64_t* keys_b = reinterpret_cast<uint64_t*>(p); p += align256(sizeof(uint64_t) * (size_t)P);
    uint32_t* vals_a = reinterpret_cast<uint32_t*>(p); p += align256(sizeof(uint32_t) * (size_t)P);
    uint32_t* vals_b = reinterpret_cast<uint32_t*>(p); p += align256(sizeof(uint32_t) * (size_t)P);
    uint32_t* cell_start = reinterpret_cast<uint32_t*>(p); p += align256(sizeof(uint32_t) * ((size_t)cells + 1));
    float4* sorted_pts = reinterpret_cast<float4*>(p); p += align256(sizeof(float4) * (size_t)P);
    void* sort_temp = p;
    int bits = 1;
    while ((1ull << bits) < (unsigned long long)cells) bits++;

    { GsrProfScope prof_("knn_init", stream);
    knn_init_kernel<<<1, 32, 0, stream>>>(bbox); }
    int bb = gsr_div_up(P, 256); if (bb > 148 * 8) bb = 148 * 8;
    { GsrProfScope prof_("knn_bbox", stream);
    knn_bbox_kernel<<<bb, 256, 0, stream>>>(P, points, bbox); }
    { GsrProfScope prof_("knn_grid", stream);
    knn_grid_kernel<<<1, 32, 0, stream>>>(bbox, G, grid); }
    { GsrProfScope prof_("knn_keys", stream);
    knn_keys_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, points, grid, keys_a, vals_a); }
    GSR_CHECK_LAUNCH();
    int in_b = 0;
    int rc = gsr_launch_sort_pairs(keys_a, keys_b, vals_a, vals_b, (uint32_t)P, 0, bits,
                                   sort_temp, gsr_sort_temp_bytes((uint32_t)P, 0, 32), &in_b, stream);
    if (rc) return rc;
    const uint64_t* sk = in_b ? keys_b : keys_a;
    const uint32_t* sv = in_b ? vals_b : vals_a;
    { GsrProfScope prof_("knn_cells", stream);
    knn_cells_kernel<<<gsr_div_up(P, 256), 256, 0, stream>>>(P, sk, sv, points, cell_start, cells, sorted_pts); }
    { GsrProfScope prof_("knn_query", stream);
    knn_query_kernel<<<gsr_div_up(P, 128), 128, 0, stream>>>(P, sorted_pts, cell_start, grid, mean_dist2); }
    GSR_CHECK_LAUNCH();
    return 0;
}
