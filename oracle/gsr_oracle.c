/* gsr_oracle.c - CPU restatement of the reference's rasterizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the tests compare the
 * sm_100a kernels against; it is never linked into, imported by or called from
 * the product (gaussian-splatting_deformable_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's baseline legs may load it.
 *
 * Scalar, single-threaded, one plain loop per reference kernel, in the
 * reference's own AoS layouts.  Each function cites the reference code it
 * follows (paths relative to submodules/diff-gaussian-rasterization/).
 *
 * Pinning: tests/test_oracle_cpu.py checks every function here against
 * tests/golden/raster_golden.npz, which was produced by the reference's own CUDA
 * code (oracle/_ref, built unmodified by oracle/build_ref.py) on a B200 with
 * tests/golden/make_raster_golden.py.
 *
 * Arithmetic: FP32 throughout, like the reference.  The geometry chain (view /
 * projection transform, cov3D, cov2D, conic, radius, pixel centre, tile rect) uses
 * fmaf() in exactly the places where nvcc fused the reference's expressions (read
 * off its PTX/SASS, see csrc/geom_exact.cuh), so the integer stages (radii, tile
 * counts, keys, order, ranges) reproduce the reference bit for bit; expf() is the
 * host libm's, so colours/alphas agree to rounding, not bit-exactly.
 * Compile with -ffp-contract=off (oracle/build_oracle.py does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16
#define BLOCK_Y 16

typedef struct {
    int P, D, M;            /* Gaussians, active SH degree, SH coefficients per Gaussian */
    int W, H;
    float tan_fovx, tan_fovy, scale_modifier;
    float bg[3];
    float view[16];         /* viewmatrix, reference memory convention */
    float proj[16];         /* projmatrix */
    float campos[3];
} orc_view;

/* auxiliary.h:22-38 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

static int grid_x(const orc_view* v) { return (v->W + BLOCK_X - 1) / BLOCK_X; }
static int grid_y(const orc_view* v) { return (v->H + BLOCK_Y - 1) / BLOCK_Y; }

/* auxiliary.h:58-77 transformPoint4x3 / 4x4, one row; fused as the reference binary. */
static float xform_row(const float* m, int r, const float* p) {
    return m[12 + r] + fmaf(p[2], m[8 + r], fmaf(p[0], m[r], p[1] * m[4 + r]));
}
/* a*b + c*d + e*f as compiled for a glm mat3 product element. */
static float dot3g(float a, float b, float c, float d, float e, float f) { return fmaf(e, f, fmaf(a, b, c * d)); }

/* forward.cu:118-152 computeCov3D */
static void cov3d(const float* scale, float mod, const float* q, float* c6) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float sx = mod * scale[0], sy = mod * scale[1], sz = mod * scale[2];
    const float yy = y * y, zz = z * z, rz = r * z, xz = x * z, rx = r * x;
    const float R00 = 1.0f - ((yy + zz) + (yy + zz));
    const float t01 = fmaf(x, y, -rz), t02 = fmaf(r, y, xz), t10 = fmaf(x, y, rz);
    const float t12 = fmaf(y, z, -rx), t20 = fmaf(-r, y, xz), t21 = fmaf(y, z, rx);
    const float a11 = fmaf(x, x, zz), a22 = fmaf(x, x, yy);
    const float R01 = t01 + t01, R02 = t02 + t02, R10 = t10 + t10, R11 = 1.0f - (a11 + a11);
    const float R12 = t12 + t12, R20 = t20 + t20, R21 = t21 + t21, R22 = 1.0f - (a22 + a22);
    const float M00 = sx * R00, M01 = sy * R01, M02 = sz * R02;
    const float M10 = sx * R10, M11 = sy * R11, M12 = sz * R12;
    const float M20 = sx * R20, M21 = sy * R21, M22 = sz * R22;
    c6[0] = dot3g(M00, M00, M01, M01, M02, M02);
    c6[1] = dot3g(M10, M00, M11, M01, M12, M02);
    c6[2] = dot3g(M20, M00, M21, M01, M22, M02);
    c6[3] = dot3g(M10, M10, M11, M11, M12, M12);
    c6[4] = dot3g(M20, M10, M21, M11, M22, M12);
    c6[5] = dot3g(M20, M20, M21, M21, M22, M22);
}

/* forward.cu:74-113 computeCov2D; T = W*J kept for the backward. */
static void ewa_T(const orc_view* v, const float* t, float focal_x, float focal_y, float* T6) {
    const float limx = v->tan_fovx * 1.3f, limy = v->tan_fovy * 1.3f;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    const float cx = fminf(limx, fmaxf(-limx, txtz)), cy = fminf(limy, fmaxf(-limy, tytz));
    const float ntz = -t[2], tz2 = t[2] * t[2];
    const float J00 = focal_x / t[2], J02 = (focal_x * (cx * ntz)) / tz2;
    const float J11 = focal_y / t[2], J12 = (focal_y * (cy * ntz)) / tz2;
    const float* V = v->view;
    T6[0] = fmaf(V[2], J02, V[0] * J00);
    T6[1] = fmaf(V[6], J02, V[4] * J00);
    T6[2] = fmaf(V[10], J02, V[8] * J00);
    T6[3] = fmaf(V[2], J12, J11 * V[1]);
    T6[4] = fmaf(V[6], J12, J11 * V[5]);
    T6[5] = fmaf(V[10], J12, J11 * V[9]);
}
static void cov2d(const float* T, const float* c, float* abc) {
    const float X00 = dot3g(T[0], c[0], T[1], c[1], T[2], c[2]);
    const float X01 = dot3g(T[3], c[0], T[4], c[1], T[5], c[2]);
    const float X10 = dot3g(T[0], c[1], T[1], c[3], T[2], c[4]);
    const float X11 = dot3g(T[3], c[1], T[4], c[3], T[5], c[4]);
    const float X20 = dot3g(T[0], c[2], T[1], c[4], T[2], c[5]);
    const float X21 = dot3g(T[3], c[2], T[4], c[4], T[5], c[5]);
    abc[0] = dot3g(T[0], X00, T[1], X10, T[2], X20) + 0.3f;
    abc[1] = dot3g(T[0], X01, T[1], X11, T[2], X21);
    abc[2] = dot3g(T[3], X01, T[4], X11, T[5], X21) + 0.3f;
}

/* auxiliary.h:41-44 */
static float ndc2pix(float v, int S) { return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5); }

/* auxiliary.h:46-56 */
static void get_rect(const float* p, int max_radius, const orc_view* v, uint32_t* rmin, uint32_t* rmax) {
    const int gx = grid_x(v), gy = grid_y(v);
    const float rf = (float)max_radius;
    int a;
    a = (int)((p[0] - rf) * 0.0625f); if (a < 0) a = 0; rmin[0] = (uint32_t)a < (uint32_t)gx ? (uint32_t)a : (uint32_t)gx;
    a = (int)((p[1] - rf) * 0.0625f); if (a < 0) a = 0; rmin[1] = (uint32_t)a < (uint32_t)gy ? (uint32_t)a : (uint32_t)gy;
    a = (int)((((p[0] + rf) + 16.0f) + -1.0f) * 0.0625f); if (a < 0) a = 0; rmax[0] = (uint32_t)a < (uint32_t)gx ? (uint32_t)a : (uint32_t)gx;
    a = (int)((((p[1] + rf) + 16.0f) + -1.0f) * 0.0625f); if (a < 0) a = 0; rmax[1] = (uint32_t)a < (uint32_t)gy ? (uint32_t)a : (uint32_t)gy;
}

/* forward.cu:20-71 computeColorFromSH */
static void sh_to_rgb(int deg, int M, const float* pos, const float* campos, const float* sh, float* rgb,
                      uint8_t* clamped) {
    float dir[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    const float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    dir[0] /= len; dir[1] /= len; dir[2] /= len;
    const float x = dir[0], y = dir[1], z = dir[2];
    (void)M;
    for (int c = 0; c < 3; c++) {
#define SH(k) sh[3 * (k) + c]
        float res = SH_C0 * SH(0);
        if (deg > 0) {
            res = res - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                res = res + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) + SH_C2[2] * (2.0f * zz - xx - yy) * SH(6) +
                      SH_C2[3] * xz * SH(7) + SH_C2[4] * (xx - yy) * SH(8);
                if (deg > 2) {
                    res = res + SH_C3[0] * y * (3.0f * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
                          SH_C3[2] * y * (4.0f * zz - xx - yy) * SH(11) +
                          SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12) +
                          SH_C3[4] * x * (4.0f * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
                          SH_C3[6] * x * (xx - 3.0f * yy) * SH(15);
                }
            }
        }
#undef SH
        res += 0.5f;
        clamped[c] = res < 0;
        rgb[c] = res > 0.0f ? res : 0.0f;
    }
}

/* forward.cu:155-256 preprocessCUDA.  Outputs follow GeometryState (rasterizer_impl.h:29-43). */
void orc_preprocess(const orc_view* v, const float* means, const float* scales, const float* rots,
                    const float* opacities, const float* shs, const float* cov3D_precomp,
                    const float* colors_precomp, int32_t* radii, float* means2D, float* depths, float* cov3Ds,
                    float* rgb, float* conic_opacity, uint32_t* tiles_touched, uint8_t* clamped) {
    const float focal_y = v->H / (2.0f * v->tan_fovy), focal_x = v->W / (2.0f * v->tan_fovx);
    for (int idx = 0; idx < v->P; idx++) {
        radii[idx] = 0; tiles_touched[idx] = 0;
        const float* p = means + 3 * idx;
        const float pz = xform_row(v->view, 2, p);
        if (pz <= 0.2f) continue;                                   /* auxiliary.h:154 */
        const float hx = xform_row(v->proj, 0, p), hy = xform_row(v->proj, 1, p), hw = xform_row(v->proj, 3, p);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float projx = hx * p_w, projy = hy * p_w;
        const float* c6;
        if (cov3D_precomp) c6 = cov3D_precomp + 6 * idx;
        else { cov3d(scales + 3 * idx, v->scale_modifier, rots + 4 * idx, cov3Ds + 6 * idx); c6 = cov3Ds + 6 * idx; }
        const float t[3] = {xform_row(v->view, 0, p), xform_row(v->view, 1, p), pz};
        float T6[6], abc[3];
        ewa_T(v, t, focal_x, focal_y, T6);
        cov2d(T6, c6, abc);
        const float det = fmaf(abc[0], abc[2], -(abc[1] * abc[1]));
        if (det == 0.0f) continue;
        const float det_inv = 1.0f / det;
        const float conic[3] = {abc[2] * det_inv, det_inv * -abc[1], abc[0] * det_inv};
        const float mid = (abc[0] + abc[2]) * 0.5f;
        const float root = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
        const float lam = fmaxf(mid + root, mid - root);
        const float my_radius = ceilf(sqrtf(lam) * 3.0f);
        const float pix[2] = {ndc2pix(projx, v->W), ndc2pix(projy, v->H)};
        uint32_t rmin[2], rmax[2];
        get_rect(pix, (int)my_radius, v, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
        if (!colors_precomp) sh_to_rgb(v->D, v->M, p, v->campos, shs + (size_t)idx * v->M * 3, rgb + 3 * idx, clamped + 3 * idx);
        depths[idx] = pz;
        radii[idx] = (int)my_radius;
        means2D[2 * idx] = pix[0]; means2D[2 * idx + 1] = pix[1];
        conic_opacity[4 * idx] = conic[0]; conic_opacity[4 * idx + 1] = conic[1];
        conic_opacity[4 * idx + 2] = conic[2]; conic_opacity[4 * idx + 3] = opacities[idx];
        tiles_touched[idx] = (rmax[1] - rmin[1]) * (rmax[0] - rmin[0]);
    }
}

/* rasterizer_impl.cu:35-50 getHigherMsb */
uint32_t orc_higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) { step /= 2; if (n >> msb) msb += step; else msb -= step; }
    if (n >> msb) msb++;
    return msb;
}

typedef struct { uint64_t key; uint32_t val; } kv_t;
static void merge_sort(kv_t* a, kv_t* tmp, size_t n, uint64_t mask) {     /* stable, ascending on key & mask */
    if (n < 2) return;
    const size_t h = n / 2;
    merge_sort(a, tmp, h, mask); merge_sort(a + h, tmp, n - h, mask);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = ((a[j].key & mask) < (a[i].key & mask)) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(kv_t));
}

/* rasterizer_impl.cu:277-318: InclusiveSum, duplicateWithKeys (:70-111), SortPairs on
 * bits [0, 32+getHigherMsb(tiles)), memset + identifyTileRanges (:116-138).
 * keys_unsorted may be NULL.  Returns num_rendered (R); arrays must hold R entries
 * (R = sum of tiles_touched, the caller computes it first with orc_count). */
uint32_t orc_count(const orc_view* v, const uint32_t* tiles_touched, uint32_t* point_offsets) {
    uint32_t s = 0;
    for (int i = 0; i < v->P; i++) { s += tiles_touched[i]; if (point_offsets) point_offsets[i] = s; }
    return s;
}
uint32_t orc_bin(const orc_view* v, const int32_t* radii, const float* means2D, const float* depths,
                 const uint32_t* tiles_touched, uint64_t* keys_unsorted, uint64_t* keys_sorted,
                 uint32_t* point_list, uint32_t* ranges /* [tiles][2] */) {
    const int gx = grid_x(v), gy = grid_y(v);
    const uint32_t R = orc_count(v, tiles_touched, NULL);
    kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (R ? R : 1));
    kv_t* tmp = (kv_t*)malloc(sizeof(kv_t) * (R ? R : 1));
    uint32_t off = 0;
    for (int idx = 0; idx < v->P; idx++) {
        if (radii[idx] > 0) {
            uint32_t rmin[2], rmax[2];
            get_rect(means2D + 2 * idx, radii[idx], v, rmin, rmax);
            uint32_t dbits;
            memcpy(&dbits, depths + idx, 4);
            for (uint32_t y = rmin[1]; y < rmax[1]; y++)
                for (uint32_t x = rmin[0]; x < rmax[0]; x++) {
                    uint64_t key = (uint64_t)(y * (uint32_t)gx + x);
                    key <<= 32; key |= dbits;
                    kv[off].key = key; kv[off].val = (uint32_t)idx;
                    if (keys_unsorted) keys_unsorted[off] = key;
                    off++;
                }
        }
    }
    const int bit = 32 + (int)orc_higher_msb((uint32_t)(gx * gy));
    const uint64_t mask = bit >= 64 ? ~0ull : ((1ull << bit) - 1ull);
    merge_sort(kv, tmp, R, mask);
    for (uint32_t i = 0; i < R; i++) { keys_sorted[i] = kv[i].key; point_list[i] = kv[i].val; }
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    for (uint32_t i = 0; i < R; i++) {
        const uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            const uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
            if (cur != prev) { ranges[2 * prev + 1] = i; ranges[2 * cur] = i; }
        }
        if (i == R - 1) ranges[2 * cur + 1] = R;
    }
    free(kv); free(tmp);
    return R;
}

/* Blend exponent with the reference binary's fusion (forward.cu:335). */
static float blend_power(float dx, float dy, const float* con) {
    return fmaf(fmaf(dx, dx * con[0], dy * (dy * con[2])), -0.5f, -(dy * (dx * con[1])));
}

/* forward.cu:261-374 renderCUDA (forward): per pixel, front to back. */
void orc_render(const orc_view* v, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                const float* features, const float* conic_opacity, float* final_T, uint32_t* n_contrib,
                float* out_color) {
    const int gx = grid_x(v);
    const size_t HW = (size_t)v->W * v->H;
    for (int py = 0; py < v->H; py++)
        for (int px = 0; px < v->W; px++) {
            const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
            const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
            float T = 1.0f, C[3] = {0, 0, 0};
            uint32_t contributor = 0, last_contributor = 0;
            for (uint32_t i = r0; i < r1; i++) {
                contributor++;
                const uint32_t id = point_list[i];
                const float dx = means2D[2 * id] - (float)px, dy = means2D[2 * id + 1] - (float)py;
                const float* con = conic_opacity + 4 * id;
                const float power = blend_power(dx, dy, con);
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, con[3] * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) break;                       /* done = true */
                for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(T, alpha * features[3 * id + ch], C[ch]);
                T = test_T;
                last_contributor = contributor;
            }
            const size_t pix = (size_t)py * v->W + px;
            final_T[pix] = T;
            n_contrib[pix] = last_contributor;
            for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = fmaf(T, v->bg[ch], C[ch]);
        }
}

/* backward.cu:399-557 renderCUDA (backward): per pixel, back to front.
 * dL_dmean2D[P,3] (x,y used), dL_dconic[P,4] (.x,.y,.w used), dL_dopacity[P], dL_dcolors[P,3]
 * must be zero-initialised by the caller (rasterize_points.cu:151-159). */
void orc_render_backward(const orc_view* v, const uint32_t* ranges, const uint32_t* point_list,
                         const float* means2D, const float* conic_opacity, const float* colors,
                         const float* final_Ts, const uint32_t* n_contrib, const float* dL_dpixels,
                         float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors) {
    const int gx = grid_x(v);
    const size_t HW = (size_t)v->W * v->H;
    const float ddelx_dx = 0.5f * v->W, ddely_dy = 0.5f * v->H;
    for (int py = 0; py < v->H; py++)
        for (int px = 0; px < v->W; px++) {
            const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
            const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
            const size_t pix = (size_t)py * v->W + px;
            const float T_final = final_Ts[pix];
            float T = T_final;
            uint32_t contributor = r1 - r0;
            const uint32_t last_contributor = n_contrib[pix];
            float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0, dL_dpixel[3];
            for (int ch = 0; ch < 3; ch++) dL_dpixel[ch] = dL_dpixels[ch * HW + pix];
            for (uint32_t k = r1; k > r0; k--) {
                const uint32_t id = point_list[k - 1];
                contributor--;
                if (contributor >= last_contributor) continue;
                const float dx = means2D[2 * id] - (float)px, dy = means2D[2 * id + 1] - (float)py;
                const float* con = conic_opacity + 4 * id;
                const float power = blend_power(dx, dy, con);
                if (power > 0.0f) continue;
                const float G = expf(power);
                const float alpha = fminf(0.99f, con[3] * G);
                if (alpha < 1.0f / 255.0f) continue;
                T = T / (1.f - alpha);
                const float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.0f;
                for (int ch = 0; ch < 3; ch++) {
                    const float c = colors[3 * id + ch];
                    accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                    last_color[ch] = c;
                    dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
                    dL_dcolors[3 * id + ch] += dchannel_dcolor * dL_dpixel[ch];
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                float bg_dot_dpixel = 0;
                for (int ch = 0; ch < 3; ch++) bg_dot_dpixel += v->bg[ch] * dL_dpixel[ch];
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                const float dL_dG = con[3] * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * con[0] - gdy * con[1];
                const float dG_ddely = -gdy * con[2] - gdx * con[1];
                dL_dmean2D[3 * id] += dL_dG * dG_ddelx * ddelx_dx;
                dL_dmean2D[3 * id + 1] += dL_dG * dG_ddely * ddely_dy;
                dL_dconic[4 * id] += -0.5f * gdx * dx * dL_dG;
                dL_dconic[4 * id + 1] += -0.5f * gdx * dy * dL_dG;
                dL_dconic[4 * id + 3] += -0.5f * gdy * dy * dL_dG;
                dL_dopacity[id] += G * dL_dalpha;
            }
        }
}

/* auxiliary.h:107-117 dnormvdv(float3) */
static void dnormvdv3(const float* v, const float* dv, float* o) {
    const float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
    o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
    o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
    o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * inv;
}

/* backward.cu:144-274 computeCov2DCUDA, :346-396 preprocessCUDA (bwd), :20-139 SH bwd,
 * :278-341 cov3D bwd.  cov3Ds = the forward's cov3D (or the precomputed one).
 * dL_dmeans3D[P,3], dL_dcov3D[P,6], dL_dsh[P,M,3], dL_dscale[P,3], dL_drot[P,4] zeroed by caller. */
void orc_preprocess_backward(const orc_view* v, const float* means, const int32_t* radii, const float* shs,
                             const uint8_t* clamped, const float* scales, const float* rots,
                             const float* cov3Ds, const float* dL_dmean2D, const float* dL_dconics,
                             const float* dL_dcolor, float* dL_dmeans, float* dL_dcov, float* dL_dsh,
                             float* dL_dscale, float* dL_drot) {
    const float h_y = v->H / (2.0f * v->tan_fovy), h_x = v->W / (2.0f * v->tan_fovx);
    const float* Vm = v->view;
    const float* proj = v->proj;
    for (int idx = 0; idx < v->P; idx++) {
        if (!(radii[idx] > 0)) continue;
        const float* mean = means + 3 * idx;
        const float* cov3D = cov3Ds + 6 * idx;
        /* ---- computeCov2DCUDA ---- */
        const float dcx = dL_dconics[4 * idx], dcy = dL_dconics[4 * idx + 1], dcz = dL_dconics[4 * idx + 3];
        float t[3] = {xform_row(Vm, 0, mean), xform_row(Vm, 1, mean), xform_row(Vm, 2, mean)};
        const float limx = 1.3f * v->tan_fovx, limy = 1.3f * v->tan_fovy;
        const float txtz = t[0] / t[2], tytz = t[1] / t[2];
        float T6[6], abc[3];
        ewa_T(v, t, h_x, h_y, T6);
        t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
        t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        cov2d(T6, cov3D, abc);
        const float a = abc[0], b = abc[1], c = abc[2];
        const float T00 = T6[0], T01 = T6[1], T02 = T6[2], T10 = T6[3], T11 = T6[4], T12 = T6[5];
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dc = dL_dcov + 6 * idx;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dc[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
            dc[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
            dc[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
            dc[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
            dc[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
            dc[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
        } else {
            for (int i = 0; i < 6; i++) dc[i] = 0;
        }
        const float V00 = cov3D[0], V01 = cov3D[1], V02 = cov3D[2], V11 = cov3D[3], V12 = cov3D[4], V22 = cov3D[5];
        const float dL_dT00 = 2 * (T00 * V00 + T01 * V01 + T02 * V02) * dL_da + (T10 * V00 + T11 * V01 + T12 * V02) * dL_db;
        const float dL_dT01 = 2 * (T00 * V01 + T01 * V11 + T02 * V12) * dL_da + (T10 * V01 + T11 * V11 + T12 * V12) * dL_db;
        const float dL_dT02 = 2 * (T00 * V02 + T01 * V12 + T02 * V22) * dL_da + (T10 * V02 + T11 * V12 + T12 * V22) * dL_db;
        const float dL_dT10 = 2 * (T10 * V00 + T11 * V01 + T12 * V02) * dL_dc + (T00 * V00 + T01 * V01 + T02 * V02) * dL_db;
        const float dL_dT11 = 2 * (T10 * V01 + T11 * V11 + T12 * V12) * dL_dc + (T00 * V01 + T01 * V11 + T02 * V12) * dL_db;
        const float dL_dT12 = 2 * (T10 * V02 + T11 * V12 + T12 * V22) * dL_dc + (T00 * V02 + T01 * V12 + T02 * V22) * dL_db;
        const float dL_dJ00 = Vm[0] * dL_dT00 + Vm[4] * dL_dT01 + Vm[8] * dL_dT02;
        const float dL_dJ02 = Vm[2] * dL_dT00 + Vm[6] * dL_dT01 + Vm[10] * dL_dT02;
        const float dL_dJ11 = Vm[1] * dL_dT10 + Vm[5] * dL_dT11 + Vm[9] * dL_dT12;
        const float dL_dJ12 = Vm[2] * dL_dT10 + Vm[6] * dL_dT11 + Vm[10] * dL_dT12;
        const float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t[0]) * tz3 * dL_dJ02 + (2 * h_y * t[1]) * tz3 * dL_dJ12;
        float* dm = dL_dmeans + 3 * idx;
        dm[0] = Vm[0] * dL_dtx + Vm[1] * dL_dty + Vm[2] * dL_dtz;       /* assignment, backward.cu:273 */
        dm[1] = Vm[4] * dL_dtx + Vm[5] * dL_dty + Vm[6] * dL_dtz;
        dm[2] = Vm[8] * dL_dtx + Vm[9] * dL_dty + Vm[10] * dL_dtz;
        /* ---- preprocessCUDA (bwd): projection ---- */
        const float m_hom_w = proj[3] * mean[0] + proj[7] * mean[1] + proj[11] * mean[2] + proj[15];
        const float m_w = 1.0f / (m_hom_w + 0.0000001f);
        const float mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
        const float g2x = dL_dmean2D[3 * idx], g2y = dL_dmean2D[3 * idx + 1];
        dm[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dm[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dm[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        /* ---- SH backward ---- */
        if (shs) {
            const int M = v->M, deg = v->D;
            const float* sh = shs + (size_t)idx * M * 3;
            float* dsh = dL_dsh + (size_t)idx * M * 3;
            float dir_orig[3] = {mean[0] - v->campos[0], mean[1] - v->campos[1], mean[2] - v->campos[2]};
            const float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
            const float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
            float dRGB[3], dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[3 * idx + ch] * (clamped[3 * idx + ch] ? 0.f : 1.f);
            for (int ch = 0; ch < 3; ch++) {
#define SH(k) sh[3 * (k) + ch]
#define DSH(k) dsh[3 * (k) + ch]
                DSH(0) = SH_C0 * dRGB[ch];
                if (deg > 0) {
                    DSH(1) = -SH_C1 * y * dRGB[ch]; DSH(2) = SH_C1 * z * dRGB[ch]; DSH(3) = -SH_C1 * x * dRGB[ch];
                    dRGBdx[ch] = -SH_C1 * SH(3); dRGBdy[ch] = -SH_C1 * SH(1); dRGBdz[ch] = SH_C1 * SH(2);
                    if (deg > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        DSH(4) = SH_C2[0] * xy * dRGB[ch]; DSH(5) = SH_C2[1] * yz * dRGB[ch];
                        DSH(6) = SH_C2[2] * (2.f * zz - xx - yy) * dRGB[ch]; DSH(7) = SH_C2[3] * xz * dRGB[ch];
                        DSH(8) = SH_C2[4] * (xx - yy) * dRGB[ch];
                        dRGBdx[ch] += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.f * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * 2.f * x * SH(8);
                        dRGBdy[ch] += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.f * -y * SH(6) + SH_C2[4] * 2.f * -y * SH(8);
                        dRGBdz[ch] += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.f * 2.f * z * SH(6) + SH_C2[3] * x * SH(7);
                        if (deg > 2) {
                            DSH(9) = SH_C3[0] * y * (3.f * xx - yy) * dRGB[ch]; DSH(10) = SH_C3[1] * xy * z * dRGB[ch];
                            DSH(11) = SH_C3[2] * y * (4.f * zz - xx - yy) * dRGB[ch];
                            DSH(12) = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * dRGB[ch];
                            DSH(13) = SH_C3[4] * x * (4.f * zz - xx - yy) * dRGB[ch];
                            DSH(14) = SH_C3[5] * z * (xx - yy) * dRGB[ch]; DSH(15) = SH_C3[6] * x * (xx - 3.f * yy) * dRGB[ch];
                            dRGBdx[ch] += (SH_C3[0] * SH(9) * 3.f * 2.f * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * -2.f * xy +
                                           SH_C3[3] * SH(12) * -3.f * 2.f * xz + SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) +
                                           SH_C3[5] * SH(14) * 2.f * xz + SH_C3[6] * SH(15) * 3.f * (xx - yy));
                            dRGBdy[ch] += (SH_C3[0] * SH(9) * 3.f * (xx - yy) + SH_C3[1] * SH(10) * xz +
                                           SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12) * -3.f * 2.f * yz +
                                           SH_C3[4] * SH(13) * -2.f * xy + SH_C3[5] * SH(14) * -2.f * yz + SH_C3[6] * SH(15) * -3.f * 2.f * xy);
                            dRGBdz[ch] += (SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.f * 2.f * yz +
                                           SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13) * 4.f * 2.f * xz +
                                           SH_C3[5] * SH(14) * (xx - yy));
                        }
                    }
                }
#undef SH
#undef DSH
            }
            const float dL_ddir[3] = {dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2],
                                      dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2],
                                      dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2]};
            float dmean_sh[3];
            dnormvdv3(dir_orig, dL_ddir, dmean_sh);
            dm[0] += dmean_sh[0]; dm[1] += dmean_sh[1]; dm[2] += dmean_sh[2];
        }
        /* ---- cov3D backward (backward.cu:278-341), column-major glm restated with [c][r] arrays ---- */
        if (scales) {
            const float* q = rots + 4 * idx;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                   {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                   {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float s[3] = {v->scale_modifier * scales[3 * idx], v->scale_modifier * scales[3 * idx + 1],
                                v->scale_modifier * scales[3 * idx + 2]};
            float M[3][3], dM[3][3], dMt[3][3];
            for (int c2 = 0; c2 < 3; c2++) for (int rr = 0; rr < 3; rr++) M[c2][rr] = s[rr] * R[c2][rr];
            const float dS[3][3] = {{dc[0], 0.5f * dc[1], 0.5f * dc[2]}, {0.5f * dc[1], dc[3], 0.5f * dc[4]},
                                    {0.5f * dc[2], 0.5f * dc[4], dc[5]}};
            for (int c2 = 0; c2 < 3; c2++) for (int rr = 0; rr < 3; rr++)
                dM[c2][rr] = 2.0f * (M[0][rr] * dS[c2][0] + M[1][rr] * dS[c2][1] + M[2][rr] * dS[c2][2]);
            for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) dMt[k][j] = dM[j][k];
            for (int k = 0; k < 3; k++)
                dL_dscale[3 * idx + k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
            for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) dMt[k][j] *= s[k];
            float* dq = dL_drot + 4 * idx;
            dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
            dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
            dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* rasterizer_impl.cu:54-66 checkFrustum */
void orc_mark_visible(const orc_view* v, const float* means, uint8_t* present) {
    for (int i = 0; i < v->P; i++) present[i] = xform_row(v->view, 2, means + 3 * i) > 0.2f;
}

/* simple-knn/simple_knn.cu:147-183 boxMeanDist semantics by brute force: mean of the
 * squared distances to the 3 nearest OTHER points (by index), O(P^2). */
void orc_knn_dist2(int P, const float* pts, float* out) {
    for (int i = 0; i < P; i++) {
        float best[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
        for (int j = 0; j < P; j++) {
            if (j == i) continue;
            const float dx = pts[3 * j] - pts[3 * i], dy = pts[3 * j + 1] - pts[3 * i + 1], dz = pts[3 * j + 2] - pts[3 * i + 2];
            float d = dx * dx + dy * dy + dz * dz;
            for (int k = 0; k < 3; k++) if (best[k] > d) { const float t = best[k]; best[k] = d; d = t; }
        }
        out[i] = (best[0] + best[1] + best[2]) / 3.0f;
    }
}
