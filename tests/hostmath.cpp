// Host build of patchmatchnet_b200/csrc/pm_math.cuh for CPU-side formula tests.
// TEST INFRASTRUCTURE: compiled by tests/conftest.py into tests/_hostmath.so with g++; it lets the
// build box (no GPU) check the per-element arithmetic the kernels inline against the oracle.
// It is never loaded by the product package.
#include "../patchmatchnet_b200/csrc/pm_math.cuh"
#include "../patchmatchnet_b200/csrc/pm_geo_math.cuh"

extern "C" {

// footprints of every (hypothesis, pixel) of one view: weights [D*HW*4], keys [D*HW]
void hm_warp_cells(const float *rt, const float *depth, int H, int W, int Hs, int Ws, int D, float *w, int *key) {
    const float sx = W > 1 ? (float)(Ws - 1) / (float)(W - 1) : 1.0f;
    const float sy = H > 1 ? (float)(Hs - 1) / (float)(H - 1) : 1.0f;
    for (int d = 0; d < D; ++d)
        for (int n = 0; n < H * W; ++n) {
            const pm::Ray ray = pm::pixel_ray(rt, (float)(n % W), (float)(n / W));
            float u, v;
            pm::project(ray, rt, depth[(size_t)d * H * W + n], W, H, sx, sy, &u, &v);
            const pm::Cell c = pm::zero_pad_cell(u, v, Hs, Ws);
            const size_t i = (size_t)d * H * W + n;
            w[4 * i] = c.w00; w[4 * i + 1] = c.w01; w[4 * i + 2] = c.w10; w[4 * i + 3] = c.w11;
            key[i] = c.key;
        }
}

// footprints of the K learned neighbours of every pixel: offsets [2K,HW] -> weights [K*HW*4], keys [K*HW]
int hm_neighbour_cells(const float *offsets, int evaluation, int K, int dilation, int H, int W, float *w, int *key) {
    for (int k = 0; k < K; ++k) {
        int dy, dx;
        if (!pm::neighbour_offset(evaluation != 0, K, dilation, k, &dy, &dx)) return -1;
        for (int n = 0; n < H * W; ++n) {
            const float ox = (float)dx + offsets[(size_t)(2 * k) * H * W + n];
            const float oy = (float)dy + offsets[(size_t)(2 * k + 1) * H * W + n];
            const pm::Cell c = pm::border_cell((float)(n % W) + ox, (float)(n / W) + oy, H, W);
            const size_t i = (size_t)k * H * W + n;
            w[4 * i] = c.w00; w[4 * i + 1] = c.w01; w[4 * i + 2] = c.w10; w[4 * i + 3] = c.w11;
            key[i] = c.key;
        }
    }
    return 0;
}

int hm_neighbour_table(int evaluation, int K, int dilation, int *dy, int *dx) {
    for (int k = 0; k < K; ++k)
        if (!pm::neighbour_offset(evaluation != 0, K, dilation, k, dy + k, dx + k)) return -1;
    return 0;
}

void hm_key_unpack(int key, int *r0, int *dx, int *dy) {
    *r0 = pm::cell_r0(key); *dx = pm::cell_dx(key); *dy = pm::cell_dy(key);
}

int hm_key_none(void) { return pm::kKeyNone; }

float hm_random_hypothesis(float u, int bin, float inv_min, float inv_max) {
    return pm::random_hypothesis(u, bin, inv_min, inv_max);
}

float hm_perturbed_hypothesis(float depth, int k, int ns, float inv_min, float inv_max, float scale) {
    return pm::perturbed_hypothesis(depth, pm::floor_div2_neg(ns) + k, inv_min, inv_max, scale);
}

float hm_depth_similarity(float xc, float xn, float scale) { return pm::depth_similarity(xc, xn, scale); }

float hm_normalised_inverse_depth(float d, float inv_min, float inv_max) {
    return pm::normalised_inverse_depth(d, inv_min, inv_max);
}

// geometric-consistency filter, pixel by pixel with the kernel's own per-pixel function (pm_geo_math.cuh)
void hm_geometric_filter(const float *ref_depth, const float *confidence, const float *src_depths, const double *cams, int V,
                         int H, int W, int Hs, int Ws, double pixel_thres, float depth_thres, float photo_thres, int mask_thres,
                         int *mask_sum, unsigned char *photo_mask, unsigned char *final_mask, double *depth_avg) {
    for (int n = 0; n < H * W; ++n) {
        const pmgeo::PixelResult r = pmgeo::filter_pixel(cams, src_depths, V, Hs, Ws, n % W, n / W, ref_depth[n], confidence[n],
                                                         pixel_thres, depth_thres, photo_thres, mask_thres);
        mask_sum[n] = r.count;
        photo_mask[n] = r.photo ? 1 : 0;
        final_mask[n] = r.final ? 1 : 0;
        depth_avg[n] = r.depth_avg;
    }
}

// fusion half: the kernel's per-pixel function applied to the surviving pixels in row-major order; returns the count
int hm_fuse_points(const unsigned char *mask, const double *depth_avg, const float *rgb, const double *cam25, int H, int W,
                   unsigned char *body) {
    int n = 0;
    for (int i = 0; i < H * W; ++i)
        if (mask[i]) pmgeo::fuse_point(cam25, i % W, i / W, depth_avg[i], rgb + (size_t)i * 3, body + (size_t)(n++) * pmgeo::kPlyVertexBytes);
    return n;
}

void hm_remap_linear(const float *src, int rows, int cols, const float *mx, const float *my, int n, float *out) {
    for (int i = 0; i < n; ++i) out[i] = pmgeo::remap_linear(src, rows, cols, mx[i], my[i]);
}

}  // extern "C"
