// pcm_slic.cpp -- SLIC superpixels of the crop, the third over-segmentation the reference offers
// (maskers/pixel_classification.py:74-75: slic(crop, n_segments=250, compactness=10, sigma=1, start_label=0)).
// SURVEY.md §8 row f-1.
//
// HOST code, like pcm_felzenszwalb.cpp: it feeds the hot path its labels and is not part of it (neither the reference's
// default config nor its sweep select SLIC).  Algorithm = scikit-image 0.17.2 (environment.yaml:12; not in the reference
// tree and not installed: parity is pinned only against oracle/slic_oracle.py, which restates slic_superpixels.py,
// _slic.pyx and _regular_grid.py; its Gaussian step is the scipy-exact one of the felzenszwalb oracle):
//   img_as_float -> Gaussian (depth axis of length 1, rows, columns) -> rgb2lab -> seeds on a regular grid ->
//   10 k-means iterations in (y, x, L/c, a/c, b/c) with +-2 step windows -> connectivity enforcement.
// Every sum runs in the order of the Cython loops with separate multiplies and adds (-ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace pcm {

static inline int reflect_index(int i, int n) {      // scipy 'reflect': d c b a | a b c d | d c b a
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
}

// symmetric correlate along one axis of an h x w x 3 image: centre tap first, then the pairs from the outside in
static void slic_smooth_axis(const std::vector<double>& in, std::vector<double>& out, int h, int w, int axis,
                             const double* k, int radius) {
    const int n = axis == 0 ? h : w;
    std::vector<int> idx((size_t)n + 2 * radius);
    for (int p = -radius; p < n + radius; ++p) idx[p + radius] = reflect_index(p, n);
    const size_t step = axis == 0 ? (size_t)w * 3 : 3;
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            const int pos = axis == 0 ? r : c;
            const size_t line0 = axis == 0 ? (size_t)c * 3 : (size_t)r * w * 3;
            const double* ctr = &in[line0 + (size_t)pos * step];
            double* o = &out[((size_t)r * w + c) * 3];
            double a0 = ctr[0] * k[radius], a1 = ctr[1] * k[radius], a2 = ctr[2] * k[radius];
            for (int j = radius; j >= 1; --j) {
                const double* lo = &in[line0 + (size_t)idx[pos - j + radius] * step];
                const double* hi = &in[line0 + (size_t)idx[pos + j + radius] * step];
                const double wj = k[radius - j];
                a0 = a0 + (lo[0] + hi[0]) * wj;
                a1 = a1 + (lo[1] + hi[1]) * wj;
                a2 = a2 + (lo[2] + hi[2]) * wj;
            }
            o[0] = a0; o[1] = a1; o[2] = a2;
        }
}

// regular_grid((1, h, w), n_points): start and step of the seed grid along rows and columns; false: every pixel is a seed
static bool slic_grid(int h, int w, int n_points, int& start_y, int& step_y, int& start_x, int& step_x) {
    const double space = (double)h * (double)w;
    if (space <= (double)n_points) return false;
    // sorted dims (1, min, max): the unit depth is smaller than the cubic step, so it gets step 1 and the two image
    // axes share the square-root step (further clipping, as in the original, when the short axis is shorter still)
    double steps[3] = {std::pow(space / n_points, 1.0 / 3), 0, 0};
    steps[1] = steps[2] = steps[0];
    const double dims[3] = {1.0, (double)std::min(h, w), (double)std::max(h, w)};
    bool any_short = false;
    for (int d = 0; d < 3; ++d) any_short = any_short || dims[d] < steps[d];
    if (any_short) {
        for (int d = 0; d < 3; ++d) {
            steps[d] = dims[d];
            double sp = 1.0;
            for (int e = d + 1; e < 3; ++e) sp *= dims[e];
            for (int e = d + 1; e < 3; ++e) steps[e] = std::pow(sp / n_points, 1.0 / (3 - d - 1));
            bool ok = true;
            for (int e = 0; e < 3; ++e) ok = ok && dims[e] >= steps[e];
            if (ok) break;
        }
    }
    const int s_min = (int)std::floor(steps[1] / 2), s_max = (int)std::floor(steps[2] / 2);
    const int t_min = (int)std::nearbyint(steps[1]), t_max = (int)std::nearbyint(steps[2]);       // np.round: half to even
    if (h <= w) { start_y = s_min; step_y = t_min; start_x = s_max; step_x = t_max; }
    else        { start_y = s_max; step_y = t_max; start_x = s_min; step_x = t_min; }
    step_y = std::max(step_y, 1); step_x = std::max(step_x, 1);
    return true;
}

// labels_out[h*w]; returns the number of labels (max label + 1 - start_label... the count of distinct new labels), -1 on bad arguments
int slic(const uint8_t* frame, int64_t stride, int cx, int cy, int w, int h, int n_segments, double compactness, double sigma,
         const double* kernel, int radius, int max_iter, int start_label, int32_t* labels_out) {
    if (w <= 0 || h <= 0 || n_segments < 1 || !(compactness > 0) || max_iter < 0) return -1;
    const size_t n = (size_t)w * h;
    std::vector<double> img(n * 3), tmp;
    for (int r = 0; r < h; ++r) {
        const uint8_t* row = frame + (int64_t)(cy + r) * stride + (int64_t)cx * 3;
        for (int i = 0; i < 3 * w; ++i) img[(size_t)r * w * 3 + i] = row[i] / 255.0;
    }
    if (sigma > 0) {
        std::vector<double> kw;
        if (!kernel) {                                   // scipy _gaussian_kernel1d, truncate = 4
            radius = (int)(4.0 * sigma + 0.5);
            kw.resize(2 * radius + 1);
            double sum = 0;
            for (int x = -radius; x <= radius; ++x) { kw[x + radius] = std::exp(-0.5 / (sigma * sigma) * (double)(x * x)); sum += kw[x + radius]; }
            for (double& v : kw) v /= sum;
            kernel = kw.data();
        }
        // depth axis (length 1): every reflected tap is the pixel itself
        for (double& v : img) {
            double a = v * kernel[radius];
            for (int j = radius; j >= 1; --j) a = a + (v + v) * kernel[radius - j];
            v = a;
        }
        tmp.resize(n * 3);
        slic_smooth_axis(img, tmp, h, w, 0, kernel, radius);
        slic_smooth_axis(tmp, img, h, w, 1, kernel, radius);
    }
    // rgb2lab of skimage.color on float pixels (channel 0 plays 'R'), then / compactness
    const double M[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};
    const double white[3] = {0.95047, 1.0, 1.08883};
    const double ratio = 1.0 / compactness;
    for (size_t i = 0; i < n; ++i) {
        double lin[3], f[3];
        for (int c = 0; c < 3; ++c) {
            const double v = img[3 * i + c];
            lin[c] = v > 0.04045 ? std::pow((v + 0.055) / 1.055, 2.4) : v / 12.92;
        }
        for (int k = 0; k < 3; ++k) {
            const double t = ((lin[0] * M[k][0] + lin[1] * M[k][1]) + lin[2] * M[k][2]) / white[k];
            f[k] = t > 0.008856 ? std::cbrt(t) : 7.787 * t + 16.0 / 116.0;
        }
        img[3 * i + 0] = (116.0 * f[1] - 16.0) * ratio;
        img[3 * i + 1] = (500.0 * (f[0] - f[1])) * ratio;
        img[3 * i + 2] = (200.0 * (f[1] - f[2])) * ratio;
    }
    // seeds
    int sy = 0, ty = 1, sx = 0, tx = 1;
    std::vector<int> ys, xs;
    if (slic_grid(h, w, n_segments, sy, ty, sx, tx)) {
        for (int y = sy; y < h; y += ty) ys.push_back(y);
        for (int x = sx; x < w; x += tx) xs.push_back(x);
    } else {
        for (int y = 0; y < h; ++y) ys.push_back(y);
        for (int x = 0; x < w; ++x) xs.push_back(x);
        ty = tx = 1;
    }
    const int K = (int)(ys.size() * xs.size());
    if (K < 1) return -1;
    std::vector<double> centers((size_t)K * 5, 0.0);                 // y, x, L, a, b
    for (size_t a = 0; a < ys.size(); ++a)
        for (size_t b = 0; b < xs.size(); ++b) {
            centers[(a * xs.size() + b) * 5 + 0] = ys[a];
            centers[(a * xs.size() + b) * 5 + 1] = xs[b];
        }
    const double step = (double)std::max(1, std::max(ty, tx));
    const double spatial_weight = 1.0 / (step * step);
    std::vector<int> nearest(n, 0);
    std::vector<double> dist(n);
    std::vector<long long> count((size_t)K);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int it = 0; it < max_iter; ++it) {
        std::fill(dist.begin(), dist.end(), std::numeric_limits<double>::infinity());
        for (int k = 0; k < K; ++k) {
            const double* ck = &centers[(size_t)k * 5];
            const double cyk = ck[0], cxk = ck[1];
            if (!(cyk == cyk)) continue;                              // a cluster that lost all its pixels
            const int y_min = (int)std::max(cyk - 2 * ty, 0.0), y_max = (int)std::min(cyk + 2 * ty + 1, (double)h);
            const int x_min = (int)std::max(cxk - 2 * tx, 0.0), x_max = (int)std::min(cxk + 2 * tx + 1, (double)w);
            for (int y = y_min; y < y_max; ++y) {
                const double ddy = (cyk - y) * (cyk - y);
                for (int x = x_min; x < x_max; ++x) {
                    const double ddx = (cxk - x) * (cxk - x);
                    double d = (0.0 + ddy + ddx) * spatial_weight;
                    const double* px = &img[((size_t)y * w + x) * 3];
                    double dc = 0.0;
                    for (int c = 0; c < 3; ++c) { const double t = px[c] - ck[2 + c]; dc += t * t; }
                    d += dc;
                    const size_t o = (size_t)y * w + x;
                    if (dist[o] > d) { nearest[o] = k; dist[o] = d; }
                }
            }
        }
        std::fill(count.begin(), count.end(), 0);
        std::fill(centers.begin(), centers.end(), 0.0);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const size_t o = (size_t)y * w + x;
                double* ck = &centers[(size_t)nearest[o] * 5];
                count[nearest[o]]++;
                ck[0] += y; ck[1] += x;
                ck[2] += img[3 * o]; ck[3] += img[3 * o + 1]; ck[4] += img[3 * o + 2];
            }
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < 5; ++c) centers[(size_t)k * 5 + c] = count[k] > 0 ? centers[(size_t)k * 5 + c] / (double)count[k] : nan;
    }
    // connectivity (min_size_factor 0.5, max_size_factor 3 of the requested segment size)
    const double segment_size = (double)h * (double)w / n_segments;
    const long long min_size = (long long)(0.5 * segment_size), max_size = (long long)(3 * segment_size);
    std::vector<int> out(n, -1);
    std::vector<int> coords((size_t)std::max<long long>(max_size, 1) * 2);
    static const int dy4[4] = {0, 0, 1, -1}, dx4[4] = {1, -1, 0, 0};
    int new_label = start_label;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t o = (size_t)y * w + x;
            if (out[o] >= 0) continue;
            int adjacent = 0;
            const int label = nearest[o];
            out[o] = new_label;
            long long size = 1, visited = 0;
            coords[0] = y; coords[1] = x;
            while (visited < size && size < max_size) {
                for (int i = 0; i < 4; ++i) {
                    const int yy = coords[2 * visited] + dy4[i], xx = coords[2 * visited + 1] + dx4[i];
                    if (xx >= 0 && xx < w && yy >= 0 && yy < h) {
                        const size_t q = (size_t)yy * w + xx;
                        if (nearest[q] == label && out[q] == -1) {
                            out[q] = new_label;
                            coords[2 * size] = yy; coords[2 * size + 1] = xx;
                            ++size;
                            if (size >= max_size) break;
                        } else if (out[q] >= 0 && out[q] != new_label) {
                            adjacent = out[q];
                        }
                    }
                }
                ++visited;
            }
            if (size < min_size) {
                for (long long i = 0; i < size; ++i) out[(size_t)coords[2 * i] * w + coords[2 * i + 1]] = adjacent;
            } else {
                ++new_label;
            }
        }
    int mx = -1;
    for (size_t i = 0; i < n; ++i) { labels_out[i] = out[i]; mx = std::max(mx, out[i]); }
    return mx + 1;
}

}  // namespace pcm
