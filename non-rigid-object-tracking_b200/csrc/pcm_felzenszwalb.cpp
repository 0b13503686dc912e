// pcm_felzenszwalb.cpp -- Felzenszwalb-Huttenlocher graph segmentation of the crop, the second
// over-segmentation the reference's sweep uses (maskers/pixel_classification.py:72-73,
// benchmark.py:47: felzenszwalb(crop, scale=100, sigma=0.5, min_size=50)).  SURVEY.md §8 row f-1.
//
// HOST code on purpose: the algorithm is a Kruskal-style pass over the edges in cost order whose
// merge test depends on everything merged before -- it does not parallelise without changing
// the result -- and it sits outside the per-frame hot path (it feeds the hot path its labels).
// Algorithm = scikit-image 0.17.2's _felzenszwalb_cython (environment.yaml:12; not in the
// reference tree: parity pinned only against oracle/felzenszwalb_oracle.py, which in turn pins the
// Gaussian step against scipy.ndimage).  Edges are sorted STABLY (cost, then edge index), where
// scikit-image leaves the order of equal costs to numpy's quicksort.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace pcm {

static inline int reflect(int i, int n) {      // scipy 'reflect': d c b a | a b c d | d c b a
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
}

// out = gaussian along one axis of an h x w x 3 float64 image (centre tap first, then the pairs
// from the outside in: the summation order of scipy's symmetric correlate1d)
static void smooth_axis(const std::vector<double>& in, std::vector<double>& out, int h, int w, int axis,
                        const double* k, int radius) {
    const int n = axis == 0 ? h : w;
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            const int pos = axis == 0 ? r : c;
            for (int ch = 0; ch < 3; ++ch) {
                auto at = [&](int p) { return axis == 0 ? in[((size_t)reflect(p, n) * w + c) * 3 + ch]
                                                        : in[((size_t)r * w + reflect(p, n)) * 3 + ch]; };
                double acc = at(pos) * k[radius];
                for (int j = radius; j >= 1; --j) acc = acc + (at(pos - j) + at(pos + j)) * k[radius - j];
                out[((size_t)r * w + c) * 3 + ch] = acc;
            }
        }
}

static int find_root(std::vector<int>& parent, int i) {
    while (parent[i] != i) i = parent[i];
    return i;
}
static void join(std::vector<int>& parent, int n, int m) {
    const int rn = find_root(parent, n), rm = find_root(parent, m);
    const int root = rn < rm ? rn : rm;
    for (int start : {n, m}) {
        int i = start;
        while (parent[i] != i) { const int nx = parent[i]; parent[i] = root; i = nx; }
        parent[i] = root;
    }
}

// labels_out[h*w]; returns the number of segments, or -1 on bad arguments
int felzenszwalb(const uint8_t* frame, int64_t stride, int cx, int cy, int w, int h, double scale, double sigma,
                 int min_size, const double* kernel, int radius, int32_t* labels_out) {
    if (w <= 0 || h <= 0) return -1;
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
        tmp.resize(n * 3);
        smooth_axis(img, tmp, h, w, 0, kernel, radius);
        smooth_axis(tmp, img, h, w, 1, kernel, radius);
    }
    const double sc = scale / 255.0;
    // edges in scikit-image's order: right, down, down-right, up-right
    struct Edge { double cost; int a, b; };
    std::vector<Edge> edges;
    edges.reserve(4 * n);
    auto cost = [&](int r0, int c0, int r1, int c1) {
        const double* p = &img[((size_t)r0 * w + c0) * 3];
        const double* q = &img[((size_t)r1 * w + c1) * 3];
        const double d0 = p[0] - q[0], d1 = p[1] - q[1], d2 = p[2] - q[2];
        return std::sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    };
    for (int r = 0; r < h; ++r) for (int c = 1; c < w; ++c) edges.push_back({cost(r, c, r, c - 1), r * w + c, r * w + c - 1});
    for (int r = 1; r < h; ++r) for (int c = 0; c < w; ++c) edges.push_back({cost(r, c, r - 1, c), r * w + c, (r - 1) * w + c});
    for (int r = 1; r < h; ++r) for (int c = 1; c < w; ++c) edges.push_back({cost(r, c, r - 1, c - 1), r * w + c, (r - 1) * w + c - 1});
    for (int r = 1; r < h; ++r) for (int c = 1; c < w; ++c) edges.push_back({cost(r, c - 1, r - 1, c), (r - 1) * w + c, r * w + c - 1});
    std::stable_sort(edges.begin(), edges.end(), [](const Edge& x, const Edge& y) { return x.cost < y.cost; });

    std::vector<int> parent(n), size(n, 1);
    std::vector<double> cint(n, 0.0);
    std::iota(parent.begin(), parent.end(), 0);
    for (const Edge& e : edges) {
        const int s0 = find_root(parent, e.a), s1 = find_root(parent, e.b);
        if (s0 == s1) continue;
        if (e.cost < std::min(cint[s0] + sc / size[s0], cint[s1] + sc / size[s1])) {
            join(parent, s0, s1);
            const int r = find_root(parent, s0);
            size[r] = size[s0] + size[s1];
            cint[r] = e.cost;
        }
    }
    for (const Edge& e : edges) {
        const int s0 = find_root(parent, e.a), s1 = find_root(parent, e.b);
        if (s0 == s1) continue;
        if (size[s0] < min_size || size[s1] < min_size) {
            join(parent, s0, s1);
            const int r = find_root(parent, s0);
            size[r] = size[s0] + size[s1];
        }
    }
    // np.unique(root, return_inverse=True)[1]: rank of the root index
    std::vector<int> rank(n, 0);
    int count = 0;
    for (size_t i = 0; i < n; ++i) rank[i] = (find_root(parent, (int)i) == (int)i) ? count++ : -1;
    for (size_t i = 0; i < n; ++i) labels_out[i] = rank[find_root(parent, (int)i)];
    return count;
}

}  // namespace pcm
