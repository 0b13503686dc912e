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
#include <cstring>
#include <numeric>
#include <vector>

namespace pcm {

static inline int reflect(int i, int n) {      // scipy 'reflect': d c b a | a b c d | d c b a
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
}

// out = gaussian along one axis of an h x w x 3 float64 image (centre tap first, then the pairs
// from the outside in: the summation order of scipy's symmetric correlate1d).  `idx` maps a
// position of the padded line [-radius, n + radius) to the reflected source position.
static void smooth_axis(const std::vector<double>& in, std::vector<double>& out, int h, int w, int axis,
                        const double* k, int radius) {
    const int n = axis == 0 ? h : w;
    std::vector<int> idx((size_t)n + 2 * radius);
    for (int p = -radius; p < n + radius; ++p) idx[p + radius] = reflect(p, n);
    const size_t step = axis == 0 ? (size_t)w * 3 : 3;       // distance between neighbours along the axis
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            const int pos = axis == 0 ? r : c;
            const size_t line0 = axis == 0 ? (size_t)c * 3 : (size_t)r * w * 3;   // element 0 of this line
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

// path halving: shortens later searches; the partition (all that matters) is unchanged
static inline int find_root(std::vector<int>& parent, int i) {
    while (parent[i] != i) {
        parent[i] = parent[parent[i]];
        i = parent[i];
    }
    return i;
}
// the smaller root index wins, as in scikit-image's join_trees
static inline int join(std::vector<int>& parent, int rn, int rm) {
    const int root = rn < rm ? rn : rm;
    parent[rn] = root;
    parent[rm] = root;
    return root;
}

struct Edge { double cost; int a, b; };

// stable LSD radix sort by cost: non-negative doubles order like their bit patterns
static void sort_edges(std::vector<Edge>& e) {
    std::vector<Edge> tmp(e.size());
    auto key = [](const Edge& x) { uint64_t u; memcpy(&u, &x.cost, 8); return u; };
    for (int shift = 0; shift < 64; shift += 11) {
        size_t count[2049] = {0};
        for (const Edge& x : e) count[((key(x) >> shift) & 2047) + 1]++;
        bool single = false;
        for (int i = 1; i <= 2048; ++i) if (count[i] == e.size()) single = true;
        if (single) continue;                                   // every key has the same digit here
        for (int i = 1; i <= 2048; ++i) count[i] += count[i - 1];
        for (const Edge& x : e) tmp[count[(key(x) >> shift) & 2047]++] = x;
        e.swap(tmp);
    }
}

// Stable cost sort, greedy merge (cost < min over both components of Int + scale / |C|), min-size pass,
// labels = rank of the root index.  Returns the number of segments.
static int merge_sorted_edges(std::vector<Edge>& edges, size_t n, double sc, int min_size, int32_t* labels_out) {
    sort_edges(edges);

    std::vector<int> parent(n), size(n, 1);
    std::vector<double> cint(n, 0.0);
    std::iota(parent.begin(), parent.end(), 0);
    for (const Edge& e : edges) {
        const int s0 = find_root(parent, e.a), s1 = find_root(parent, e.b);
        if (s0 == s1) continue;
        if (e.cost < std::min(cint[s0] + sc / size[s0], cint[s1] + sc / size[s1])) {
            const int total = size[s0] + size[s1];
            const int r = join(parent, s0, s1);
            size[r] = total;
            cint[r] = e.cost;
        }
    }
    for (const Edge& e : edges) {
        const int s0 = find_root(parent, e.a), s1 = find_root(parent, e.b);
        if (s0 == s1) continue;
        if (size[s0] < min_size || size[s1] < min_size) {
            const int total = size[s0] + size[s1];
            size[join(parent, s0, s1)] = total;
        }
    }
    // np.unique(root, return_inverse=True)[1]: rank of the root index
    std::vector<int> rank(n, 0);
    int count = 0;
    for (size_t i = 0; i < n; ++i) rank[i] = (find_root(parent, (int)i) == (int)i) ? count++ : -1;
    for (size_t i = 0; i < n; ++i) labels_out[i] = rank[find_root(parent, (int)i)];
    return count;
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
    return merge_sorted_edges(edges, n, sc, min_size, labels_out);
}

// Parity tap (pcm_felzenszwalb_graph): the passes above on a caller-provided edge list.
int felzenszwalb_graph(int n_vertices, int n_edges, const int32_t* a, const int32_t* b, const double* cost, double scale,
                       int min_size, int32_t* labels_out) {
    if (n_vertices <= 0 || n_edges < 0) return -1;
    std::vector<Edge> edges((size_t)n_edges);
    for (int i = 0; i < n_edges; ++i) {
        if (a[i] < 0 || a[i] >= n_vertices || b[i] < 0 || b[i] >= n_vertices || !(cost[i] >= 0.0)) return -1;
        edges[i] = {cost[i], a[i], b[i]};
    }
    return merge_sorted_edges(edges, (size_t)n_vertices, scale, min_size, labels_out);
}

}  // namespace pcm
