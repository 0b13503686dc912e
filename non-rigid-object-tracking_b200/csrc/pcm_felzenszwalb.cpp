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

// Working memory of one call, kept per thread: a call needs ~170 bytes per pixel in half a dozen arrays, and fresh
// allocations of that size come from the kernel page by page (a fifth of the time of a call on a 300 x 240 crop).
struct Scratch {
    std::vector<double> img, tmp, limit;
    std::vector<Edge> edges, edges_tmp;
    std::vector<int> parent, size, rank;
};
static Scratch& scratch() {
    static thread_local Scratch s;
    return s;
}
// after a call on a large image (more than a megapixel: > 170 MB of working memory) nothing is kept
static void trim_scratch(size_t n_vertices) {
    if (n_vertices > ((size_t)1 << 20)) scratch() = Scratch();
}

// Stable sort by cost (non-negative doubles order like their bit patterns; equal costs keep their edge order).
// The 64-bit keys are not sorted whole: an LSD radix sort orders the edges by the UPPER key half (four 8-bit digits, one
// histogram pass for all of them, digits that are the same everywhere skipped), and the runs that still tie there --
// short ones: the upper half holds the exponent and 20 mantissa bits -- are ordered by the lower half, stably.
static void sort_edges(std::vector<Edge>& e, std::vector<Edge>& tmp) {
    const size_t n = e.size();
    if (n < 2) return;
    auto key = [](const Edge& x) { uint64_t u; memcpy(&u, &x.cost, 8); return u; };
    tmp.resize(n);
    size_t count[4][257] = {};
    for (size_t i = 0; i < n; ++i) {
        const uint64_t k = key(e[i]);
        count[0][((k >> 32) & 255) + 1]++;
        count[1][((k >> 40) & 255) + 1]++;
        count[2][((k >> 48) & 255) + 1]++;
        count[3][((k >> 56) & 255) + 1]++;
    }
    for (int p = 0; p < 4; ++p) {
        size_t* c = count[p];
        bool single = false;
        for (int i = 1; i <= 256; ++i) if (c[i] == n) single = true;
        if (single) continue;                                   // every key has the same digit here
        for (int i = 1; i <= 256; ++i) c[i] += c[i - 1];
        const int shift = 32 + 8 * p;
        for (size_t i = 0; i < n; ++i) tmp[c[(key(e[i]) >> shift) & 255]++] = e[i];
        e.swap(tmp);
    }
    for (size_t i = 0; i < n;) {
        const uint64_t hi = key(e[i]) >> 32;
        size_t j = i + 1;
        bool differ = false;
        for (; j < n && (key(e[j]) >> 32) == hi; ++j) differ |= key(e[j]) != key(e[i]);
        if (differ)
            std::stable_sort(e.begin() + i, e.begin() + j, [&](const Edge& x, const Edge& y) { return key(x) < key(y); });
        i = j;
    }
}

// Stable cost sort, greedy merge (cost < min over both components of Int + scale / |C|), min-size pass,
// labels = rank of the root index.  Returns the number of segments.
static int merge_sorted_edges(std::vector<Edge>& edges, size_t n, double sc, int min_size, int32_t* labels_out) {
    Scratch& w = scratch();
    sort_edges(edges, w.edges_tmp);

    std::vector<int>& parent = w.parent;
    std::vector<int>& size = w.size;
    parent.resize(n);
    size.assign(n, 1);
    // limit[root] = Int(C) + scale / |C|, the merge threshold of the component: it changes only when the component does,
    // so it is evaluated at the merge (the same two operations on the same operands) instead of at every test
    std::vector<double>& limit = w.limit;
    limit.assign(n, 0.0 + sc / 1);
    std::iota(parent.begin(), parent.end(), 0);
    for (const Edge& e : edges) {
        const int s0 = find_root(parent, e.a), s1 = find_root(parent, e.b);
        if (s0 == s1) continue;
        if (e.cost < std::min(limit[s0], limit[s1])) {
            const int total = size[s0] + size[s1];
            const int r = join(parent, s0, s1);
            size[r] = total;
            limit[r] = e.cost + sc / total;
        }
    }
    // min-size pass: same edge order; it ends as soon as no component below min_size is left
    size_t small = 0;
    for (size_t i = 0; i < n; ++i) small += parent[i] == (int)i && size[i] < min_size;
    for (const Edge& e : edges) {
        if (small == 0) break;
        const int s0 = find_root(parent, e.a), s1 = find_root(parent, e.b);
        if (s0 == s1) continue;
        if (size[s0] < min_size || size[s1] < min_size) {
            const int total = size[s0] + size[s1];
            small -= (size[s0] < min_size) + (size[s1] < min_size);
            small += total < min_size;
            size[join(parent, s0, s1)] = total;
        }
    }
    // np.unique(root, return_inverse=True)[1]: rank of the root index
    std::vector<int>& rank = w.rank;
    rank.resize(n);
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
    Scratch& ws = scratch();
    std::vector<double>& img = ws.img;
    std::vector<double>& tmp = ws.tmp;
    img.resize(n * 3);
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
    std::vector<Edge>& edges = ws.edges;
    edges.resize((size_t)h * (w - 1) + (size_t)(h - 1) * w + 2 * (size_t)(h - 1) * (w - 1));
    Edge* out = edges.data();
    auto cost = [&](const double* p, const double* q) {
        const double d0 = p[0] - q[0], d1 = p[1] - q[1], d2 = p[2] - q[2];
        return std::sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    };
    const double* im = img.data();
    const size_t rs = (size_t)w * 3;
    for (int r = 0; r < h; ++r) {
        const double* row = im + r * rs;
        for (int c = 1; c < w; ++c) *out++ = {cost(row + 3 * c, row + 3 * c - 3), r * w + c, r * w + c - 1};
    }
    for (int r = 1; r < h; ++r) {
        const double* row = im + r * rs;
        for (int c = 0; c < w; ++c) *out++ = {cost(row + 3 * c, row + 3 * c - rs), r * w + c, (r - 1) * w + c};
    }
    for (int r = 1; r < h; ++r) {
        const double* row = im + r * rs;
        for (int c = 1; c < w; ++c) *out++ = {cost(row + 3 * c, row + 3 * c - rs - 3), r * w + c, (r - 1) * w + c - 1};
    }
    for (int r = 1; r < h; ++r) {
        const double* row = im + r * rs;
        for (int c = 1; c < w; ++c) *out++ = {cost(row + 3 * c - 3, row + 3 * c - rs), (r - 1) * w + c, r * w + c - 1};
    }
    const int count = merge_sorted_edges(edges, n, sc, min_size, labels_out);
    trim_scratch(n);
    return count;
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
    const int count = merge_sorted_edges(edges, (size_t)n_vertices, scale, min_size, labels_out);
    trim_scratch((size_t)n_vertices);
    return count;
}

}  // namespace pcm
