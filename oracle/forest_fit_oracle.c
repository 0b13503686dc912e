/* forest_fit_oracle.c -- TEST INFRASTRUCTURE (oracle).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may build or call this file; the product never does.
 *
 * CPU restatement of how scikit-learn 1.9.0 grows ONE tree of the RandomForestClassifier that the
 * reference fits in addModel (maskers/pixel_classification.py:199-200:
 * RandomForestClassifier(random_state=42, n_estimators, max_depth).fit(X, labels)), for the data this
 * path feeds it: features are integers v in {-1..255} presented as float32(v / 255.0) (:55 `X/255`,
 * sklearn casts to float32), two classes, bootstrap sample counts as integer sample weights.
 * scikit-learn is a third-party dependency of the reference (environment.yaml:14 pins 0.24.1; this image
 * has 1.9.0, whose Cython sources are restated here):
 *   sklearn/tree/_tree.pyx        DepthFirstTreeBuilder.build        (:150-340)  stack order, leaf tests
 *   sklearn/tree/_splitter.pyx    node_split_best                    (:262-504)  Fisher-Yates feature draw,
 *                                                                                 constant-feature bookkeeping
 *   sklearn/tree/_partitioner.pyx DensePartitioner.next_p / FEATURE_THRESHOLD = 1e-7
 *   sklearn/tree/_criterion.pyx   Gini.node_impurity / children_impurity (:620-688),
 *                                 Criterion.proxy_impurity_improvement / impurity_improvement (:147-199),
 *                                 ClassificationCriterion.node_value (:473-486)
 *   sklearn/utils/_random.pxd     our_rand_r (:20-34), sklearn/tree/_utils.pyx rand_int (:51-54)
 * Pinned by tests/test_forest_fit_oracle.py against scikit-learn itself (tree_ arrays equal).
 *
 * Because a feature takes at most 257 distinct values, sorting the node's samples by feature value
 * (what scikit-learn does) and scanning the sorted run is restated as a 257-bin histogram of weighted
 * class counts scanned in ascending value order: the candidate positions, the weighted sums at each of
 * them (integers, exact in float64) and therefore every float64 expression are the same.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o _build/libforest_fit_oracle.so forest_fit_oracle.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NBINS 257

static uint32_t our_rand_r(uint32_t* seed) {
    if (*seed == 0) *seed = 1;
    *seed ^= (uint32_t)(*seed << 13);
    *seed ^= (uint32_t)(*seed >> 17);
    *seed ^= (uint32_t)(*seed << 5);
    return *seed % ((uint32_t)2147483647 + 1);
}
static long rand_int(long low, long high, uint32_t* st) { return low + (long)(our_rand_r(st) % (uint32_t)(high - low)); }

typedef struct {
    int start, end, depth, parent, is_left, n_constant;
    double impurity;
} StackRec;

static double gini(double c0, double c1, double w) {
    double sq = 0.0;
    sq += c0 * c0;
    sq += c1 * c1;
    double g = 0.0;
    g += 1.0 - sq / (w * w);
    return g / 1.0;
}

/* X: [n][F] int16 row-major (-1..255); y: [n] 0/1; w: [n] bootstrap counts (>= 0).
 * Outputs (capacity max_nodes): feature (-2 leaf), threshold (-2.0 leaf), left/right (-1 leaf),
 * value0/value1 = class fractions of the node, n_node_samples, weighted_n, impurity.
 * Returns the node count, or -1 when max_nodes is too small. */
int forest_fit_oracle_tree(const int16_t* X, const uint8_t* y, const int32_t* w, int n, int F, int max_depth,
                           int max_features, uint32_t rand_r_state, int max_nodes, int32_t* feature, double* threshold,
                           int32_t* left, int32_t* right, double* value0, double* value1, int32_t* n_node_samples,
                           double* weighted_n, double* impurity_out) {
    const double EPSILON = 2.220446049250313e-16;     /* np.finfo('double').eps, _tree.pyx:42 */
    float fv[NBINS];
    for (int b = 0; b < NBINS; ++b) fv[b] = (float)((double)(b - 1) / 255.0);
    int* samples = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int ns = 0;
    double weighted_n_samples = 0.0;
    for (int i = 0; i < n; ++i) {
        if (w[i] != 0) samples[ns++] = i;
        weighted_n_samples += (double)w[i];
    }
    long* features = (long*)malloc(sizeof(long) * (size_t)F);
    long* constant_features = (long*)malloc(sizeof(long) * (size_t)F);
    for (int f = 0; f < F; ++f) features[f] = f;
    StackRec* stack = (StackRec*)malloc(sizeof(StackRec) * (size_t)(2 * (max_depth > 0 ? max_depth : 1) + 8));
    int sp = 0, node_count = 0, first = 1, rc = 0;
    stack[sp++] = (StackRec){0, ns, 0, -1, 0, 0, INFINITY};
    double h0[NBINS], h1[NBINS];
    int hc[NBINS];
    while (sp > 0) {
        StackRec r = stack[--sp];
        const int start = r.start, end = r.end;
        const int nn = end - start;
        double t0 = 0.0, t1 = 0.0, W = 0.0;
        for (int p = start; p < end; ++p) {               /* Criterion.init: sums in sample order (integers: exact) */
            const int i = samples[p];
            if (y[i]) t1 += (double)w[i]; else t0 += (double)w[i];
            W += (double)w[i];
        }
        int is_leaf = (r.depth >= max_depth) || (nn < 2) || (nn < 2 * 1) || (W < 2 * 0.0);
        double impurity = r.impurity;
        if (first) { impurity = gini(t0, t1, W); first = 0; }
        is_leaf = is_leaf || impurity <= EPSILON;

        int best_feature = -1, best_bin = -1, best_found = 0;     /* best_bin: last bin that goes left */
        double best_thr = 0.0, best_proxy = -INFINITY;
        double best_l0 = 0, best_l1 = 0;
        int n_total_constants = r.n_constant;
        if (!is_leaf) {
            long f_i = F, f_j;
            int n_visited = 0, n_found = 0, n_drawn = 0;
            const int n_known = r.n_constant;
            while (f_i > n_total_constants && (n_visited < max_features || n_visited <= n_found + n_drawn)) {
                n_visited++;
                f_j = rand_int(n_drawn, f_i - n_found, &rand_r_state);
                if (f_j < n_known) {
                    long t = features[n_drawn]; features[n_drawn] = features[f_j]; features[f_j] = t;
                    n_drawn++;
                    continue;
                }
                f_j += n_found;
                const long cf = features[f_j];
                memset(h0, 0, sizeof h0); memset(h1, 0, sizeof h1); memset(hc, 0, sizeof hc);
                int lo = NBINS, hi = -1;
                for (int p = start; p < end; ++p) {
                    const int i = samples[p];
                    const int b = (int)X[(size_t)i * F + cf] + 1;
                    if (y[i]) h1[b] += (double)w[i]; else h0[b] += (double)w[i];
                    hc[b]++;
                    if (b < lo) lo = b;
                    if (b > hi) hi = b;
                }
                if (fv[hi] <= fv[lo] + 1e-7f) {           /* constant in this node (FEATURE_THRESHOLD) */
                    long t = features[f_j]; features[f_j] = features[n_total_constants]; features[n_total_constants] = t;
                    n_found++;
                    n_total_constants++;
                    continue;
                }
                f_i--;
                { long t = features[f_i]; features[f_i] = features[f_j]; features[f_j] = t; }
                /* scan: position p after the run of equal values in bin b, candidate iff a later bin is non-empty */
                double l0 = 0.0, l1 = 0.0;
                int prev = -1;
                for (int b = lo; b <= hi; ++b) {
                    if (!hc[b]) continue;
                    if (prev >= 0) {
                        /* split between bins prev and b: left = everything up to prev */
                        const double wl = l0 + l1, r0 = t0 - l0, r1 = t1 - l1, wr = W - wl;
                        const double gl = gini(l0, l1, wl), gr = gini(r0, r1, wr);
                        const double proxy = (-wr * gr) - (wl * gl);
                        if (proxy > best_proxy) {
                            best_proxy = proxy;
                            best_feature = (int)cf;
                            best_bin = prev;
                            best_thr = (double)fv[prev] / 2.0 + (double)fv[b] / 2.0;
                            best_l0 = l0; best_l1 = l1;
                            best_found = 1;
                        }
                    }
                    l0 += h0[b]; l1 += h1[b];
                    prev = b;
                }
            }
            memcpy(features, constant_features, sizeof(long) * (size_t)n_known);
            memcpy(constant_features + n_known, features + n_known, sizeof(long) * (size_t)n_found);
        }
        double imp_l = 0, imp_r = 0, improvement = 0;
        int pos = end;
        if (!is_leaf && best_found) {
            /* partition_samples_final: X[sample, feature] <= threshold goes left */
            int p = start, q = end;
            while (p < q) {
                if ((int)X[(size_t)samples[p] * F + best_feature] + 1 <= best_bin) p++;
                else { q--; int t = samples[p]; samples[p] = samples[q]; samples[q] = t; }
            }
            pos = p;
            const double wl = best_l0 + best_l1, wr = W - wl;
            imp_l = gini(best_l0, best_l1, wl);
            imp_r = gini(t0 - best_l0, t1 - best_l1, wr);
            improvement = (W / weighted_n_samples) * (impurity - (wr / W * imp_r) - (wl / W * imp_l));
        }
        if (!is_leaf) is_leaf = (pos >= end) || (improvement + EPSILON < 0.0);

        if (node_count >= max_nodes) { rc = -1; break; }
        const int id = node_count++;
        if (r.parent >= 0) { if (r.is_left) left[r.parent] = id; else right[r.parent] = id; }
        left[id] = right[id] = -1;
        feature[id] = is_leaf ? -2 : best_feature;
        threshold[id] = is_leaf ? -2.0 : best_thr;
        value0[id] = t0 / W;
        value1[id] = t1 / W;
        n_node_samples[id] = nn;
        weighted_n[id] = W;
        impurity_out[id] = impurity;
        if (!is_leaf) {
            stack[sp++] = (StackRec){pos, end, r.depth + 1, id, 0, n_total_constants, imp_r};
            stack[sp++] = (StackRec){start, pos, r.depth + 1, id, 1, n_total_constants, imp_l};
        }
    }
    free(samples); free(features); free(constant_features); free(stack);
    return rc < 0 ? -1 : node_count;
}
