// pcm_kernels.cuh -- CUDA kernels of the PC masker hot path (sm_100a).
//
//  K1 score_kernel        BGR crop -> HSV/LAB planes (smem) -> star taps -> forest(s)
//                         [-> PCA novelty error] [-> temporal blend] -> P(fg) f64 [, err f64]
//                         replaces cvtColor + getFeatures + X/255 + predict_proba + PCA
//                         (reference maskers/pixel_classification.py:53-63, :80-95)
//  K2 segment_reduce      per-label sums / areas (np.unique :97 + first loop of
//                         compileSaliencyMap :235-238), warp-aggregated atomics
//  K2b segment_decide     per-label score and decision (:240-242) + guard band
//  K2c segment_resolve    exact sequential-float32 re-evaluation of guard-band labels
//  K3 mask_dilate         decision -> 0/255 map (:242-246) fused with cv.dilate (:112)
//  K5 iou_kernel          computeBenchmark counts (benchmark.py:8-14)
//  + convert_kernel / gather_kernel: parity taps (pcm_convert, pcm_gather_features)
#pragma once
#include "pcm_device.cuh"

namespace pcm {

// ------------------------------------------------------------------------------
// packed forest (built on the host by encode_forest in pcm_api.cu)
//   Every tree owns a contiguous run of 8-byte entries: its internal nodes followed
//   by one self-looping pseudo-node per leaf, so that the traversal is branch-free
//   and a thread that has reached a leaf simply stays there.
//   node.x = thr << 24 | tap byte offset inside the plane tile (24 bits)
//   node.y = left | right << 16, BYTE offsets of the children from the tree's base
//   leaf pseudo-node: x = 0xff000000 (never goes right), y = self | self << 16
//   trees[t] = {node base (bytes), leaf-value base (bytes, biased so that
//               value address = base + entry offset), root offset, depth}
// A tap value v (u8, 0 outside the crop) goes RIGHT iff v > thr  <=>  (v << 24) > node.x.
// The crop-border sentinel -1 of the reference (:263) is handled by the encoder:
// nodes with integer threshold -1 test the validity plane instead (0 outside the crop).
// ------------------------------------------------------------------------------
struct DevForest {
    const uint2* nodes;
    const double* leaves;
    const int4* trees;
    int n_trees, n_nodes, n_leaves;
};

struct DevPCA {
    const double* comp;      // components_[0][F]
    const double* comp255;   // components_[0][F] / 255
    const double* mean;      // mean_[F]
    double mean_dot_comp;    // mean_ . components_[0]
};

struct ScoreArgs {
    const uint8_t* frame;            // BGR, rows `stride` bytes apart
    const uint8_t* frame_lo;         // first / one-past-last readable byte of the frame
    const uint8_t* frame_hi;
    long long stride;
    int cx, cy, cw, ch;              // crop rectangle (frame coords)
    int tiles_x, tiles_y;
    Geom g;
    const ColorTables* tables;
    DevForest f0, f1;
    int blend;                       // 0/1: f1 (and pca1) valid
    double w0, w1;                   // np.average weights
    int novelty;                     // 0/1
    DevPCA pca0, pca1;
    double* p1_out;                  // [ch*cw]
    double* sa_out;                  // [ch*cw] (novelty only)
};

__device__ __forceinline__ uint32_t ldg_word_checked(const uint8_t* wp, const uint8_t* lo, const uint8_t* hi) {
    if (wp >= lo && wp + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t*>(wp));
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (wp + i >= lo && wp + i < hi) w |= (uint32_t)__ldg(wp + i) << (8 * i);
    return w;
}

// One forest over the thread's PIX_PER_THREAD pixels; leaf fractions are added in
// estimator order (sklearn ensemble/_forest.py: all_proba += prediction).
// Branch-free: PIX_PER_THREAD independent pointer chases are interleaved by the
// compiler, which is what hides the two dependent shared-memory loads per visit.
__device__ __forceinline__ void traverse_forest(const uint8_t* __restrict__ nodes, const uint8_t* __restrict__ leaves,
                                                const int4* __restrict__ trees, int n_trees,
                                                const uint8_t* __restrict__ pixbase, int RS,
                                                double (&acc)[PIX_PER_THREAD]) {
    for (int t = 0; t < n_trees; ++t) {
        const int4 ti = trees[t];
        const uint8_t* nb = nodes + ti.x;
        const uint8_t* lb = leaves + ti.y;
        unsigned ref[PIX_PER_THREAD];
#pragma unroll
        for (int g = 0; g < PIX_PER_THREAD; ++g) ref[g] = (unsigned)ti.z;
        for (int lvl = 0; lvl < ti.w; ++lvl) {
#pragma unroll
            for (int g = 0; g < PIX_PER_THREAD; ++g) {
                const uint2 nd = *reinterpret_cast<const uint2*>(nb + ref[g]);
                const unsigned v = pixbase[g * RS + (nd.x & 0xffffffu)];
                ref[g] = __byte_perm(nd.y, 0u, ((v << 24) > nd.x) ? 0x4432u : 0x4410u);
            }
        }
#pragma unroll
        for (int g = 0; g < PIX_PER_THREAD; ++g)
            acc[g] = __dadd_rn(acc[g], *reinterpret_cast<const double*>(lb + ref[g]));
    }
}

// L1 reconstruction error of a rank-1 PCA over the star features (:58-60):
//   t = sum_f x_f c_f - mean.c ;  err = sum_f |x_f - (t c_f + mean_f)|,  x_f = v_f / 255, v_f = -1 off-crop
__device__ __forceinline__ void novelty_error(const double* __restrict__ comp, const double* __restrict__ comp255,
                                              const double* __restrict__ mean, const double mdc,
                                              const int* __restrict__ sp,
                                              const Geom& g, const uint8_t* __restrict__ pixbase,
                                              double (&err)[PIX_PER_THREAD]) {
    const int nch = 3 * g.n_spaces;
    const int vplane = nch * g.PS;
    double t[PIX_PER_THREAD];
#pragma unroll
    for (int i = 0; i < PIX_PER_THREAD; ++i) t[i] = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int k = 0; k < g.K; ++k) {
            const int so = sp[k];
            unsigned ok = 0;
#pragma unroll
            for (int i = 0; i < PIX_PER_THREAD; ++i) ok |= (unsigned)(pixbase[i * g.RS + vplane + so] != 0) << i;
            for (int p = 0; p < nch; ++p) {
                const int f = (p / 3) * 3 * g.K + 3 * k + (p % 3);
                const uint8_t* src = pixbase + p * g.PS + so;
                if (pass == 0) {
                    const double c255 = comp255[f];
#pragma unroll
                    for (int i = 0; i < PIX_PER_THREAD; ++i) {
                        const double v = ((ok >> i) & 1u) ? u8_to_double(src[i * g.RS]) : -1.0;
                        t[i] = fma(v, c255, t[i]);
                    }
                } else {
                    const double c = comp[f], mu = mean[f];
#pragma unroll
                    for (int i = 0; i < PIX_PER_THREAD; ++i) {
                        const double v = ((ok >> i) & 1u) ? u8_to_double(src[i * g.RS]) : -1.0;
                        err[i] += fabs(fma(v, 1.0 / 255.0, -fma(t[i], c, mu)));
                    }
                }
            }
        }
        if (pass == 0) {
#pragma unroll
            for (int i = 0; i < PIX_PER_THREAD; ++i) t[i] -= mdc;
        }
    }
}

__device__ __forceinline__ double blend2(double a, double b, double w0, double w1) {
    // np.average([a, b], weights=[w0, w1]) = (a*w0 + b*w1) / (w0 + w1), no contraction
    return __ddiv_rn(__dadd_rn(__dmul_rn(a, w0), __dmul_rn(b, w1)), __dadd_rn(w0, w1));
}

struct ScoreSmem {
    uint32_t planes, raw, tables, sp, f0_nodes, f0_leaves, f0_trees, f1_nodes, f1_leaves, f1_trees,
        pca0, pca1, total;
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// Shared-memory carve-up, identical on host (sizing) and device.
__host__ __device__ inline ScoreSmem score_smem_layout(const Geom& g, const DevForest& f0, const DevForest& f1,
                                                       bool blend, bool novelty, bool forest_smem) {
    ScoreSmem s;
    uint32_t o = 0;
    s.planes = o; o = align_up(o + g.n_planes * g.PS, 16);
    s.raw = o;    o = align_up(o + g.PH * g.RAWS, 16);
    s.tables = o; o = align_up(o + (uint32_t)sizeof(ColorTables), 16);
    s.sp = o;     o = align_up(o + 4 * g.K, 16);
    s.f0_nodes = s.f0_leaves = s.f0_trees = s.f1_nodes = s.f1_leaves = s.f1_trees = 0;
    if (forest_smem) {
        s.f0_nodes = o;  o = align_up(o + 8 * f0.n_nodes, 16);
        s.f0_leaves = o; o = align_up(o + 8 * f0.n_leaves, 16);
        s.f0_trees = o;  o = align_up(o + 16 * f0.n_trees, 16);
        if (blend) {
            s.f1_nodes = o;  o = align_up(o + 8 * f1.n_nodes, 16);
            s.f1_leaves = o; o = align_up(o + 8 * f1.n_leaves, 16);
            s.f1_trees = o;  o = align_up(o + 16 * f1.n_trees, 16);
        }
    }
    s.pca0 = s.pca1 = 0;
    if (novelty) {
        s.pca0 = o; o = align_up(o + 24 * g.F, 16);
        if (blend) { s.pca1 = o; o = align_up(o + 24 * g.F, 16); }
    }
    s.total = o;
    return s;
}

template <typename T>
__device__ __forceinline__ void copy_to_smem(T* dst, const T* __restrict__ src, int n) {
    for (int i = threadIdx.x; i < n; i += NTHREADS) dst[i] = src[i];
}

// K1 -----------------------------------------------------------------------------
template <bool FOREST_SMEM>
__global__ void __launch_bounds__(NTHREADS) score_kernel(const ScoreArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const Geom& g = a.g;
    const ScoreSmem L = score_smem_layout(g, a.f0, a.f1, a.blend != 0, a.novelty != 0, FOREST_SMEM);
    uint8_t* planes = smem + L.planes;
    uint8_t* raw = smem + L.raw;
    ColorTables* tab = reinterpret_cast<ColorTables*>(smem + L.tables);
    int* sp = reinterpret_cast<int*>(smem + L.sp);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- once per CTA: tables, tap offsets, forests, PCA vectors -> smem ------------
    copy_to_smem(reinterpret_cast<uint32_t*>(tab), reinterpret_cast<const uint32_t*>(a.tables),
                 (int)(sizeof(ColorTables) / 4));
    for (int k = tid; k < g.K; k += NTHREADS) {
        int dr, dc;
        star_tap(k, dr, dc);
        sp[k] = (dr + g.n) * g.RS + (dc + g.n);
    }
    const uint8_t* f0n = reinterpret_cast<const uint8_t*>(a.f0.nodes);
    const uint8_t* f0l = reinterpret_cast<const uint8_t*>(a.f0.leaves);
    const int4* f0t = a.f0.trees;
    const uint8_t* f1n = reinterpret_cast<const uint8_t*>(a.f1.nodes);
    const uint8_t* f1l = reinterpret_cast<const uint8_t*>(a.f1.leaves);
    const int4* f1t = a.f1.trees;
    if (FOREST_SMEM) {
        copy_to_smem(reinterpret_cast<uint2*>(smem + L.f0_nodes), a.f0.nodes, a.f0.n_nodes);
        copy_to_smem(reinterpret_cast<double*>(smem + L.f0_leaves), a.f0.leaves, a.f0.n_leaves);
        copy_to_smem(reinterpret_cast<int4*>(smem + L.f0_trees), a.f0.trees, a.f0.n_trees);
        f0n = smem + L.f0_nodes;
        f0l = smem + L.f0_leaves;
        f0t = reinterpret_cast<const int4*>(smem + L.f0_trees);
        if (a.blend) {
            copy_to_smem(reinterpret_cast<uint2*>(smem + L.f1_nodes), a.f1.nodes, a.f1.n_nodes);
            copy_to_smem(reinterpret_cast<double*>(smem + L.f1_leaves), a.f1.leaves, a.f1.n_leaves);
            copy_to_smem(reinterpret_cast<int4*>(smem + L.f1_trees), a.f1.trees, a.f1.n_trees);
            f1n = smem + L.f1_nodes;
            f1l = smem + L.f1_leaves;
            f1t = reinterpret_cast<const int4*>(smem + L.f1_trees);
        }
    }
    const double *p0c = nullptr, *p0c255 = nullptr, *p0m = nullptr, *p1c = nullptr, *p1c255 = nullptr, *p1m = nullptr;
    if (a.novelty) {
        double* d0 = reinterpret_cast<double*>(smem + L.pca0);
        copy_to_smem(d0, a.pca0.comp, g.F);
        copy_to_smem(d0 + g.F, a.pca0.comp255, g.F);
        copy_to_smem(d0 + 2 * g.F, a.pca0.mean, g.F);
        p0c = d0; p0c255 = d0 + g.F; p0m = d0 + 2 * g.F;
        if (a.blend) {
            double* d1 = reinterpret_cast<double*>(smem + L.pca1);
            copy_to_smem(d1, a.pca1.comp, g.F);
            copy_to_smem(d1 + g.F, a.pca1.comp255, g.F);
            copy_to_smem(d1 + 2 * g.F, a.pca1.mean, g.F);
            p1c = d1; p1c255 = d1 + g.F; p1m = d1 + 2 * g.F;
        }
    }
    __syncthreads();

    const int n_tiles = a.tiles_x * a.tiles_y;
    const int nch = 3 * g.n_spaces;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx0 = (tile % a.tiles_x) * TILE_W;
        const int ty0 = (tile / a.tiles_x) * TILE_H;
        const int lo = max(tx0 - g.n, 0);
        const int hi = min(tx0 - g.n + g.PW, a.cw);

        // ---- stage 1: warp w stages and converts plane rows w, w+8, ... ----------------
        for (int r = warp; r < g.PH; r += NTHREADS / 32) {
            const int gy = ty0 - g.n + r;
            const bool rowok = (gy >= 0) && (gy < a.ch) && (hi > lo);
            int m = 0;
            uint8_t* rawrow = raw + r * g.RAWS;
            if (rowok) {
                const uint8_t* a0 = a.frame + (long long)(a.cy + gy) * a.stride + (long long)(a.cx + lo) * 3;
                m = (int)(reinterpret_cast<uintptr_t>(a0) & 3);
                const uint8_t* A0 = a0 - m;
                const int nwords = (m + (hi - lo) * 3 + 3) >> 2;
                for (int j = lane; j < nwords; j += 32)
                    reinterpret_cast<uint32_t*>(rawrow)[j] = ldg_word_checked(A0 + 4 * j, a.frame_lo, a.frame_hi);
            }
            __syncwarp();
            for (int c = lane; c < g.PW; c += 32) {
                const int gx = tx0 - g.n + c;
                uint8_t* dst = planes + r * g.RS + c;
                if (rowok && gx >= lo && gx < hi) {
                    const uint8_t* px = rawrow + m + (gx - lo) * 3;
                    const int b = px[0], gg = px[1], rr = px[2];
                    for (int q = 0; q < g.n_spaces; ++q) {
                        int c0, c1, c2;
                        const int sid = g.space_id[q];
                        if (sid == 1) bgr2hsv_px(b, gg, rr, tab->sdiv, tab->hdiv, c0, c1, c2);
                        else if (sid == 2) bgr2lab_px(b, gg, rr, tab->gamma, tab->cbrt_tab, c0, c1, c2);
                        else { c0 = b; c1 = gg; c2 = rr; }
                        dst[(3 * q + 0) * g.PS] = (uint8_t)c0;
                        dst[(3 * q + 1) * g.PS] = (uint8_t)c1;
                        dst[(3 * q + 2) * g.PS] = (uint8_t)c2;
                    }
                    dst[nch * g.PS] = 1;
                } else {
                    for (int p = 0; p <= nch; ++p) dst[p * g.PS] = 0;
                }
            }
        }
        __syncthreads();

        // ---- stage 2: forests (+ novelty) for 8 rows of one column per thread -----------
        const int col = (warp & 1) * 32 + lane;
        const int row0 = (warp >> 1) * PIX_PER_THREAD;
        const int ox = tx0 + col;
        unsigned active = 0;
#pragma unroll
        for (int i = 0; i < PIX_PER_THREAD; ++i)
            active |= (unsigned)((ox < a.cw) && (ty0 + row0 + i < a.ch)) << i;
        const uint8_t* pixbase = planes + row0 * g.RS + col;

        double p[PIX_PER_THREAD];
#pragma unroll
        for (int i = 0; i < PIX_PER_THREAD; ++i) p[i] = 0.0;
        traverse_forest(f0n, f0l, f0t, a.f0.n_trees, pixbase, g.RS, p);
        const double T0 = (double)a.f0.n_trees;
#pragma unroll
        for (int i = 0; i < PIX_PER_THREAD; ++i) p[i] = __ddiv_rn(p[i], T0);
        if (a.blend) {
            double q[PIX_PER_THREAD];
#pragma unroll
            for (int i = 0; i < PIX_PER_THREAD; ++i) q[i] = 0.0;
            traverse_forest(f1n, f1l, f1t, a.f1.n_trees, pixbase, g.RS, q);
            const double T1 = (double)a.f1.n_trees;
#pragma unroll
            for (int i = 0; i < PIX_PER_THREAD; ++i) p[i] = blend2(p[i], __ddiv_rn(q[i], T1), a.w0, a.w1);
        }
#pragma unroll
        for (int i = 0; i < PIX_PER_THREAD; ++i)
            if ((active >> i) & 1u) a.p1_out[(size_t)(ty0 + row0 + i) * a.cw + ox] = p[i];

        if (a.novelty) {
            double e[PIX_PER_THREAD];
#pragma unroll
            for (int i = 0; i < PIX_PER_THREAD; ++i) e[i] = 0.0;
            novelty_error(p0c, p0c255, p0m, a.pca0.mean_dot_comp, sp, g, pixbase, e);
            if (a.blend) {
                double e1[PIX_PER_THREAD];
#pragma unroll
                for (int i = 0; i < PIX_PER_THREAD; ++i) e1[i] = 0.0;
                novelty_error(p1c, p1c255, p1m, a.pca1.mean_dot_comp, sp, g, pixbase, e1);
#pragma unroll
                for (int i = 0; i < PIX_PER_THREAD; ++i) e[i] = blend2(e[i], e1[i], a.w0, a.w1);
            }
#pragma unroll
            for (int i = 0; i < PIX_PER_THREAD; ++i)
                if ((active >> i) & 1u) a.sa_out[(size_t)(ty0 + row0 + i) * a.cw + ox] = e[i];
        }
        __syncthreads();   // planes / raw are rewritten by the next tile
    }
}

// K2 -----------------------------------------------------------------------------
// d = p1 - (max(sa, thr) - thr) per pixel (:237); per label: sum d, sum |d|, area,
// first/last row.  One pixel per lane; lanes with equal labels are combined with
// shuffles so that a warp issues one set of atomics per distinct label.
struct SegArgs {
    const double* p1;
    const double* sa;          // nullptr when novelty is off (sa == 0, thr == 0)
    const int32_t* labels;
    int n_px, cw, n_labels;
    double thr;
    double* sum;               // [S]
    double* asum;              // [S]
    int* area;                 // [S]
    int* rmin;                 // [S] init INT_MAX
    int* rmax;                 // [S] init -1
    int* err;                  // set to 1 on an out-of-range label
};

__device__ __forceinline__ double contribution(double p1, double sa, double thr) {
    return __dsub_rn(p1, __dsub_rn(fmax(sa, thr), thr));
}

__global__ void __launch_bounds__(256) segment_reduce_kernel(const SegArgs a) {
    const int lane = threadIdx.x & 31;
    const int n_warps_total = (gridDim.x * blockDim.x) >> 5;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int base = warp_global * 32; base < a.n_px; base += n_warps_total * 32) {
        const int idx = base + lane;
        const bool in = idx < a.n_px;
        int lab = -1;
        double d = 0.0;
        int row = 0;
        if (in) {
            lab = a.labels[idx];
            if (lab < 0 || lab >= a.n_labels) { *a.err = 1; lab = -1; }
            else {
                d = contribution(a.p1[idx], a.sa ? a.sa[idx] : 0.0, a.thr);
                row = idx / a.cw;
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, lab >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int L = __shfl_sync(0xffffffffu, lab, leader);
            const bool mine = (lab == L);
            const unsigned members = __ballot_sync(0xffffffffu, mine);
            double s = mine ? d : 0.0, as = mine ? fabs(d) : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                as += __shfl_xor_sync(0xffffffffu, as, o);
            }
            const int r0 = __reduce_min_sync(0xffffffffu, mine ? row : 0x7fffffff);
            const int r1 = __reduce_max_sync(0xffffffffu, mine ? row : -1);
            if (lane == leader) {
                atomicAdd(a.sum + L, s);
                atomicAdd(a.asum + L, as);
                atomicAdd(a.area + L, __popc(members));
                atomicMin(a.rmin + L, r0);
                atomicMax(a.rmax + L, r1);
            }
            todo &= ~members;
        }
    }
}

// K2b ----------------------------------------------------------------------------
// score = f32( (acc/area) * (1-w) + prior * w ) > 0.5 (:241-242), acc being the
// reference's sequential float32 accumulator.  The parallel float64 sum differs from
// it by at most 2^-24 * sum|d| per unit area; labels whose score is that close to
// 0.5 are queued for the exact path (K2c), every other label is decided here.
struct DecideArgs {
    const double* sum;
    const double* asum;
    const int* area;
    const float* priors;       // nullptr -> all -1
    int n_labels;
    double prior_weight;
    uint8_t* decision;         // [S] 0/1
    float* scores;             // [S]
    int* flagged;              // [S] queue
    int* n_flagged;
};

__global__ void __launch_bounds__(256) segment_decide_kernel(const DecideArgs a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_labels) return;
    const int area = a.area[s];
    if (area <= 0) { a.decision[s] = 0; a.scores[s] = 0.f; return; }
    const double w = a.prior_weight;
    const double prior = a.priors ? (double)a.priors[s] : -1.0;
    const double omw = __dsub_rn(1.0, w);
    const double sc = __dadd_rn(__dmul_rn(__ddiv_rn(a.sum[s], (double)area), omw), __dmul_rn(prior, w));
    const double band = 2.0 * 5.9604644775390625e-08 * (a.asum[s] / (double)area) * fabs(omw) + 2.4e-7;
    if (fabs(sc - 0.5) <= band) {
        a.flagged[atomicAdd(a.n_flagged, 1)] = s;
        a.decision[s] = 0;
    } else {
        a.decision[s] = sc > 0.5;
    }
    a.scores[s] = (float)sc;
}

// K2c ----------------------------------------------------------------------------
// Exact compileSaliencyMap accumulation for one label: float32 accumulator, each
// `+=` evaluated in float64 and rounded to float32, pixels in raster order (:235-238).
struct ResolveArgs {
    const double* p1;
    const double* sa;
    const int32_t* labels;
    int cw;
    double thr;
    const int* area;
    const int* rmin;
    const int* rmax;
    const float* priors;
    double prior_weight;
    const int* flagged;
    const int* n_flagged;
    uint8_t* decision;
    float* scores;
};

__global__ void __launch_bounds__(256) segment_resolve_kernel(const ResolveArgs a) {
    __shared__ double buf[256];
    __shared__ int warp_count[8];
    const int n = *a.n_flagged;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int L = a.flagged[q];
        const int lo = a.rmin[L] * a.cw, hi = (a.rmax[L] + 1) * a.cw;
        float acc = 0.f;
        for (int base = lo; base < hi; base += 256) {
            const int idx = base + threadIdx.x;
            const bool mine = (idx < hi) && (a.labels[idx] == L);
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (lane == 0) warp_count[warp] = __popc(bal);
            __syncthreads();
            int off = 0, total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (w < warp) off += warp_count[w];
                total += warp_count[w];
            }
            if (mine)
                buf[off + __popc(bal & ((1u << lane) - 1u))] =
                    contribution(a.p1[idx], a.sa ? a.sa[idx] : 0.0, a.thr);
            __syncthreads();
            if (threadIdx.x == 0)
                for (int i = 0; i < total; ++i) acc = __double2float_rn(__dadd_rn((double)acc, buf[i]));
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double w = a.prior_weight;
            const double prior = a.priors ? (double)a.priors[L] : -1.0;
            const float sc = __double2float_rn(__dadd_rn(
                __dmul_rn(__ddiv_rn((double)acc, (double)a.area[L]), __dsub_rn(1.0, w)), __dmul_rn(prior, w)));
            a.scores[L] = sc;
            a.decision[L] = sc > 0.5f;
        }
    }
}

// K3 -----------------------------------------------------------------------------
// map[r][c] = 255 * decision[label[r][c]] (:242-245) followed by cv.dilate with a
// k x k box, anchor (k/2, k/2), neighbours outside the crop ignored (:112).
struct DilateArgs {
    const int32_t* labels;
    const uint8_t* decision;
    int cw, ch, k, n_labels;
    uint8_t* mask;             // dense plane; element (cy + r, cx + c)
    long long mask_stride;
    int cx, cy;
    uint8_t* pre;              // optional [ch*cw] pre-dilation map
};

constexpr int DIL_TW = 64, DIL_TH = 32, DIL_MAXK = 33;

__global__ void __launch_bounds__(256) mask_dilate_kernel(const DilateArgs a) {
    __shared__ uint8_t s0[(DIL_TH + DIL_MAXK) * (DIL_TW + DIL_MAXK)];
    __shared__ uint8_t s1[(DIL_TH + DIL_MAXK) * DIL_TW];
    const int k = a.k, before = k / 2, SW = DIL_TW + k - 1, SH = DIL_TH + k - 1;
    const int tx0 = blockIdx.x * DIL_TW, ty0 = blockIdx.y * DIL_TH;
    for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) {
        const int r = i / SW, c = i - r * SW;
        const int y = ty0 - before + r, x = tx0 - before + c;
        uint8_t v = 0;
        if (y >= 0 && y < a.ch && x >= 0 && x < a.cw) {
            const int lab = a.labels[(size_t)y * a.cw + x];
            v = (lab >= 0 && lab < a.n_labels) ? a.decision[lab] : 0;
            if (a.pre && r >= before && r < before + DIL_TH && c >= before && c < before + DIL_TW)
                a.pre[(size_t)y * a.cw + x] = v ? 255 : 0;
        }
        s0[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SH * DIL_TW; i += blockDim.x) {
        const int r = i / DIL_TW, c = i - r * DIL_TW;
        uint8_t v = 0;
        for (int d = 0; d < k; ++d) v |= s0[r * SW + c + d];
        s1[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DIL_TH * DIL_TW; i += blockDim.x) {
        const int r = i / DIL_TW, c = i - r * DIL_TW;
        const int y = ty0 + r, x = tx0 + c;
        if (y < a.ch && x < a.cw) {
            uint8_t v = 0;
            for (int d = 0; d < k; ++d) v |= s1[(r + d) * DIL_TW + c];
            a.mask[(size_t)(a.cy + y) * a.mask_stride + (a.cx + x)] = v ? 255 : 0;
        }
    }
}

// K5 -----------------------------------------------------------------------------
// counts[0] += #(m != 0 && t != 0), counts[1] += #(m != 0 || t != 0) (benchmark.py:12-13);
// truth either gray or BGR (converted like cv.cvtColor(BGR2GRAY), main.py:285).
__global__ void __launch_bounds__(256) iou_kernel(const uint8_t* __restrict__ mask, long long mask_stride,
                                                  const uint8_t* __restrict__ truth, long long truth_stride,
                                                  int truth_channels, int h, int w,
                                                  unsigned long long* __restrict__ counts) {
    unsigned inter = 0, uni = 0;
    const int chunks_per_row = (w + 31) / 32;
    const long long n_chunks = (long long)h * chunks_per_row;
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long ch = warp_global; ch < n_chunks; ch += n_warps) {
        const int r = (int)(ch / chunks_per_row);
        const int c = (int)(ch - (long long)r * chunks_per_row) * 32 + lane;
        bool m = false, t = false;
        if (c < w) {
            m = mask[(size_t)r * mask_stride + c] != 0;
            const uint8_t* tp = truth + (size_t)r * truth_stride + (size_t)c * truth_channels;
            t = (truth_channels == 3) ? (bgr2gray_px(tp[0], tp[1], tp[2]) != 0) : (tp[0] != 0);
        }
        inter += __popc(__ballot_sync(0xffffffffu, m && t));
        uni += __popc(__ballot_sync(0xffffffffu, m || t));
    }
    // every lane of a warp holds the same totals; one atomic pair per block
    __shared__ unsigned s_i[8], s_u[8];
    if (lane == 0) { s_i[threadIdx.x >> 5] = inter; s_u[threadIdx.x >> 5] = uni; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ti = 0, tu = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ti += s_i[i]; tu += s_u[i]; }
        if (ti) atomicAdd(counts, ti);
        if (tu) atomicAdd(counts + 1, tu);
    }
}

// parity taps ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) convert_kernel(const uint8_t* __restrict__ bgr, long long stride, int h, int w,
                                                      int space, const ColorTables* __restrict__ tab,
                                                      uint8_t* __restrict__ out, long long out_stride) {
    const long long n = (long long)h * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / w), c = (int)(i - (long long)r * w);
        const uint8_t* px = bgr + (size_t)r * stride + 3 * (size_t)c;
        int c0, c1, c2;
        if (space == 1) bgr2hsv_px(px[0], px[1], px[2], tab->sdiv, tab->hdiv, c0, c1, c2);
        else if (space == 2) bgr2lab_px(px[0], px[1], px[2], tab->gamma, tab->cbrt_tab, c0, c1, c2);
        else { c0 = px[0]; c1 = px[1]; c2 = px[2]; }
        uint8_t* o = out + (size_t)r * out_stride + 3 * (size_t)c;
        o[0] = (uint8_t)c0; o[1] = (uint8_t)c1; o[2] = (uint8_t)c2;
    }
}

// X[row, q*3K + 3k + ch] = plane value at (r + dr_k, c + dc_k) or -1 outside the crop (:263-272)
__global__ void __launch_bounds__(256) gather_kernel(const uint8_t* __restrict__ frame, long long stride, int cx, int cy,
                                                     int cw, int ch, Geom g, const ColorTables* __restrict__ tab,
                                                     int16_t* __restrict__ X) {
    const long long n = (long long)ch * cw * g.K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % g.K);
        const long long pix = i / g.K;
        const int r = (int)(pix / cw), c = (int)(pix - (long long)r * cw);
        int dr, dc;
        star_tap(k, dr, dc);
        const int rr = r + dr, cc = c + dc;
        const bool ok = rr >= 0 && rr < ch && cc >= 0 && cc < cw;
        int b = 0, gg = 0, red = 0;
        if (ok) {
            const uint8_t* px = frame + (size_t)(cy + rr) * stride + 3 * (size_t)(cx + cc);
            b = px[0]; gg = px[1]; red = px[2];
        }
        for (int q = 0; q < g.n_spaces; ++q) {
            int c0 = -1, c1 = -1, c2 = -1;
            if (ok) {
                const int sid = g.space_id[q];
                if (sid == 1) bgr2hsv_px(b, gg, red, tab->sdiv, tab->hdiv, c0, c1, c2);
                else if (sid == 2) bgr2lab_px(b, gg, red, tab->gamma, tab->cbrt_tab, c0, c1, c2);
                else { c0 = b; c1 = gg; c2 = red; }
            }
            int16_t* o = X + (size_t)pix * g.F + (size_t)q * 3 * g.K + 3 * k;
            o[0] = (int16_t)c0; o[1] = (int16_t)c1; o[2] = (int16_t)c2;
        }
    }
}

}  // namespace pcm
