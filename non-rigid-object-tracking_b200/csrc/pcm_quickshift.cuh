// pcm_quickshift.cuh -- quickshift over-segmentation of the crop on the GPU ("next" row f-1 of
// the step that feeds labels to the hot path,
// reference maskers/pixel_classification.py:70-71
//     quickshift(crop_frame, kernel_size=3, max_dist=6, ratio=0.5, random_seed=42)).
//
// Algorithm = scikit-image 0.17.2 (environment.yaml:12; source not in the reference tree, so
// parity is pinned only against the restatement in oracle/quickshift_oracle.py):
//   Q1 qs_lab_kernel      u8 crop -> float64 Lab * ratio, planar   (skimage rgb2xyz + xyz2lab)
//   Q2 qs_density_kernel  density = sum over the (2w+1)^2 window of exp(-d2 / (2 ks^2)) + tie noise
//   Q3 qs_parent_kernel   parent = nearest window pixel of higher density; cut links > max_dist
//   Q4 qs_root_kernel     follow parents to the root
//   Q5 qs_flag / scan / qs_label   labels = rank of the root among all roots (np.unique inverse)
// All sums run in the window's raster order with separate multiplies and adds (no FMA
// contraction), exactly as the Cython loops do, so equal inputs give equal float64 densities
// up to the last-ulp differences of exp() between math libraries.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

namespace pcm {

struct QsArgs {
    const uint8_t* frame;      // BGR interleaved
    long long stride;
    int cx, cy, cw, ch;
    const double* lin;         // [256] sRGB -> linear
    double ratio;
    double* lab;               // [3][ch][cw]
    double* dens;              // [ch][cw]
    const double* noise;       // [ch][cw] or nullptr
    int* parent;               // [ch*cw]
    int* root;                 // [ch*cw]
    int kw;                    // window half width = ceil(3 * kernel_size)
    double inv;                // -0.5 / kernel_size^2
    double max_dist;
    const double* exp_tab;     // [64][2]: 2^(j/64) as a high and a low part (qs_exp_neg)
    int exp_guard;             // 1: exp arguments may fall below -700 (see qs_exp_neg)
    int pw;                    // half width of the PARENT search window: min(kw, floor(max_dist)), see qs_parent_kernel
};

// exp(x) for x <= 0, table driven: x = (64 m + j) ln2/64 + r, |r| <= ln2/128,
//   exp(x) = 2^m * 2^(j/64) * (1 + r + r^2/2 + r^3/6 + r^4/24 + r^5/120)        (truncation 3.5e-17 relative)
// 2^(j/64) comes from a 64-entry table held as high + low parts (shared memory, `tab`).  About 1 ulp, like the CUDA math
// library's exp() -- which needs about twice the float64 operations, and this kernel is bound by the float64 pipe
// (profiles/README.md).  GUARD: arguments below -700 are clamped to -700 (the exponent arithmetic below assumes a normal
// result).  That changes no density: a window always contains its own centre, whose term is exp(0) = 1, so a term of
// 1e-304 or less -- exact or clamped -- is absorbed without trace by the first normal-sized term added after it.  The host
// drops the guard when the parameters cannot produce such an argument (true for the reference's ratio 0.5, kernel 3).
template <bool GUARD>
__device__ __forceinline__ double qs_exp_neg(double x, const double* __restrict__ tab) {
    if (GUARD && (unsigned)__double2hiint(x) > 0xC085E000u) x = -700.0;      // x < -700 (integer compare: keeps the float64 pipe free)
    const double t = fma(x, 92.33248261689366, 6755399441055744.0);          // 64 / ln2; 1.5 * 2^52 rounds to an integer
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -0x1.62e42fefa0000p-7, x);                             // ln2 / 64, high 36 bits: exact product
    r = fma(kd, -0x1.cf79abc9e3b3ap-46, r);
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    q = fma(r * r, q, r);                                                      // exp(r) - 1
    const int j = k & 63;
    const double hi = tab[2 * j], lo = tab[2 * j + 1];
    const double v = hi + fma(hi, q, lo);
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}

__global__ void __launch_bounds__(256) qs_lab_kernel(const QsArgs a) {
    const int n = a.cw * a.ch;
    const size_t ps = (size_t)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / a.cw, c = i - r * a.cw;
        const uint8_t* px = a.frame + (long long)(a.cy + r) * a.stride + (long long)(a.cx + c) * 3;
        // the reference hands an OpenCV BGR crop to a function that expects RGB: channel 0 plays 'R'
        const double c0 = a.lin[px[0]], c1 = a.lin[px[1]], c2 = a.lin[px[2]];
        const double M[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};
        const double white[3] = {0.95047, 1.0, 1.08883};
        double f[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double x = __dadd_rn(__dadd_rn(__dmul_rn(c0, M[k][0]), __dmul_rn(c1, M[k][1])), __dmul_rn(c2, M[k][2]));
            const double t = __ddiv_rn(x, white[k]);
            f[k] = t > 0.008856 ? cbrt(t) : __dadd_rn(__dmul_rn(7.787, t), 16.0 / 116.0);
        }
        const double L = __dsub_rn(__dmul_rn(116.0, f[1]), 16.0);
        const double A = __dmul_rn(500.0, __dsub_rn(f[0], f[1]));
        const double B = __dmul_rn(200.0, __dsub_rn(f[1], f[2]));
        a.lab[i] = __dmul_rn(L, a.ratio);
        a.lab[ps + i] = __dmul_rn(A, a.ratio);
        a.lab[2 * ps + i] = __dmul_rn(B, a.ratio);
    }
}

// squared 5-D distance in the Cython loop's order: channels, then rows, then columns
__device__ __forceinline__ double qs_dist(double l0, double a0, double b0, double l1, double a1, double b1, int dr, int dc) {
    double t = __dsub_rn(l0, l1);
    double d = __dmul_rn(t, t);
    t = __dsub_rn(a0, a1);
    d = __dadd_rn(d, __dmul_rn(t, t));
    t = __dsub_rn(b0, b1);
    d = __dadd_rn(d, __dmul_rn(t, t));
    t = (double)dr;
    d = __dadd_rn(d, __dmul_rn(t, t));
    t = (double)dc;
    d = __dadd_rn(d, __dmul_rn(t, t));
    return d;
}

// 32 x 8 pixels per block; the Lab values of the block's window (tile + halo) are staged in
// shared memory once and reused by the (2w+1)^2 taps of every pixel.
constexpr int QS_BW = 32, QS_BH = 8, QS_MAX_KW = 15;

// PARENT pass window: the reference takes the nearest higher-density pixel of the whole (2 kw + 1)^2 window and then
// cuts the link if it is longer than max_dist (:parent_flat[dist_parent_flat > max_dist] = self).  A pixel farther than
// max_dist in the image plane alone can therefore never end up as the parent, and if no candidate within that radius is
// close enough the answer is "self" whatever lies outside: searching |dr|, |dc| <= pw = min(kw, floor(max_dist)) in the
// same raster order gives the identical result with (2 pw + 1)^2 instead of (2 kw + 1)^2 taps (169 instead of 361 for
// the reference's parameters).
__global__ void __launch_bounds__(QS_BW * QS_BH) qs_parent_kernel(const QsArgs a) {
    extern __shared__ double qs_smem[];
    const int kw = a.pw, TW = QS_BW + 2 * kw, TH = QS_BH + 2 * kw;
    double* sL = qs_smem;
    double* sA = sL + TW * TH;
    double* sB = sA + TW * TH;
    double* sD = sB + TW * TH;            // densities
    const int x0 = blockIdx.x * QS_BW, y0 = blockIdx.y * QS_BH;
    const size_t ps = (size_t)a.cw * a.ch;
    for (int i = threadIdx.x; i < TW * TH; i += QS_BW * QS_BH) {
        const int tr = i / TW, tc = i - tr * TW;
        const int y = y0 - kw + tr, x = x0 - kw + tc;
        const bool in = y >= 0 && y < a.ch && x >= 0 && x < a.cw;
        const size_t o = in ? (size_t)y * a.cw + x : 0;
        sL[i] = in ? a.lab[o] : 0.0;
        sA[i] = in ? a.lab[ps + o] : 0.0;
        sB[i] = in ? a.lab[2 * ps + o] : 0.0;
        sD[i] = in ? a.dens[o] : 0.0;
    }
    __syncthreads();
    const int lx = threadIdx.x % QS_BW, ly = threadIdx.x / QS_BW;
    const int c = x0 + lx, r = y0 + ly;
    if (c >= a.cw || r >= a.ch) return;
    const int r_min = max(r - kw, 0), r_max = min(r + kw + 1, a.ch);
    const int c_min = max(c - kw, 0), c_max = min(c + kw + 1, a.cw);
    const int me = (ly + kw) * TW + (lx + kw);
    const double l0 = sL[me], a0 = sA[me], b0 = sB[me];
    const double cur = sD[me];
    double closest = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    int closest_ceil = 0x7fffffff;         // ceil(closest): an integer s is < closest exactly when s < closest_ceil
    int best = r * a.cw + c;
    for (int r_ = r_min; r_ < r_max; ++r_) {
        const int row = (r_ - y0 + kw) * TW + kw - x0;
        const int dr2 = (r - r_) * (r - r_);
        for (int c_ = c_min; c_ < c_max; ++c_) {
            const int j = row + c_;
            // the squared distance is at least its (exactly representable) image-plane part, whatever the roundings of the
            // colour terms: a tap that far away cannot beat the current candidate and its Lab values are not even fetched
            if (sD[j] > cur && dr2 + (c - c_) * (c - c_) < closest_ceil) {
                const double d = qs_dist(l0, a0, b0, sL[j], sA[j], sB[j], r - r_, c - c_);
                if (d < closest) {
                    closest = d; best = r_ * a.cw + c_;
                    closest_ceil = d < 1073741824.0 ? (int)ceil(d) : 0x7fffffff;
                }
            }
        }
    }
    // parent_flat[dist_parent_flat > max_dist] = self, dist_parent = sqrt(closest)
    if (sqrt(closest) > a.max_dist) best = r * a.cw + c;
    a.parent[r * a.cw + c] = best;
}

// Density pass.  A thread owns QS_DR vertically adjacent pixels and fetches the Lab values of a window position from
// shared memory once for all of them; every pixel still adds its window in raster order (rows ascending, columns
// ascending inside a row), exactly like the Cython loop.  Block = 32 x 8 threads = 32 x (8 QS_DR) pixels.
// Measured at 1080p (round 2, profiles/README.md): QS_DR = 1 / 2 / 3 / 4 / 6 / 8 -> 1.43 / 1.47 / 1.53 / 1.57 / 1.65 /
// 2.68 ms.  Sharing the fetches (0.9 instead of 3 shared-memory loads per pixel-tap at QS_DR = 4) does NOT pay: the kernel
// waits on float64 dependency chains (top stall `wait`, float64 pipe 61 %), and one pixel per thread keeps more warps
// resident (31 KB of shared memory per block instead of 61 KB).  Default: 1.
#ifndef PCM_QS_UNROLL
#define PCM_QS_UNROLL 2             // taps of a window row in flight per thread (independent exp chains): 1 / 2 / 4 -> 1.43 / 1.38 / 1.38 ms
#endif
#ifndef PCM_QS_DR
#define PCM_QS_DR 1                 // rows per thread (see above)
#endif
constexpr int QS_DR = PCM_QS_DR, QS_DBW = 32, QS_DBH = 8 * QS_DR, QS_UNROLL = PCM_QS_UNROLL;

template <bool GUARD>
__global__ void __launch_bounds__(256) qs_density_kernel(const QsArgs a) {
    extern __shared__ double qs_smem[];
    const int kw = a.kw, TW = QS_DBW + 2 * kw, TH = QS_DBH + 2 * kw;
    double* sL = qs_smem;
    double* sA = sL + TW * TH;
    double* sB = sA + TW * TH;
    double* tab = sB + TW * TH;            // [128] exp table
    double* sq = tab + 128;                // [2 kw + 1] squares of the column offsets
    const int x0 = blockIdx.x * QS_DBW, y0 = blockIdx.y * QS_DBH;
    const size_t ps = (size_t)a.cw * a.ch;
    for (int i = threadIdx.x; i < TW * TH; i += 256) {
        const int tr = i / TW, tc = i - tr * TW;
        const int y = y0 - kw + tr, x = x0 - kw + tc;
        const bool in = y >= 0 && y < a.ch && x >= 0 && x < a.cw;
        const size_t o = in ? (size_t)y * a.cw + x : 0;
        sL[i] = in ? a.lab[o] : 0.0;
        sA[i] = in ? a.lab[ps + o] : 0.0;
        sB[i] = in ? a.lab[2 * ps + o] : 0.0;
    }
    for (int i = threadIdx.x; i < 128; i += 256) tab[i] = a.exp_tab[i];
    for (int i = threadIdx.x; i < 2 * kw + 1; i += 256) { const double t = (double)(i - kw); sq[i] = __dmul_rn(t, t); }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int c = x0 + lx, rb = y0 + ly * QS_DR;            // the thread's pixels: (rb + p, c), p < QS_DR
    if (c >= a.cw || rb >= a.ch) return;
    const int c_min = max(c - kw, 0), c_max = min(c + kw + 1, a.cw);
    double l0[QS_DR], a0[QS_DR], b0[QS_DR], acc[QS_DR];
#pragma unroll
    for (int p = 0; p < QS_DR; ++p) {
        const int me = (ly * QS_DR + p + kw) * TW + (lx + kw);
        l0[p] = sL[me]; a0[p] = sA[me]; b0[p] = sB[me]; acc[p] = 0.0;       // rows below the crop read the zero padding
    }
    const double* sqc = sq + kw + c;                                               // sqc[-c_] = (c - c_)^2
    const double inv = a.inv;
    // one window row for the thread's pixels; ALL: the row lies in the window of every one of them (no tests in the loop)
    auto window_row = [&](int r_, auto all) {
        constexpr bool ALL = decltype(all)::value;
        const int row = (r_ - y0 + kw) * TW + kw - x0;
        double dr2[QS_DR];
        bool on[QS_DR];
#pragma unroll
        for (int p = 0; p < QS_DR; ++p) {
            const int dr = rb + p - r_;
            on[p] = ALL || (dr >= -kw && dr <= kw);
            const double t = (double)dr;
            dr2[p] = __dmul_rn(t, t);
        }
#pragma unroll QS_UNROLL
        for (int c_ = c_min; c_ < c_max; ++c_) {
            const int j = row + c_;
            const double L = sL[j], A = sA[j], B = sB[j], dc2 = sqc[-c_];
#pragma unroll
            for (int p = 0; p < QS_DR; ++p) {
                // the Cython loop's order: channels, then rows, then columns; separate multiplies and adds
                double t = __dsub_rn(l0[p], L);
                double d = __dmul_rn(t, t);
                t = __dsub_rn(a0[p], A);
                d = __dadd_rn(d, __dmul_rn(t, t));
                t = __dsub_rn(b0[p], B);
                d = __dadd_rn(d, __dmul_rn(t, t));
                d = __dadd_rn(d, dr2[p]);
                d = __dadd_rn(d, dc2);
                const double e = qs_exp_neg<GUARD>(__dmul_rn(d, inv), tab);
                if (ALL) acc[p] = __dadd_rn(acc[p], e);
                else if (on[p]) acc[p] = __dadd_rn(acc[p], e);       // (a select: rows outside this pixel's window add nothing)
            }
        }
    };
    // rows of the union of the windows, clipped to the crop; in [all_lo, all_hi) every pixel of the thread takes part
    const int r_lo = max(rb - kw, 0), r_hi = min(rb + QS_DR - 1 + kw + 1, a.ch);
    const int all_lo = min(max(rb + QS_DR - 1 - kw, r_lo), r_hi), all_hi = max(min(rb + kw + 1, r_hi), all_lo);
    for (int r_ = r_lo; r_ < all_lo; ++r_) window_row(r_, std::false_type{});
    for (int r_ = all_lo; r_ < all_hi; ++r_) window_row(r_, std::true_type{});
    for (int r_ = all_hi; r_ < r_hi; ++r_) window_row(r_, std::false_type{});
#pragma unroll
    for (int p = 0; p < QS_DR; ++p) {
        if (rb + p < a.ch) {
            const size_t o = (size_t)(rb + p) * a.cw + c;
            a.dens[o] = a.noise ? __dadd_rn(acc[p], a.noise[o]) : acc[p];
        }
    }
}

__global__ void __launch_bounds__(256) qs_root_kernel(const int* __restrict__ parent, int* __restrict__ root,
                                                      int* __restrict__ flag, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int p = i, q = __ldg(parent + i);
        while (q != p) { p = q; q = __ldg(parent + p); }      // density strictly increases along links: no cycles
        root[i] = p;
        if (p == i) flag[i] = 1;
    }
}

// ---- exclusive scan of the root flags (np.unique ranks), 1024 elements per block ------------
constexpr int QS_SCAN_ELEMS = 1024;

__device__ __forceinline__ int qs_block_exclusive(int v, int& total) {   // 256 threads
    __shared__ int warp_sums[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) base += warp_sums[w]; tot += warp_sums[w]; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(256) qs_scan_reduce_kernel(const int* __restrict__ flag, int n, int* __restrict__ block_sums) {
    const int base = blockIdx.x * QS_SCAN_ELEMS + threadIdx.x * 4;
    int s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += (base + j < n) ? flag[base + j] : 0;
    int total;
    qs_block_exclusive(s, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) qs_scan_sums_kernel(int* __restrict__ block_sums, int n_blocks, int* __restrict__ n_labels) {
    int carry = 0;                       // single block; chunks of 256 block sums
    for (int b0 = 0; b0 < n_blocks; b0 += 256) {
        const int i = b0 + threadIdx.x;
        const int v = i < n_blocks ? block_sums[i] : 0;
        int total;
        const int ex = qs_block_exclusive(v, total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *n_labels = carry;
}

__global__ void __launch_bounds__(256) qs_scan_apply_kernel(const int* __restrict__ flag, int n, const int* __restrict__ block_offsets,
                                                            int* __restrict__ rank) {
    const int base = blockIdx.x * QS_SCAN_ELEMS + threadIdx.x * 4;
    int f[4], s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[j] = (base + j < n) ? flag[base + j] : 0; s += f[j]; }
    int total;
    int ex = qs_block_exclusive(s, total) + block_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (base + j < n) rank[base + j] = ex;
        ex += f[j];
    }
}

__global__ void __launch_bounds__(256) qs_label_kernel(const int* __restrict__ root, const int* __restrict__ rank,
                                                       int32_t* __restrict__ labels, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) labels[i] = rank[root[i]];
}

}  // namespace pcm
