// pcm_prior.cuh -- the SIFT-match prior of the PC masker (reference maskers/pixel_classification.py:129-163,
// computePriors) on the GPU (sm_100a), SURVEY.md §8 row f-2.
//
// The reference detects SIFT keypoints on the previous crop (inside the previous foreground mask) and on the
// current crop, matches them with FLANN (kd-trees, Lowe ratio 0.7), drops matches whose displacement exceeds the
// 90th percentile and gives prior +1 to the superpixel under every surviving keypoint of the current crop.
// Keypoint DETECTION stays with OpenCV on the host (once per clip frame); everything after it runs here per frame:
//   P1 prior_match_kernel   one warp per previous-crop keypoint: mask filter (KeyPointsFilter::runByPixelsMask:
//                           mask[(int)(y + 0.5f)][(int)(x + 0.5f)] != 0), EXACT 2-nearest-neighbour search over the
//                           current crop's descriptors (integer squared L2 over the 128 uint8 entries, __vabsdiffu4 +
//                           __dp4a), ratio test on the float32 square roots as FLANN reports them, displacement (f64)
//   P2 prior_finish_kernel  one block: priors = -1; np.percentile(dist, 90) by rank counting (numpy's virtual index
//                           and two-sided lerp); +1 at labels[int(y)][int(x)] of the kept matches
// FLANN's randomised kd-trees are an APPROXIMATE nearest-neighbour search whose result differs from run to run; the
// exact search here is what it approximates (oracle/prior_oracle.py restates this path; tests compare bit for bit).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pcm {

struct PriorArgs {
    const float* pts1;        // [m1][2] keypoints (x, y) of the previous crop, unmasked detection
    const uint8_t* des1;      // [m1][128]
    int m1;
    const uint8_t* prev_mask; // element (0, 0) = top-left pixel of the previous crop inside the mask plane
    long long prev_stride;
    int prev_w, prev_h;
    const float* pts2;        // [m2][2] keypoints of the current crop
    const uint8_t* des2;      // [m2][128]
    int m2;
    const int32_t* labels;    // [ch][cw] over-segmentation of the current crop
    int cw, ch, n_labels;
    float* priors;            // [n_labels] out
    // scratch
    double* q_dist;           // [m1] displacement of query i's match (valid when q_j >= 0)
    int* q_j;                 // [m1] matched current keypoint, or -1
    double* g_dist;           // [m1] compacted
    int* g_j;                 // [m1]
};

__device__ __forceinline__ void top2_insert(unsigned d, int j, unsigned& d1, int& j1, unsigned& d2, int& j2) {
    if (d < d1 || (d == d1 && j < j1)) { d2 = d1; j2 = j1; d1 = d; j1 = j; }
    else if (d < d2 || (d == d2 && j < j2)) { d2 = d; j2 = j; }
}

__global__ void __launch_bounds__(256) prior_match_kernel(const PriorArgs a) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= a.m1) return;
    const float px = a.pts1[2 * i], py = a.pts1[2 * i + 1];
    const int mx = (int)(px + 0.5f), my = (int)(py + 0.5f);
    bool ok = a.m2 >= 2 && mx >= 0 && my >= 0 && mx < a.prev_w && my < a.prev_h &&
              a.prev_mask[(long long)my * a.prev_stride + mx] != 0;
    if (!ok) { if (lane == 0) a.q_j[i] = -1; return; }
    uint32_t q[32];
    const uint4* qp = reinterpret_cast<const uint4*>(a.des1 + (size_t)i * 128);
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint4 v = __ldg(qp + w); q[4 * w] = v.x; q[4 * w + 1] = v.y; q[4 * w + 2] = v.z; q[4 * w + 3] = v.w; }
    unsigned d1 = 0xffffffffu, d2 = 0xffffffffu;
    int j1 = 0x7fffffff, j2 = 0x7fffffff;
    for (int j = lane; j < a.m2; j += 32) {
        const uint4* cp = reinterpret_cast<const uint4*>(a.des2 + (size_t)j * 128);
        unsigned acc = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint4 v = __ldg(cp + w);
            unsigned t;
            t = __vabsdiffu4(q[4 * w], v.x); acc = __dp4a(t, t, acc);
            t = __vabsdiffu4(q[4 * w + 1], v.y); acc = __dp4a(t, t, acc);
            t = __vabsdiffu4(q[4 * w + 2], v.z); acc = __dp4a(t, t, acc);
            t = __vabsdiffu4(q[4 * w + 3], v.w); acc = __dp4a(t, t, acc);
        }
        top2_insert(acc, j, d1, j1, d2, j2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned e1 = __shfl_xor_sync(0xffffffffu, d1, o), e2 = __shfl_xor_sync(0xffffffffu, d2, o);
        const int k1 = __shfl_xor_sync(0xffffffffu, j1, o), k2 = __shfl_xor_sync(0xffffffffu, j2, o);
        top2_insert(e1, k1, d1, j1, d2, j2);
        top2_insert(e2, k2, d1, j1, d2, j2);
    }
    if (lane == 0) {
        // FLANN reports sqrt of the float32 squared distance; the reference compares them as Python floats (:147)
        const double s1 = (double)__fsqrt_rn((float)d1), s2 = (double)__fsqrt_rn((float)d2);
        if (s1 < __dmul_rn(0.7, s2)) {
            const double dx = __dsub_rn((double)a.pts2[2 * j1], (double)px), dy = __dsub_rn((double)a.pts2[2 * j1 + 1], (double)py);
            a.q_dist[i] = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            a.q_j[i] = j1;
        } else {
            a.q_j[i] = -1;
        }
    }
}

__global__ void __launch_bounds__(1024) prior_finish_kernel(const PriorArgs a) {
    __shared__ int s_G;
    __shared__ double s_lo, s_hi;
    const int tid = threadIdx.x;
    for (int s = tid; s < a.n_labels; s += blockDim.x) a.priors[s] = -1.f;
    if (tid == 0) s_G = 0;
    __syncthreads();
    for (int i = tid; i < a.m1; i += blockDim.x) {
        const int j = a.q_j[i];
        if (j >= 0) { const int g = atomicAdd(&s_G, 1); a.g_dist[g] = a.q_dist[i]; a.g_j[g] = j; }
    }
    __syncthreads();
    const int G = s_G;
    if (G == 0) return;
    // np.percentile(dist, 90), method "linear": virtual index n q + (alpha + q (1 - alpha - beta)) - 1, alpha = beta = 1
    const double q = 90.0 / 100.0;
    const double vi = __dsub_rn(__dadd_rn(__dmul_rn((double)G, q), __dadd_rn(1.0, __dmul_rn(q, -1.0))), 1.0);
    int lo = (int)floor(vi), hi = lo + 1;
    double gamma = __dsub_rn(vi, floor(vi));
    if (vi >= (double)(G - 1)) { lo = hi = G - 1; }
    if (vi < 0.0) { lo = hi = 0; }
    for (int g = tid; g < G; g += blockDim.x) {
        const double d = a.g_dist[g];
        int rank = 0;
        for (int k = 0; k < G; ++k) {
            const double e = a.g_dist[k];
            rank += (e < d) || (e == d && k < g);
        }
        if (rank == lo) s_lo = d;
        if (rank == hi) s_hi = d;
    }
    __syncthreads();
    double thr;
    {
        const double lo_v = s_lo, hi_v = s_hi, diff = __dsub_rn(hi_v, lo_v);
        thr = (lo == hi) ? lo_v
              : (gamma >= 0.5 ? __dsub_rn(hi_v, __dmul_rn(diff, __dsub_rn(1.0, gamma))) : __dadd_rn(lo_v, __dmul_rn(diff, gamma)));
    }
    for (int g = tid; g < G; g += blockDim.x) {
        if (a.g_dist[g] <= thr) {
            const int j = a.g_j[g];
            const int x = (int)a.pts2[2 * j], y = (int)a.pts2[2 * j + 1];
            if (x >= 0 && y >= 0 && x < a.cw && y < a.ch) {
                const int l = a.labels[(size_t)y * a.cw + x];
                if (l >= 0 && l < a.n_labels) a.priors[l] = 1.f;
            }
        }
    }
}

}  // namespace pcm
