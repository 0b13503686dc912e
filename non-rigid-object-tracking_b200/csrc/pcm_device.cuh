// pcm_device.cuh -- device-side building blocks of the PC masker hot path (sm_100a).
//
// Colour conversions follow OpenCV's 8-bit fixed-point algorithms, which the
// reference reaches through cv.cvtColor (maskers/pixel_classification.py:305,307;
// main.py:285).  They are integer-exact; tests/test_gpu_parity.py checks them over
// all 2^24 colours.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pcm {

// ---- tile geometry of the fused score kernel ---------------------------------
// A tile is TILE_W columns x 4 * PPT rows: 8 warps = 4 row groups x 2 column halves, a thread walks PPT
// vertically adjacent pixels of one column.  PPT is a compile-time parameter of the kernel (6, 7 or 8);
// the host picks it per launch (pcm_api.cu, enqueue_update): 8 whenever there are at least as many
// tiles as CTA slots, fewer rows for small crops where every CTA has a single tile.
constexpr int TILE_W = 64;            // output pixels per tile row
#ifndef PCM_NTHREADS
#define PCM_NTHREADS 256
#endif
#ifndef PCM_MIN_CTAS
#define PCM_MIN_CTAS 2
#endif
constexpr int NTHREADS = PCM_NTHREADS;   // warp = (row group, 32-column half of the tile)
constexpr int ROW_GROUPS = NTHREADS / 64;
constexpr int MIN_PPT = 6, MAX_PPT = 8;
constexpr int MAX_TILE_H = ROW_GROUPS * MAX_PPT;
constexpr int MAX_NEIGHBORS = 16;
constexpr int MAX_SPACES = 3;
constexpr int LAB_CBRT_SIZE = 2041;
constexpr int LAB_CBRT_PAD = 2048;

static_assert(TILE_W == 64 && NTHREADS % 64 == 0, "thread mapping: NTHREADS/64 row groups x 2 column halves");

// Lookup tables built on the host at pcm_create (pcm_api.cu: build_tables).
struct ColorTables {
    uint16_t gamma[256];              // rint(2040 * srgb^-1(i/255))
    uint16_t cbrt_tab[LAB_CBRT_PAD];  // rint(32768 * f(i/2040)), i <= 2040
    int32_t sdiv[256];                // rint(255*4096 / i)
    int32_t hdiv[256];                // rint(180*4096 / (6 i))
};

// Geometry shared by host (tap-offset encoding) and device.
struct Geom {
    int n;            // n_neighbors
    int n_spaces;     // Q
    int space_id[MAX_SPACES];
    int K;            // 1 + 8 n taps
    int F;            // 3 K Q features
    int n_planes;     // 3 Q + 1 (last plane = in-crop validity)
    int HX;           // horizontal halo of the smem tile in samples: the TMA box must start on a
                      // 16-byte boundary of the row (measured: any other x faults on sm_100a) -> 16
    int RS;           // plane row stride in the smem tile (bytes) = TMA box width = TILE_W + 2 HX
    int PS;           // plane stride in the smem tile (bytes): RS * (MAX_TILE_H + 2 n) rounded up to 128 -- the
                      // SAME for every tile height, so that the tap offsets packed into the forests do not
                      // depend on the tile height chosen per launch (each plane of a tile is its own TMA box)
};

__host__ __device__ inline void star_tap(int k, int& dr, int& dc) {
    // order of pixel_classification.py:254-258
    if (k == 0) { dr = 0; dc = 0; return; }
    int i = (k - 1) / 8 + 1;
    switch ((k - 1) % 8) {
        case 0: dr = -i; dc = 0; break;
        case 1: dr = +i; dc = 0; break;
        case 2: dr = 0; dc = -i; break;
        case 3: dr = 0; dc = +i; break;
        case 4: dr = +i; dc = +i; break;
        case 5: dr = -i; dc = -i; break;
        case 6: dr = +i; dc = -i; break;
        default: dr = -i; dc = +i; break;
    }
}

// Programmatic dependent launch: returns once the preceding kernel of the stream has completed and
// its writes are visible (a no-op for a kernel launched without the attribute).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the NEXT kernel of the stream start once every CTA of this grid has got here: its CTAs take
// the slots this grid frees, run whatever precedes their own grid_dependency_wait(), and wait.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- BGR -> HSV (OpenCV RGB2HSV_b, hrange 180) ----------------------------------
__device__ __forceinline__ void bgr2hsv_px(int b, int g, int r, const int32_t* __restrict__ sdiv,
                                           const int32_t* __restrict__ hdiv, int& H, int& S, int& V) {
    int v = max(max(b, g), r);
    int vmin = min(min(b, g), r);
    int diff = v - vmin;
    int h = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff));
    S = (diff * sdiv[v] + (1 << 11)) >> 12;
    h = (h * hdiv[diff] + (1 << 11)) >> 12;   // arithmetic shift == floor, as OpenCV
    H = h + ((h < 0) ? 180 : 0);
    V = v;
}

// ---- BGR -> Lab (OpenCV RGB2Lab_b, sRGB, D65) -------------------------------------
__device__ __forceinline__ void bgr2lab_px(int b, int g, int r, const uint16_t* __restrict__ gamma,
                                           const uint16_t* __restrict__ cbrt_tab, int& L, int& A, int& Bc) {
    int R = gamma[r], G = gamma[g], B = gamma[b];
    int fX = cbrt_tab[(R * 1777 + G * 1541 + B * 778 + 2048) >> 12];
    int fY = cbrt_tab[(R * 871 + G * 2929 + B * 296 + 2048) >> 12];
    int fZ = cbrt_tab[(R * 73 + G * 448 + B * 3575 + 2048) >> 12];
    int l = (296 * fY - 1336934 + 16384) >> 15;
    int a = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
    int bb = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
    L = min(max(l, 0), 255);
    A = min(max(a, 0), 255);
    Bc = min(max(bb, 0), 255);
}

// ---- BGR -> gray (OpenCV RGB2Gray 8-bit) -------------------------------------------
__device__ __forceinline__ int bgr2gray_px(int b, int g, int r) {
    return (3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15;
}

// exact u8 -> double without the slow I2F unit: (2^52 + v) - 2^52
__device__ __forceinline__ double u8_to_double(unsigned v) {
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

}  // namespace pcm
