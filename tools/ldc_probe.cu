// ldc_probe.cu -- does a DIVERGENT indexed constant-bank load (LDC.64, <= 4 distinct addresses per
// warp) ride for free next to a kernel that saturates the shared-memory/LSU pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ldc_probe tools/ldc_probe.cu && tools/ldc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Table { uint2 n[48][8]; };

template <int MODE>   // 0: 2 LDS.64 / iter; 1: + LDC.64 divergent(4); 2: + third LDS.64 instead; 3: + LDC.64 uniform
__global__ void __launch_bounds__(256, 2) probe(const __grid_constant__ Table tab, uint32_t* out, int iters) {
    __shared__ uint2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_uint2(i * 8 + 8, (i * 37) & 2047);
    __syncthreads();
    uint32_t a = threadIdx.x & 2047, b = (threadIdx.x * 7) & 2047, c = (threadIdx.x * 13) & 2047, acc = 0;
    uint32_t sel = threadIdx.x & 3;
    for (int i = 0; i < iters; ++i) {
        const uint2 x = sm[a], y = sm[b];
        a = x.y; b = y.y;
        if (MODE == 1) { const uint2 z = tab.n[i % 48][4 + sel]; acc += z.x; sel = (sel + z.y) & 3; }
        if (MODE == 2) { const uint2 z = sm[c]; c = z.y; acc += z.x; }
        if (MODE == 3) { const uint2 z = tab.n[i % 48][4 + (i & 3)]; acc += z.x + z.y; }
        acc += x.x ^ y.x;
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc + a + b + c + sel;
}

int main() {
    Table t;
    for (int i = 0; i < 48; ++i) for (int j = 0; j < 8; ++j) t.n[i][j] = make_uint2(i * 8 + j, (i + j) & 3);
    uint32_t* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<296, 256>>>(t, out, iters);
            if (mode == 1) probe<1><<<296, 256>>>(t, out, iters);
            if (mode == 2) probe<2><<<296, 256>>>(t, out, iters);
            if (mode == 3) probe<3><<<296, 256>>>(t, out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        // warp-iterations per SM per clock: 16 warps/SM * iters / (ms * clock)
        printf("mode %d: %.3f ms  (%.2f SM-cycles per warp-iteration at 1.965 GHz)\n", mode, best,
               best * 1e-3 * 1.965e9 / (16.0 * iters));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
