// ldc_probe.cu -- throughput of indexed constant-bank loads (LDC.64) vs shared loads
// (LDS.64) for the node fetch of the forest traversal.  Debug/measurement aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ uint2 c_nodes[4096];

template <int MODE>   // 0 = LDC, 1 = LDS
__global__ void chase(const uint2* g_nodes, int iters, int spread, unsigned* out) {
    __shared__ uint2 s_nodes[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_nodes[i] = g_nodes[i];
    __syncthreads();
    // 8 independent chains per thread, like the traversal; `spread` distinct start nodes per warp
    unsigned ref[8], acc = 0;
    for (int g = 0; g < 8; ++g) ref[g] = ((threadIdx.x & 31) % spread) * 64 + g;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            uint2 nd = MODE == 0 ? c_nodes[ref[g]] : s_nodes[ref[g]];
            acc += nd.y;
            ref[g] = nd.x;
        }
    }
    if (acc == 0xdeadbeef) out[0] = acc + ref[0];
    out[1 + blockIdx.x * blockDim.x + threadIdx.x] = ref[3] + acc;
}

int main() {
    uint2 h[4096];
    for (int i = 0; i < 4096; ++i) { h[i].x = (i / 64) * 64 + (i * 7 + 3) % 64; h[i].y = i; }
    cudaMemcpyToSymbol(c_nodes, h, sizeof h);
    uint2* g; cudaMalloc(&g, sizeof h); cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
    unsigned* out; cudaMalloc(&out, 4 * (1 + 148 * 8 * 256));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 2000, blocks = 148 * 2, threads = 256;
    for (int mode = 0; mode < 2; ++mode)
        for (int spread : {1, 2, 4, 8, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) chase<0><<<blocks, threads>>>(g, iters, spread, out);
                else chase<1><<<blocks, threads>>>(g, iters, spread, out);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            double loads = (double)blocks * threads / 32 * iters * 8;   // warp-level loads
            double per_sm_per_clk = loads / 148 / (ms * 1e-3 * 1.965e9);
            printf("%s spread %2d: %.3f ms, %.3f warp-loads/clk/SM (%s)\n", mode == 0 ? "LDC.64" : "LDS.64", spread, ms,
                   per_sm_per_clk, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
