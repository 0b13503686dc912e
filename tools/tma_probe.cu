// tma_probe.cu -- isolates the TMA tile load used by score_kernel (debug aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "../non-rigid-object-tracking_b200/csrc/pcm_kernels.cuh"
using namespace pcm;

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t dst = smem_u32(smem);
    const uint32_t bar = smem_u32(smem + 65536);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, bytes);
        tma_load_3d(dst, &tmap, x, y, 0, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
    const int cw = 224, ch = 139, np = 7, RS = 80, PH = 48;
    const long long pitch = 256, ps = pitch * ch;
    std::vector<uint8_t> h((size_t)ps * np);
    for (int p = 0; p < np; ++p) for (int r = 0; r < ch; ++r) for (int c = 0; c < pitch; ++c)
        h[p * ps + r * pitch + c] = (uint8_t)(1 + (p * 31 + r * 7 + c) % 250);
    uint8_t *d, *out;
    CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    const int bytes = RS * PH * np;
    CK(cudaMalloc(&out, bytes));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)cw, (cuuint64_t)ch, (cuuint64_t)np};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)ps};
    cuuint32_t box[3] = {RS, PH, np};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64));
    for (int t = 0; t < 3; ++t) {
        const int x = t == 0 ? -8 : (t == 1 ? 56 : 184), y = t == 0 ? -8 : (t == 1 ? 24 : 120);
        probe<<<1, 256, 65536 + 64>>>(tm, x, y, bytes, out);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> o(bytes);
        CK(cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (int p = 0; p < np; ++p) for (int rr = 0; rr < PH; ++rr) for (int c = 0; c < RS; ++c) {
            const int gy = y + rr, gx = x + c;
            uint8_t want = (gy >= 0 && gy < ch && gx >= 0 && gx < cw) ? h[p * ps + gy * pitch + gx] : 0;
            bad += o[(p * PH + rr) * RS + c] != want;
        }
        printf("box at (%d,%d): %ld mismatches\n", x, y, bad);
    }
    return 0;
}
