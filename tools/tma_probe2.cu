// tma_probe2.cu -- variants of a TMA tile load, one per process (argv[1] = variant).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cuda/ptx>
namespace ptx = cuda::ptx;
using barrier_t = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// variant using libcu++ wrappers (official path)
__global__ void k_lib2d(const __grid_constant__ CUtensorMap tmap, int x, int y, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); ptx::fence_proxy_async(ptx::space_shared); }
    __syncthreads();
    barrier_t::arrival_token tok;
    if (threadIdx.x == 0) {
        cuda::device::experimental::cp_async_bulk_tensor_2d_global_to_shared(smem, &tmap, x, y, bar);
        tok = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else tok = bar.arrive();
    bar.wait(std::move(tok));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
__global__ void k_lib3d(const __grid_constant__ CUtensorMap tmap, int x, int y, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); ptx::fence_proxy_async(ptx::space_shared); }
    __syncthreads();
    barrier_t::arrival_token tok;
    if (threadIdx.x == 0) {
        cuda::device::experimental::cp_async_bulk_tensor_3d_global_to_shared(smem, &tmap, x, y, 0, bar);
        tok = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else tok = bar.arrive();
    bar.wait(std::move(tok));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
// raw PTX variant (what pcm_kernels.cuh does), rank given
template <int RANK>
__global__ void k_raw(const __grid_constant__ CUtensorMap tmap, int x, int y, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long barmem;
    const uint32_t dst = smem_u32(smem), bar = smem_u32(&barmem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(&tmap), "r"(x), "r"(y), "r"(bar) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(&tmap), "r"(x), "r"(y), "r"(0), "r"(bar) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int rank = (variant == 0 || variant == 2) ? 2 : 3;
    const bool lib = variant <= 1;
    int bw = 80, bh = 48, np = 7;
    if (argc > 2) bw = atoi(argv[2]);
    if (argc > 3) bh = atoi(argv[3]);
    if (argc > 4) np = atoi(argv[4]);
    int x = argc > 5 ? atoi(argv[5]) : 0, y = argc > 6 ? atoi(argv[6]) : 0;
    const int cw = 224, ch = 139;
    const long long pitch = 256, ps = pitch * ch;
    std::vector<uint8_t> h((size_t)ps * np);
    for (int p = 0; p < np; ++p) for (int r = 0; r < ch; ++r) for (int c = 0; c < pitch; ++c)
        h[p * ps + r * pitch + c] = (uint8_t)(1 + (p * 31 + r * 7 + c) % 250);
    uint8_t *d, *out;
    CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    const int planes = rank == 3 ? np : 1;
    const int bytes = bw * bh * planes;
    CK(cudaMalloc(&out, bytes));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q));
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)cw, (cuuint64_t)ch, (cuuint64_t)np};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)ps};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)np};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d rank %d box %dx%dx%d at (%d,%d): encode -> %d; ", variant, rank, bw, bh, planes, x, y, (int)r);
    const int smem = 65536;
    if (lib && rank == 2) { CK(cudaFuncSetAttribute(k_lib2d, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_lib2d<<<1, 256, smem>>>(tm, x, y, bytes, out); }
    else if (lib) { CK(cudaFuncSetAttribute(k_lib3d, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_lib3d<<<1, 256, smem>>>(tm, x, y, bytes, out); }
    else if (rank == 2) { CK(cudaFuncSetAttribute(k_raw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_raw<2><<<1, 256, smem>>>(tm, x, y, bytes, out); }
    else { CK(cudaFuncSetAttribute(k_raw<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_raw<3><<<1, 256, smem>>>(tm, x, y, bytes, out); }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> o(bytes);
    CK(cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int p = 0; p < planes; ++p) for (int rr = 0; rr < bh; ++rr) for (int c = 0; c < bw; ++c) {
        const int gy = y + rr, gx = x + c;
        uint8_t want = (gy >= 0 && gy < ch && gx >= 0 && gx < cw) ? h[p * ps + gy * pitch + gx] : 0;
        bad += o[(p * bh + rr) * bw + c] != want;
    }
    printf("%ld mismatches\n", bad);
    return 0;
}
