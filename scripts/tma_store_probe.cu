// Probe: TMA box load + box store of fp64 tiles whose box is larger than / partly outside a small tensor.
// mode 0: load only, 1: store only (block of ones), 2: load, add 1, store.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../ibamr_b200/csrc/ibk_device.cuh"
using namespace ibk;
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int BX, int BY>
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int mode)
{
    extern __shared__ __align__(128) unsigned char raw[];
    double* su = (double*)raw;
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (mode == 1) for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) su[i] = 1.0;
    __syncthreads();
    if (mode != 1)
    {
        if (threadIdx.x == 0) { mbar_expect_tx(&bar, BX * BY * 8); tma_load_2d(su, &map, &bar, c0, c1); }
        mbar_wait(&bar, 0);
        for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) su[i] += 1.0;
    }
    if (mode == 0) return;
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) { tma_store_2d(&map, su, c0, c1); tma_store_commit_and_wait_read(); }
}
int main()
{
    PFN enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    const int sizes[3][3] = { { 13, 12, 16 }, { 40, 30, 48 }, { 22, 19, 32 } };
    constexpr int BX = 20, BY = 18;
    for (int s = 0; s < 3; ++s)
    {
        const int n0 = sizes[s][0], n1 = sizes[s][1], pitch = sizes[s][2];
        std::vector<double> h((size_t)pitch * n1);
        double* d; cudaMalloc(&d, h.size() * 8);
        CUtensorMap m;
        cuuint64_t dims[2] = { (cuuint64_t)n0, (cuuint64_t)n1 }; cuuint64_t str[1] = { (cuuint64_t)pitch * 8 };
        cuuint32_t box[2] = { BX, BY }, es[2] = { 1, 1 };
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("tensor %dx%d pitch %d encode -> %d\n", n0, n1, pitch, (int)r);
        if (r) continue;
        const int cs[5][2] = { { 0, 0 }, { 10, 11 }, { 0, -5 }, { -6, 0 }, { -6, -5 } };
        for (int mode = 1; mode < 3; ++mode)
            for (int t = 0; t < 5; ++t)
            {
                for (size_t i = 0; i < h.size(); ++i) h[i] = 100.0 + (double)i;
                cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
                probe<BX, BY><<<1, 128, BX * BY * 8>>>(m, cs[t][0], cs[t][1], mode);
                cudaError_t e = cudaDeviceSynchronize();
                int bad = 0;
                if (!e && mode > 0)
                {
                    std::vector<double> o(h.size());
                    cudaMemcpy(o.data(), d, o.size() * 8, cudaMemcpyDeviceToHost);
                    for (int j = 0; j < n1; ++j) for (int i = 0; i < pitch; ++i)
                    {
                        const bool in = i < n0 && i >= cs[t][0] && i < cs[t][0] + BX && j >= cs[t][1] && j < cs[t][1] + BY;
                        const double ex = in ? (mode == 1 ? 1.0 : h[(size_t)j * pitch + i] + 1.0) : h[(size_t)j * pitch + i];
                        if (o[(size_t)j * pitch + i] != ex) ++bad;
                    }
                }
                printf("  mode %d coords (%d,%d) -> %s, mismatches %d\n", mode, cs[t][0], cs[t][1], cudaGetErrorString(e), bad);
                if (e) { cudaDeviceReset(); return 1; }
            }
        cudaFree(d);
    }
    return 0;
}
