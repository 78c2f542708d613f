// Probe: TMA box loads of fp64 tiles, rank 2 and rank 3, box possibly larger than / outside the tensor.
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
template <int NDIM, int S>
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, double* out)
{
    extern __shared__ __align__(128) unsigned char raw[];
    double* su = (double*)raw;
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_expect_tx(&bar, (NDIM == 3 ? S * S * S : S * S) * 8);
        if (NDIM == 3) tma_load_3d(su, &map, &bar, c0, c1, c2); else tma_load_2d(su, &map, &bar, c0, c1);
    }
    mbar_wait(&bar, 0);
    const int n = NDIM == 3 ? S * S * S : S * S;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = su[i];
}
int main()
{
    PFN enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    printf("entry %p q=%d\n", (void*)enc, (int)q);
    const int n0 = 27, n1 = 26, n2 = 5, pitch = 32, S = 20;
    std::vector<double> h((size_t)pitch * n1 * n2);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *out; cudaMalloc(&d, h.size() * 8); cudaMalloc(&out, S * S * S * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    for (int ndim = 2; ndim <= 3; ++ndim)
    {
        CUtensorMap m;
        cuuint64_t dims[3] = { n0, n1, n2 }; cuuint64_t str[2] = { pitch * 8, (cuuint64_t)pitch * 8 * n1 };
        cuuint32_t box[3] = { S, S, S }, es[3] = { 1, 1, 1 };
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, ndim, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("ndim %d encode -> %d\n", ndim, (int)r);
        if (r) continue;
        const int cs[6][3] = { { -2, 3, -1 }, { 25, 3, 0 }, { 26, 26, 0 }, { 27, 0, 0 }, { 0, 26, 0 }, { -25, -30, -40 } };
        for (int t = 0; t < 6; ++t) {
        if (ndim == 2) probe<2, S><<<1, 128, S * S * 8>>>(m, cs[t][0], cs[t][1], 0, out);
        else { cudaFuncSetAttribute(probe<3, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * S * S * 8); probe<3, S><<<1, 128, S * S * S * 8>>>(m, cs[t][0], cs[t][1], cs[t][2], out); }
        cudaError_t e = cudaDeviceSynchronize();
        printf("ndim %d coords (%d,%d,%d) run -> %s\n", ndim, cs[t][0], cs[t][1], cs[t][2], cudaGetErrorString(e));
        if (e) return 1; }
        if (ndim == 2) probe<2, S><<<1, 128, S * S * 8>>>(m, -2, 3, 0, out); else probe<3, S><<<1, 128, S * S * S * 8>>>(m, -2, 3, -1, out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<double> o(S * S * S); cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
        // expected element (i,j[,k]) = h[((c2+k)*n1 + (c1+j))*pitch + c0+i] or 0 outside
        int bad = 0;
        for (int k = 0; k < (ndim == 3 ? S : 1); ++k) for (int j = 0; j < S; ++j) for (int i = 0; i < S; ++i)
        {
            int gi = -2 + i, gj = 3 + j, gk = ndim == 3 ? -1 + k : 0;
            double ex = (gi >= 0 && gi < n0 && gj >= 0 && gj < n1 && gk >= 0 && gk < n2) ? h[((size_t)gk * n1 + gj) * pitch + gi] : 0.0;
            if (o[(k * S + j) * S + i] != ex) ++bad;
        }
        printf("ndim %d mismatches %d\n", ndim, bad);
    }
    return 0;
}
