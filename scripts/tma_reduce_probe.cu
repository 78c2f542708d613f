// Probe: does TMA's reducing store (cp.reduce.async.bulk.tensor ... .add) work on an fp64 tensor on this GPU?
// A dense 18^3 box of shared memory is ADDED to a box of a small 3D fp64 tensor (inside, and partly outside: clipped).
// If it does, the spread kernel can add a CTA's share of a cluster tile to f without loading f first.
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
constexpr int B = 18;
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2)
{
    extern __shared__ __align__(128) unsigned char raw[];
    double* su = (double*)raw;
    for (int i = threadIdx.x; i < B * B * B; i += blockDim.x) su[i] = 0.5 + 1e-3 * i;
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&map),
                     "r"(smem_u32(su)), "r"(c0), "r"(c1), "r"(c2)
                     : "memory");
        tma_store_commit_and_wait_read();
    }
}
int main()
{
    PFN enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    const int n0 = 40, n1 = 30, n2 = 24, pitch = 48;
    std::vector<double> h((size_t)pitch * n1 * n2);
    double* d;
    cudaMalloc(&d, h.size() * 8);
    CUtensorMap m;
    cuuint64_t dims[3] = { n0, n1, n2 };
    cuuint64_t str[2] = { (cuuint64_t)pitch * 8, (cuuint64_t)pitch * 8 * n1 };
    cuuint32_t box[3] = { B, B, B }, es[3] = { 1, 1, 1 };
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    if (r) return 1;
    const int cs[4][3] = { { 2, 3, 1 }, { 30, 20, 10 }, { 1, 0, 0 }, { 0, 0, 0 } };
    for (int t = 0; t < 4; ++t)
    {
        for (size_t i = 0; i < h.size(); ++i) h[i] = 100.0 + (double)i;
        cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        probe<<<1, 256, B * B * B * 8>>>(m, cs[t][0], cs[t][1], cs[t][2]);
        cudaError_t e = cudaDeviceSynchronize();
        int bad = 0, hit = 0;
        if (!e)
        {
            std::vector<double> o(h.size());
            cudaMemcpy(o.data(), d, o.size() * 8, cudaMemcpyDeviceToHost);
            for (int k = 0; k < n2; ++k)
                for (int j = 0; j < n1; ++j)
                    for (int i = 0; i < pitch; ++i)
                    {
                        const size_t g = ((size_t)k * n1 + j) * pitch + i;
                        const int bi = i - cs[t][0], bj = j - cs[t][1], bk = k - cs[t][2];
                        const bool in = i < n0 && bi >= 0 && bi < B && bj >= 0 && bj < B && bk >= 0 && bk < B;
                        const double ex = in ? h[g] + (0.5 + 1e-3 * ((bk * B + bj) * B + bi)) : h[g];
                        hit += in;
                        if (o[g] != ex) ++bad;
                    }
        }
        printf("reduce-add fp64 at (%d,%d,%d) -> %s, %d points in the box, mismatches %d\n", cs[t][0], cs[t][1], cs[t][2], cudaGetErrorString(e), hit, bad);
        if (e)
        {
            cudaDeviceReset();
            return 1;
        }
    }
    return 0;
}
