// ibk_force.cu -- Lagrangian force evaluation and marker-column algebra on the device:
// with X, U and F resident, the step either side of spread/interpolate needs no host round trip.
//
// Reference: IBStandardForceGen::computeLagrangianSpringForce / BeamForce / TargetPointForce
// (src/IB/IBStandardForceGen.cpp:813-930, 1037-1147, 1201-1299) called from
// IBStandardForceGen::computeLagrangianForce (:253-303) after IBMethod::computeLagrangianForce zeroed F
// (src/IB/IBMethod.cpp:834-858); the position updates are VecWAXPY / VecAXPBYPCZ on LData
// (IBMethod.cpp:714-826, reinitMidpointData :1900-1912).
//
// The reference loops over the force elements and scatters +-F into the nodes (serial, so no conflicts).
// Here every NODE gathers: one thread per marker walks the node's incidence lists (springs, then beams, then
// target points, each in element order = the order in which the reference's loop reaches this node), so the
// sum at every node is formed in the reference's order, with no atomics: bit-reproducible.  An element's force
// is recomputed by each of its 2 (3) nodes.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>

#include "ibk_engine.h"

namespace ibk
{
template <int NDIM>
__global__ void lagrangian_force_kernel(ForceTables t, const double* __restrict__ X, const double* __restrict__ U,
                                        double* __restrict__ F, long long stride, const uint32_t* __restrict__ id_of_pos,
                                        const int* __restrict__ pos_of_id, int n, int* __restrict__ missing)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = (int)id_of_pos[i];
    double acc[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) acc[d] = 0.0;
    if (l < t.n_nodes)
    {
        // springs: F_mastr += T/R * D, F_slave -= T/R * D with D = X_slave - X_mastr (IBStandardForceGen.cpp:858-884)
        if (t.spring_ptr)
            for (int q = t.spring_ptr[l]; q < t.spring_ptr[l + 1]; ++q)
            {
                const int item = t.spring_items[q];
                const int k = item >> 1, is_slave = item & 1;
                const int pm = pos_of_id[t.spring_mastr[k]], ps = pos_of_id[t.spring_slave[k]];
                if (pm < 0 || ps < 0)
                {
                    atomicAdd(missing, 1);
                    continue;
                }
                double D[NDIM], R2 = 0.0;
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
                {
                    D[d] = X[d * stride + ps] - X[d * stride + pm];
                    R2 = __dadd_rn(R2, __dmul_rn(D[d], D[d]));
                }
                const double R = sqrt(R2);
                if (R < DBL_EPSILON) continue;
                const double T_over_R = __ddiv_rn(__dmul_rn(t.spring_kappa[k], R - t.spring_rest[k]), R);
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
                {
                    const double f = __dmul_rn(T_over_R, D[d]);
                    acc[d] = is_slave ? __dsub_rn(acc[d], f) : __dadd_rn(acc[d], f);
                }
            }
        // beams: F = K (X_next + X_prev - 2 X_mastr - D2X0); mastr += 2F, next -= F, prev -= F (:1078-1096)
        if (t.beam_ptr)
            for (int q = t.beam_ptr[l]; q < t.beam_ptr[l + 1]; ++q)
            {
                const int item = t.beam_items[q];
                const int k = item >> 2, role = item & 3;
                const int pm = pos_of_id[t.beam_mastr[k]], pn = pos_of_id[t.beam_next[k]], pp = pos_of_id[t.beam_prev[k]];
                if (pm < 0 || pn < 0 || pp < 0)
                {
                    atomicAdd(missing, 1);
                    continue;
                }
                const double K = t.beam_rigidity[k];
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
                {
                    const double s = __dsub_rn(__dsub_rn(__dadd_rn(X[d * stride + pn], X[d * stride + pp]),
                                                         __dmul_rn(2.0, X[d * stride + pm])),
                                               t.beam_curvature[(size_t)NDIM * k + d]);
                    const double f = __dmul_rn(K, s);
                    acc[d] = (role == 0) ? __dadd_rn(acc[d], __dmul_rn(2.0, f)) : __dsub_rn(acc[d], f);
                }
            }
        // target points: F += kappa (X0 - X) - eta U (:1241-1246)
        if (t.target_ptr)
            for (int q = t.target_ptr[l]; q < t.target_ptr[l + 1]; ++q)
            {
                const int k = t.target_items[q];
#pragma unroll
                for (int d = 0; d < NDIM; ++d)
                {
                    const double f = __dsub_rn(__dmul_rn(t.target_kappa[k], __dsub_rn(t.target_X0[(size_t)NDIM * k + d], X[d * stride + i])),
                                               __dmul_rn(t.target_eta[k], U[d * stride + i]));
                    acc[d] = __dadd_rn(acc[d], f);
                }
            }
    }
#pragma unroll
    for (int d = 0; d < NDIM; ++d) F[d * stride + i] = acc[d];
}

cudaError_t launch_lagrangian_force(Launcher& L, int ndim, const ForceTables& t, const double* X, const double* U, double* F,
                                    long long stride, const uint32_t* id_of_pos, const int* pos_of_id, int n, int* d_missing)
{
    if (n <= 0) return cudaSuccess;
    const unsigned nb = (unsigned)((n + 127) / 128);
    if (ndim == 3)
        lagrangian_force_kernel<3><<<nb, 128, 0, L.stream>>>(t, X, U, F, stride, id_of_pos, pos_of_id, n, d_missing);
    else
        lagrangian_force_kernel<2><<<nb, 128, 0, L.stream>>>(t, X, U, F, stride, id_of_pos, pos_of_id, n, d_missing);
    L.launches++;
    return cudaGetLastError();
}

__global__ void pos_of_id_kernel(const uint32_t* __restrict__ id_of_pos, int n, int* __restrict__ pos_of_id)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos_of_id[id_of_pos[i]] = i;
}
cudaError_t launch_pos_of_id(Launcher& L, const uint32_t* id_of_pos, int n, int* pos_of_id, int id_bound)
{
    cudaError_t e = cudaMemsetAsync(pos_of_id, 0xff, sizeof(int) * (size_t)id_bound, L.stream); // -1: not on this rank
    if (e != cudaSuccess || n <= 0) return e;
    pos_of_id_kernel<<<(unsigned)((n + 255) / 256), 256, 0, L.stream>>>(id_of_pos, n, pos_of_id);
    L.launches++;
    return cudaGetLastError();
}

// dst = alpha * a + beta * b over the first n entries of every dimension's column (VecWAXPY / VecAXPBYPCZ)
__global__ void lincomb_kernel(double* __restrict__ dst, double alpha, const double* __restrict__ a, double beta,
                               const double* __restrict__ b, long long stride, int n, int ndim)
{
    const long long total = (long long)n * ndim;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int d = (int)(q / n), i = (int)(q - (long long)d * n);
        const long long o = d * stride + i;
        // the reference forms alpha*a + beta*b with separate multiplies and one add (PETSc BLAS-1 kernels)
        dst[o] = __dadd_rn(__dmul_rn(alpha, a[o]), __dmul_rn(beta, b[o]));
    }
}
cudaError_t launch_lincomb(Launcher& L, double* dst, double alpha, const double* a, double beta, const double* b, long long stride,
                           int n, int ndim)
{
    if (n <= 0) return cudaSuccess;
    const long long total = (long long)n * ndim;
    const unsigned nb = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
    lincomb_kernel<<<nb, 256, 0, L.stream>>>(dst, alpha, a, beta, b, stride, n, ndim);
    L.launches++;
    return cudaGetLastError();
}

// dst[d][i] = src[d][i] * ds[row(i)]: the F * ds product of LDataManager::spread (LDataManager.cpp:416-447); ds is given
// per host row (Lagrangian order), row_of_pos maps the storage position to it
__global__ void scale_rows_kernel(double* __restrict__ dst, const double* __restrict__ src, long long stride, int n, int ndim,
                                  const double* __restrict__ ds, const uint32_t* __restrict__ row_of_pos)
{
    const long long total = (long long)n * ndim;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x)
    {
        const int d = (int)(q / n), i = (int)(q - (long long)d * n);
        dst[d * stride + i] = __dmul_rn(src[d * stride + i], ds[row_of_pos[i]]);
    }
}
cudaError_t launch_scale_rows(Launcher& L, double* dst, const double* src, long long stride, int n, int ndim, const double* d_ds,
                              const uint32_t* row_of_pos)
{
    if (n <= 0) return cudaSuccess;
    const long long total = (long long)n * ndim;
    scale_rows_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, L.stream>>>(dst, src, stride, n, ndim, d_ds,
                                                                                                       row_of_pos);
    L.launches++;
    return cudaGetLastError();
}

__global__ void zero_rows_kernel(double* __restrict__ col, long long stride, int ndim, const int* __restrict__ ids, int n_ids,
                                 const int* __restrict__ pos_of_id, int id_bound)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ids) return;
    const int l = ids[k];
    if (l < 0 || l >= id_bound) return;
    const int i = pos_of_id[l];
    if (i < 0) return; // not on this rank
    for (int d = 0; d < ndim; ++d) col[d * stride + i] = 0.0;
}
cudaError_t launch_zero_rows(Launcher& L, double* col, long long stride, int ndim, const int* d_ids, int n_ids, const int* pos_of_id,
                             int id_bound)
{
    if (n_ids <= 0) return cudaSuccess;
    zero_rows_kernel<<<(unsigned)((n_ids + 255) / 256), 256, 0, L.stream>>>(col, stride, ndim, d_ids, n_ids, pos_of_id, id_bound);
    L.launches++;
    return cudaGetLastError();
}

} // namespace ibk
