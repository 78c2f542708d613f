// ibk_ctx.h -- the context object behind the C ABI (internal).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "ibk_engine.h"

namespace ibk
{
// growable device buffer
struct DevBuf
{
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as()
    {
        return reinterpret_cast<T*>(p);
    }
};

struct PatchState
{
    int lower[3], upper[3];
    PatchBin pb;
    // per axis: u and f arrays (pitched), extents incl. ghosts
    double* u[3] = { nullptr, nullptr, nullptr };
    double* f[3] = { nullptr, nullptr, nullptr };
    int n[3][3];
    long long pitch[3];
    size_t elems[3];
};

struct LevelState
{
    bool valid = false;
    int ndim = 0;
    int domain_lower[3], domain_upper[3], periodic[3], gcw[3];
    double x_lower[3], x_upper[3], dx[3];
    int G = 0;
    std::vector<PatchState> patches;
    std::vector<PatchBin> h_bins;
    PatchBin* d_bins = nullptr;
    // markers (storage order; after a rebin: sorted order)
    int n = 0;
    long long stride = 0;
    double* X = nullptr;
    double* U = nullptr;
    double* F = nullptr;
    double* tmp = nullptr;    // [ndim][stride] permutation scratch
    double* extra[3] = { nullptr, nullptr, nullptr }; // optional columns 3..5 (X_current, X_new, auxiliary), lazily allocated
    // second copy of every column in use: the re-bin permutes all columns in ONE kernel from the column into its shadow and
    // swaps the two (allocated at the first re-bin, dropped when the capacity changes)
    double* shadow[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    long long shadow_stride = 0;
    // Lagrangian force elements (ibk_force_set_*), device tables + the inverse of the storage order
    ForceTables force;
    std::vector<void*> force_allocs; // every device array behind `force`
    int* pos_of_id = nullptr;        // [pos_cap] Lagrangian index -> storage position (-1: not here)
    int pos_cap = 0;
    bool pos_valid = false;
    int* d_missing = nullptr;        // device counter: force elements with an endpoint that is not on this rank
    uint32_t* lag = nullptr;  // storage position -> row of the host AoS arrays (= Lagrangian index unless gid is set)
    uint32_t* gid = nullptr;  // optional: storage position -> global Lagrangian index (multi-rank; ibk_markers_set_ids)
    uint32_t id_bound = 0;    // exclusive upper bound of the global indices (sizes the tie bits of the sort key)
    // migration state between ibk_migrate_plan and ibk_migrate_unpack
    const uint32_t* mig_order = nullptr; // tail markers grouped by destination rank
    int mig_n = -1;
    uint32_t* lag_prev = nullptr; // the same map as it was when the binning products were written
    int* cells = nullptr;     // [n][ndim], in lag_prev storage order
    int* owner = nullptr;     // [n]
    int* escaped = nullptr;   // device counter
    Bins bins;
    bool binned = false;
    int n_owned = 0;
};

} // namespace ibk

struct ibk_ctx
{
    int device = 0;
    ibk::Launcher L;
    std::string err;
    // scratch for the raw / patch seams
    ibk::Bins sbins;
    ibk::DevBuf b_Xe, b_Xr, b_Xes, b_Xrs, b_src, b_patchbin;
    ibk::DevBuf b_io[8]; // staging for *_host entry points
    ibk::LevelState lv;
    bool timing = false;
    cudaEvent_t ev[3][2];
    bool ev_valid[3] = { false, false, false };
    bool ev_created = false;
    // asynchronous grid transfers (ibk_grid_upload_async / ibk_grid_download_async): one stream per direction,
    // one event per (direction, array) that the compute stream waits on before it touches the array again
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2], ev_out[2], ev_order;
    bool xfer_created = false;
    bool pend_in[2] = { false, false }, pend_out[2] = { false, false };
    ibk::DevBuf b_mig[10];  // marker migration scratch (keys/vals ping-pong, sort temp, box list, offsets; [8], [9]: send / receive rows of ibk_migrate)
    ibk::DevBuf b_stage[3]; // staging blocks of the grid transfers: compute stream, copy-in stream, copy-out stream
    ibk::DevBuf b_user[10]; // USER_DEFINED kernel (ibk_user.cu): ranges, weights, rows; contribution keys / values of the spread
};
