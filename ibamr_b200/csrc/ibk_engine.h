// ibk_engine.h -- host-side internal interfaces between the translation units of libibk.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ibk_device.cuh"

namespace ibk
{
// ---------------------------------------------------------------------------------------------
// binning geometry of one patch (device-visible POD)
// ---------------------------------------------------------------------------------------------
struct PatchBin
{
    int ndim;
    int lower[3], upper[3];   // patch box (level cell indices)
    int accept_lo[3], accept_hi[3]; // a marker is binned into this patch iff its cell is in this box
    int G;                    // margin: cc = cell - lower + G  (cc >= 0 for every accepted cell)
    int nb[3], nt[3];         // bricks / marker tiles per dimension
    int brick_base;           // first brick id of this patch
    int nbricks;              // nt[0]*nt[1]*nt[2] * 4^ndim
};

// How cells are computed from positions.
struct CellGeom
{
    int ndim;
    double x_lower[3], x_upper[3], dx[3];
    int ilower[3], iupper[3];
    int two_branch; // 1: IndexUtilities::getCellIndex (two-branch), 0: floor((X - x_lower)/dx) + ilower
};

struct DomainGeom
{
    int ndim;
    double x_lower[3], x_upper[3];
    int periodic[3];
};

void fill_patch_bin(PatchBin& pb, int ndim, const int* lower, const int* upper, const int* accept_lo, const int* accept_hi,
                    int G, int brick_base);

// Sorted marker bins (device).  Capacity-managed by the owner.
struct Bins
{
    int n_entries = 0;        // entries handed to the binning
    int n_active = 0;         // entries accepted by some patch (sorted to the front); filled lazily
    int total_bricks = 0;
    uint64_t* keys[2] = { nullptr, nullptr }; // ping-pong
    uint32_t* vals[2] = { nullptr, nullptr };
    int sorted_in = 0;        // which ping-pong buffer holds the sorted result
    int* brick_start = nullptr; // [total_bricks + 1]
    int tie_bits = 32;        // low key bits holding the tie-break id
    void* sort_temp = nullptr;
    size_t sort_temp_bytes = 0;
    int capacity = 0;
    int brick_capacity = 0;
    // per patch (keyed by its first brick id): sorted positions [first, last) of its markers, read
    // back once per binning so that launches need no device->host round trip
    std::vector<int> range_base, range_first, range_last;
    // bricks holding more than DENSE_BRICK_MARKERS markers (structures: tens of markers per cell), in no particular
    // order; the spread gives them a kernel of their own (ibk_spread.cu, spread_dense_kernel)
    int* dense_list = nullptr; // [dense_capacity] brick ids
    int* dense_count = nullptr; // device counter
    int dense_capacity = 0;
    int n_dense = 0;
    // spread exceptions: (entry, component) pairs whose stencil does not fit the accumulator of their tile are flagged here by
    // the tile kernels and spread by spread_fixup_kernel.  One byte per sorted position (bit a = component a), held as 32-bit
    // words: the list cannot overflow.  All zero between spreads (the fix-up clears what it consumes).
    // 3D march spread: [0] the ticket counter of its persistent CTAs, [1 ...] one done flag per (march tile, component)
    int* march_sync = nullptr;
    size_t march_sync_capacity = 0;
    unsigned* exc_flags = nullptr; // [capacity / 4 + 1]
    int* exc_count = nullptr;      // number of flagged pairs (an upper bound: a pair may be flagged twice)
};
constexpr int DENSE_BRICK_MARKERS = 48;

struct Launcher
{
    cudaStream_t stream = 0;
    long long launches = 0;
    // SMs the persistent kernels leave free: with a communicator the messages of the halo exchange (NCCL kernels on the
    // communication stream) must be able to start while the spread's persistent CTAs hold their SMs for milliseconds
    int reserve_sms = 0;
};

// ibk_sort.cu
size_t radix_sort_temp_bytes(int n);
int radix_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int n, int begin_bit,
                     int end_bit, void* temp, cudaStream_t stream, long long* launches);

// ibk_bin.cu
cudaError_t bins_reserve(Bins& b, int n_entries, int total_bricks);
void bins_free(Bins& b);
// X: SoA [ndim][stride].  tie[i] = tie-break id of entry i (Lagrangian index) or nullptr for i.
// cells_out ([n][ndim], optional) and owner_out ([n], optional) receive the binning products.
cudaError_t bins_build(Bins& b, Launcher& L, const CellGeom& cg, const PatchBin* d_patches, int n_patches,
                       const PatchBin* h_patches, const double* d_X, long long x_stride, const uint32_t* d_tie,
                       uint32_t tie_bound, int n_entries, int* d_cells_out, int* d_owner_out);
// check_only: count the points outside a non-periodic domain without touching X
cudaError_t wrap_positions(Launcher& L, const DomainGeom& dg, double* d_X, long long x_stride, int n, int* d_escaped,
                           bool check_only = false);
// out[c][i] = in[c][perm[i]] for c < ncols (SoA gather through the sort permutation)
cudaError_t gather_columns(Launcher& L, const double* d_in, long long in_stride, double* d_out, long long out_stride,
                           const uint32_t* d_perm, int n, int ncols);
// out[k][c][i] = in[k][c][perm[i]] for up to 6 column sets of ncols columns in one launch (one read of perm, all gathers of a
// marker in flight together)
struct GatherSets
{
    const double* in[6];
    double* out[6];
    int nsets;
};
cudaError_t gather_column_sets(Launcher& L, const GatherSets& sets, long long stride, const uint32_t* d_perm, int n, int ncols);
// within every run of keys that agree above `tie_bits`, orders the (key, value) pairs by the full key (insertion sort by the
// thread of the run's first element; runs are cells: a handful of markers)
cudaError_t sort_ties(Launcher& L, uint64_t* keys, uint32_t* vals, int n, int tie_bits);
cudaError_t scatter_columns(Launcher& L, const double* d_in, long long in_stride, double* d_out, long long out_stride,
                            const uint32_t* d_perm, int n, int ncols);
cudaError_t extract_low32(Launcher& L, const uint64_t* d_keys, uint32_t* d_out, int n, int bits);

// ibk_halo.cu: all regions of one message in one launch
struct HaloItem
{
    double* ptr;       // array of the (patch, axis)
    long long pitch;
    int n1;
    int off[3], ext[3]; // region in array coordinates
    long long buf_off;  // first element of the region in the message buffer
    long long count;
    unsigned block0, nblocks; // the CTAs of the launch that work on this item (in proportion to its size)
};
// work (in warp work items of the row walker, ibk_device.cuh::slab_rows) and CTAs of a region
long long region_work(const int* ext);
unsigned region_blocks(const int* ext);
// op 0: buffer <- regions (pack); 1: regions <- buffer (copy); 2: regions += buffer.  total_blocks = sum of nblocks.
cudaError_t launch_halo_items(Launcher& L, const HaloItem* d_items, int n_items, unsigned total_blocks, double* buf, int op);

// ibk_force.cu
// Force elements by Lagrangian index plus, per node, the elements it takes part in (CSR), all on the device.
struct ForceTables
{
    int n_nodes = 0; // the CSR pointers cover Lagrangian indices [0, n_nodes)
    // springs
    const int* spring_ptr = nullptr;   // [n_nodes + 1]
    const int* spring_items = nullptr; // spring k * 2 + (0: node is the master, 1: the slave), ascending k per node
    const int* spring_mastr = nullptr;
    const int* spring_slave = nullptr;
    const double* spring_kappa = nullptr;
    const double* spring_rest = nullptr;
    // beams
    const int* beam_ptr = nullptr;
    const int* beam_items = nullptr; // beam k * 4 + (0: master, 1: next, 2: prev)
    const int* beam_mastr = nullptr;
    const int* beam_next = nullptr;
    const int* beam_prev = nullptr;
    const double* beam_rigidity = nullptr;
    const double* beam_curvature = nullptr; // [n_beams][ndim]
    // target points
    const int* target_ptr = nullptr;
    const int* target_items = nullptr; // target k
    const double* target_kappa = nullptr;
    const double* target_eta = nullptr;
    const double* target_X0 = nullptr; // [n_targets][ndim]
};
cudaError_t launch_lagrangian_force(Launcher& L, int ndim, const ForceTables& t, const double* X, const double* U, double* F,
                                    long long stride, const uint32_t* id_of_pos, const int* pos_of_id, int n, int* d_missing);
cudaError_t launch_pos_of_id(Launcher& L, const uint32_t* id_of_pos, int n, int* pos_of_id, int id_bound);
cudaError_t launch_lincomb(Launcher& L, double* dst, double alpha, const double* a, double beta, const double* b, long long stride,
                           int n, int ndim);
cudaError_t launch_scale_rows(Launcher& L, double* dst, const double* src, long long stride, int n, int ndim, const double* d_ds,
                              const uint32_t* row_of_pos);
cudaError_t launch_zero_rows(Launcher& L, double* col, long long stride, int ndim, const int* d_ids, int n_ids, const int* pos_of_id,
                             int id_bound);

// ibk_migrate.cu
cudaError_t migrate_dest(Launcher& L, const CellGeom& cg, const int* d_plo, const int* d_phi, const int* d_prank, int n_patches,
                         int n_ranks, const double* X, long long stride, int first, int n_tail, uint64_t* keys, uint32_t* vals);
cudaError_t bucket_offsets(Launcher& L, const uint64_t* keys_sorted, int n, int n_buckets, int* d_start);
cudaError_t migrate_pack(Launcher& L, const uint32_t* order, int n_send, const double* X, const double* U, const double* F,
                         long long stride, int ndim, const uint32_t* gid, double* buf);
cudaError_t migrate_append(Launcher& L, const double* buf, int n_recv, int at, double* X, double* U, double* F, long long stride,
                           int ndim, uint32_t* gid);
cudaError_t id_keys(Launcher& L, const uint32_t* gid, int n, uint64_t* keys, uint32_t* vals);
cudaError_t rank_scatter(Launcher& L, const uint32_t* sorted_pos, int n, uint32_t* row);
cudaError_t gather_u32(Launcher& L, const uint32_t* in, const uint32_t* perm, int n, uint32_t* out);

// ibk_interp.cu / ibk_spread.cu
// Marker data for the tile kernels, in SORTED order (entry i of the bins).
struct MarkerView
{
    const double* X;      // SoA [ndim][x_stride]: X + Xshift
    const double* Xraw;   // SoA, unshifted positions (BSPLINE_4 quirk) or nullptr (= X)
    long long x_stride;
    double* V;            // values: column c of entry i is V[c * v_cstride + src(i) * v_istride]
    long long v_cstride;
    long long v_istride;
    const uint32_t* src;  // optional gather index (sorted position -> value row); nullptr = identity
    // optional restriction to a subset of the marker tiles (overlap of the inter-rank halo exchange with the
    // tiles that do not touch it): part 0 = all tiles, 1 = tiles with sel_lo <= index <= sel_hi in every
    // dimension, 2 = the others
    int part = 0;
    int sel_lo[3] = { 0, 0, 0 }, sel_hi[3] = { 0, 0, 0 };
    // true: every entry was binned into the interior of a patch whose ghost width covers the kernel, so no stencil reaches
    // outside the arrays (the resident level).  The spread may then start its first accumulator block per dimension at the
    // array's first element instead of M points before the first tile.
    bool clip_free = false;
};

struct TmaMaps; // opaque, ibk_interp.cu

cudaError_t launch_interp(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, const MarkerView& mv,
                          std::string& err);
cudaError_t launch_spread(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, const MarkerView& mv,
                          std::string& err);
long long count_touched(Launcher& L, int kernel, const TileParams& tp, const Bins& bins, std::string& err);

// ibk_halo.cu
struct RegionCopy
{
    double* dst;
    const double* src;
    long long dst_pitch, src_pitch;
    int dst_n1, src_n1;   // rows per plane
    int dst_off[3], src_off[3];
    int ext[3];
};
cudaError_t launch_region_ops(Launcher& L, const std::vector<RegionCopy>& ops, int mode /*0 copy, 1 add*/);
cudaError_t launch_pack(Launcher& L, const double* arr, long long pitch, int n1, const int* off, const int* ext, double* buf,
                        int ndim);
cudaError_t launch_unpack(Launcher& L, double* arr, long long pitch, int n1, const int* off, const int* ext,
                          const double* buf, int ndim, int mode);
cudaError_t launch_fill(Launcher& L, double* ptr, size_t count, double value);
// layout conversion: reference (dense Fortran) <-> pitched device array
// `stage` (device, `stage_bytes`): optional staging block; with it the bytes cross the link as flat copies
cudaError_t copy_dense_to_pitched(Launcher& L, const double* src_dense, double* dst, long long pitch, const int* n, int ndim,
                                  cudaMemcpyKind kind, void* stage = nullptr, size_t stage_bytes = 0);
cudaError_t copy_pitched_to_dense(Launcher& L, const double* src, long long pitch, double* dst_dense, const int* n, int ndim,
                                  cudaMemcpyKind kind, void* stage = nullptr, size_t stage_bytes = 0);
// AoS [n][depth] <-> SoA [depth][stride]
cudaError_t aos_to_soa(Launcher& L, const double* d_aos, double* d_soa, long long stride, int n, int depth);
cudaError_t soa_to_aos(Launcher& L, const double* d_soa, long long stride, double* d_aos, int n, int depth);
// entries for the raw / indexed seams: Xe[d][l] = X[idx[l]][d] + shift[l][d], Xr[d][l] = X[idx[l]][d]
cudaError_t build_entries(Launcher& L, const double* d_X_aos, const int* d_idx, const double* d_shift, int n, int ndim,
                          double* d_Xe, double* d_Xr, long long stride);
cudaError_t zero_discarded(Launcher& L, const int* brick_start, int total_bricks, int n, const uint32_t* src, double* V,
                           long long v_cstride, long long v_istride, int ncol);
// the same for the entries whose cell lies in [box_lo, box_hi] only (Xs: the entries' positions in sorted order)
cudaError_t zero_discarded_in_box(Launcher& L, const int* brick_start, int total_bricks, int n, const double* Xs, long long stride,
                                  const CellGeom& cg, const int* box_lo, const int* box_hi, const uint32_t* src, double* V,
                                  long long v_cstride, long long v_istride, int ncol);
cudaError_t compose_index(Launcher& L, const int* d_idx, const uint32_t* d_perm, uint32_t* d_out, int n);

} // namespace ibk
