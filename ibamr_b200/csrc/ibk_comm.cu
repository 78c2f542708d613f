// ibk_comm.cu -- the multi-rank layer behind the C ABI: halo plan, inter-process ghost fill / ghost accumulation and
// marker migration, one process (rank) per GPU.
//
// Replaces, across processes, what the reference reaches from C++ around the hot path:
//   fill        u ghost cells <- owner interiors, periodic wrap included
//               (u_ghost_fill_scheds[ln]->fillData, ibtk/src/lagrangian/LDataManager.cpp:744)
//   accumulate  owner interiors += every other copy of the DOF (ghost copies and the interior copy of a face
//               shared by two patches): SAMRAIGhostDataAccumulator::accumulateGhostData's reverse scatter
//               (ibtk/src/math/SAMRAIGhostDataAccumulator.cpp:327-344; called LDataManager.cpp:597-620)
//   migrate     LDataManager::endDataRedistribution's scatter of the marker rows (LDataManager.cpp:1824-1837)
//
// The PLAN is host code without a context (ibk_halo_plan_*): derived identically on every rank from the global box
// list, so no metadata is exchanged; items are in a canonical order, unpack-adds run in ascending source rank then
// item order: sums are reproducible.  The TRANSPORT is NCCL (ncclSend / ncclRecv in one group per exchange on a
// communication stream of the context, so the messages fly while the tiles that do not touch the exchanged regions
// are processed), resolved with dlopen so that libibk.so loads without NCCL; or, for several contexts of ONE
// process (tests on one GPU; one process driving several GPUs), device-to-device copies ordered by events.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ibk_ctx.h"

// ---------------------------------------------------------------------------------------------
// plan (host only)
// ---------------------------------------------------------------------------------------------
struct ibk_halo_item
{
    int axis;
    int src, dst;             // global patch numbers
    int src_lo[3], src_hi[3]; // region in the source patch's index space
    int dst_lo[3], dst_hi[3]; // the same region where the destination patch sees it (periodic image)
    long long count;
};
struct ibk_halo_message
{
    int src_rank, dst_rank;
    std::vector<ibk_halo_item> items;
    long long count = 0;
};
struct ibk_halo_plan
{
    int ndim = 0, my_rank = 0;
    std::vector<std::array<int, 3>> lower, upper;
    std::vector<int> rank, local_id;
    std::vector<ibk_halo_message> table[2]; // 0 fill, 1 accumulate; sorted by (src_rank, dst_rank)
};

namespace
{
struct BoxI
{
    int lo[3], hi[3];
};
bool intersect(int ndim, const BoxI& a, const BoxI& b, BoxI& r)
{
    for (int d = 0; d < ndim; ++d)
    {
        r.lo[d] = std::max(a.lo[d], b.lo[d]);
        r.hi[d] = std::min(a.hi[d], b.hi[d]);
        if (r.hi[d] < r.lo[d]) return false;
    }
    return true;
}
// a \ b as disjoint boxes, highest dimension first
void box_minus(int ndim, const BoxI& a, const BoxI& b, std::vector<BoxI>& out)
{
    BoxI in;
    if (!intersect(ndim, a, b, in))
    {
        out.push_back(a);
        return;
    }
    BoxI cur = a;
    for (int d = ndim - 1; d >= 0; --d)
    {
        if (cur.lo[d] < in.lo[d])
        {
            BoxI p = cur;
            p.hi[d] = in.lo[d] - 1;
            out.push_back(p);
        }
        if (cur.hi[d] > in.hi[d])
        {
            BoxI p = cur;
            p.lo[d] = in.hi[d] + 1;
            out.push_back(p);
        }
        cur.lo[d] = in.lo[d];
        cur.hi[d] = in.hi[d];
    }
}
void side_boxes(int ndim, const int* lower, const int* upper, int axis, const int* gcw, BoxI& interior, BoxI& all)
{
    for (int d = 0; d < ndim; ++d)
    {
        interior.lo[d] = lower[d];
        interior.hi[d] = upper[d] + (d == axis ? 1 : 0);
        all.lo[d] = interior.lo[d] - gcw[d];
        all.hi[d] = interior.hi[d] + gcw[d];
    }
}
} // namespace

extern "C" int ibk_halo_plan_create(int ndim, int n_patches, const int* lower, const int* upper, const int* rank, const int* domain_ncells,
                                    const int* periodic, const int* gcw, int my_rank, ibk_halo_plan** out)
{
    if (!out || (ndim != 2 && ndim != 3) || n_patches < 0 || (n_patches > 0 && (!lower || !upper || !rank)) || !domain_ncells ||
        !periodic || !gcw)
        return IBK_ERR_INVALID;
    ibk_halo_plan* pl = new ibk_halo_plan();
    pl->ndim = ndim;
    pl->my_rank = my_rank;
    std::map<int, int> per_rank;
    for (int p = 0; p < n_patches; ++p)
    {
        std::array<int, 3> lo = { 0, 0, 0 }, hi = { 0, 0, 0 };
        for (int d = 0; d < ndim; ++d)
        {
            lo[d] = lower[p * ndim + d];
            hi[d] = upper[p * ndim + d];
        }
        pl->lower.push_back(lo);
        pl->upper.push_back(hi);
        pl->rank.push_back(rank[p]);
        pl->local_id.push_back(per_rank[rank[p]]++);
    }
    std::map<std::pair<int, int>, ibk_halo_message> msgs[2];
    int nshift[3] = { 1, 1, 1 };
    for (int d = 0; d < ndim; ++d) nshift[d] = periodic[d] ? 3 : 1;
    for (int axis = 0; axis < ndim; ++axis)
        for (int dst = 0; dst < n_patches; ++dst)
        {
            BoxI d_in, d_all;
            side_boxes(ndim, pl->lower[dst].data(), pl->upper[dst].data(), axis, gcw, d_in, d_all);
            std::vector<BoxI> d_ghost;
            box_minus(ndim, d_all, d_in, d_ghost);
            for (int src = 0; src < n_patches; ++src)
            {
                if (rank[src] == rank[dst]) continue; // same process: ibk_halo_local
                if (rank[src] != my_rank && rank[dst] != my_rank) continue;
                BoxI s_in, s_all;
                side_boxes(ndim, pl->lower[src].data(), pl->upper[src].data(), axis, gcw, s_in, s_all);
                for (int o2 = 0; o2 < nshift[2]; ++o2)
                    for (int o1 = 0; o1 < nshift[1]; ++o1)
                        for (int o0 = 0; o0 < nshift[0]; ++o0)
                        {
                            // (offsets in the order -1, 0, 1 per dimension, the last dimension slowest ... the item order is
                            // fixed by the sort below anyway)
                            const int o[3] = { periodic[0] ? o0 - 1 : 0, ndim > 1 && periodic[1] ? o1 - 1 : 0, ndim > 2 && periodic[2] ? o2 - 1 : 0 };
                            int sh[3] = { 0, 0, 0 };
                            for (int d = 0; d < ndim; ++d) sh[d] = o[d] * domain_ncells[d];
                            auto add = [&](int table, const BoxI& r) {
                                ibk_halo_message& m = msgs[table][{ rank[src], rank[dst] }];
                                m.src_rank = rank[src];
                                m.dst_rank = rank[dst];
                                ibk_halo_item it;
                                std::memset(&it, 0, sizeof(it));
                                it.axis = axis;
                                it.src = src;
                                it.dst = dst;
                                it.count = 1;
                                for (int d = 0; d < ndim; ++d)
                                {
                                    it.src_lo[d] = r.lo[d] - sh[d];
                                    it.src_hi[d] = r.hi[d] - sh[d];
                                    it.dst_lo[d] = r.lo[d];
                                    it.dst_hi[d] = r.hi[d];
                                    it.count *= (long long)(r.hi[d] - r.lo[d] + 1);
                                }
                                m.items.push_back(it);
                            };
                            BoxI si = s_in, sa = s_all, r;
                            for (int d = 0; d < ndim; ++d)
                            {
                                si.lo[d] += sh[d];
                                si.hi[d] += sh[d];
                                sa.lo[d] += sh[d];
                                sa.hi[d] += sh[d];
                            }
                            // fill: ghost region of dst <- interior of src (shifted)
                            for (const BoxI& g : d_ghost)
                                if (intersect(ndim, g, si, r)) add(0, r);
                            // accumulate: interior of dst += every copy src holds of it (ghosts, shared face)
                            if (intersect(ndim, d_in, sa, r)) add(1, r);
                        }
            }
        }
    for (int t = 0; t < 2; ++t)
        for (auto& kv : msgs[t])
        {
            ibk_halo_message& m = kv.second;
            // canonical item order: (axis, dst local id, src rank, src local id, dst_lo, dst_hi)
            std::stable_sort(m.items.begin(), m.items.end(), [&](const ibk_halo_item& x, const ibk_halo_item& y) {
                if (x.axis != y.axis) return x.axis < y.axis;
                if (pl->local_id[x.dst] != pl->local_id[y.dst]) return pl->local_id[x.dst] < pl->local_id[y.dst];
                if (pl->rank[x.src] != pl->rank[y.src]) return pl->rank[x.src] < pl->rank[y.src];
                if (pl->local_id[x.src] != pl->local_id[y.src]) return pl->local_id[x.src] < pl->local_id[y.src];
                for (int d = 0; d < 3; ++d)
                    if (x.dst_lo[d] != y.dst_lo[d]) return x.dst_lo[d] < y.dst_lo[d];
                for (int d = 0; d < 3; ++d)
                    if (x.dst_hi[d] != y.dst_hi[d]) return x.dst_hi[d] < y.dst_hi[d];
                return false;
            });
            m.count = 0;
            for (const auto& it : m.items) m.count += it.count;
            pl->table[t].push_back(m); // (std::map iterates in (src_rank, dst_rank) order)
        }
    *out = pl;
    return IBK_OK;
}
extern "C" void ibk_halo_plan_destroy(ibk_halo_plan* plan)
{
    delete plan;
}
extern "C" int ibk_halo_plan_messages(const ibk_halo_plan* plan, int table)
{
    if (!plan || table < 0 || table > 1) return IBK_ERR_INVALID;
    return (int)plan->table[table].size();
}
extern "C" int ibk_halo_plan_message(const ibk_halo_plan* plan, int table, int k, int* src_rank, int* dst_rank, int* n_items, long long* count)
{
    if (!plan || table < 0 || table > 1 || k < 0 || k >= (int)plan->table[table].size()) return IBK_ERR_INVALID;
    const ibk_halo_message& m = plan->table[table][k];
    if (src_rank) *src_rank = m.src_rank;
    if (dst_rank) *dst_rank = m.dst_rank;
    if (n_items) *n_items = (int)m.items.size();
    if (count) *count = m.count;
    return IBK_OK;
}
extern "C" int ibk_halo_plan_items(const ibk_halo_plan* plan, int table, int k, int* axis, int* src_local, int* dst_local, int* src_lo,
                                   int* src_hi, int* dst_lo, int* dst_hi)
{
    if (!plan || table < 0 || table > 1 || k < 0 || k >= (int)plan->table[table].size()) return IBK_ERR_INVALID;
    const ibk_halo_message& m = plan->table[table][k];
    const int ndim = plan->ndim;
    for (size_t i = 0; i < m.items.size(); ++i)
    {
        const ibk_halo_item& it = m.items[i];
        if (axis) axis[i] = it.axis;
        if (src_local) src_local[i] = plan->local_id[it.src];
        if (dst_local) dst_local[i] = plan->local_id[it.dst];
        for (int d = 0; d < ndim; ++d)
        {
            if (src_lo) src_lo[i * ndim + d] = it.src_lo[d];
            if (src_hi) src_hi[i * ndim + d] = it.src_hi[d];
            if (dst_lo) dst_lo[i * ndim + d] = it.dst_lo[d];
            if (dst_hi) dst_hi[i * ndim + d] = it.dst_hi[d];
        }
    }
    return IBK_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------------------------
namespace
{
struct NcclApi
{
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    struct Uid // ncclUniqueId, passed by value
    {
        char internal[128];
    };
    int (*CommInitRank)(void**, int, Uid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};
NcclApi& nccl()
{
    static NcclApi api;
    if (api.lib || !api.error.empty()) return api;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* n : names)
        if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.lib)
    {
        api.error = "NCCL not found (dlopen libnccl.so.2): put its directory on LD_LIBRARY_PATH, or import torch first";
        return api;
    }
    auto sym = [&](const char* s) -> void* {
        void* p = dlsym(api.lib, s);
        if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s;
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    return api;
}
constexpr int NCCL_FLOAT64 = 8; // ncclDataType_t values (nccl.h)

struct Message
{
    int peer = -1;
    long long count = 0;
    double* d_buf = nullptr;
    // item arrays of this rank's side of the message (ibk_halo_pack_many / ibk_halo_unpack_many)
    std::vector<int> patch, axis, lo, hi;
    std::vector<long long> offs;
    cudaEvent_t ev_ready = nullptr;    // loopback: the send buffer is packed
    cudaEvent_t ev_consumed = nullptr; // loopback: the peer has copied the send buffer
    bool consumed_pending = false;
};
struct Comm
{
    int rank = 0, nranks = 1;
    int transport = 0; // 1 NCCL, 2 loopback
    void* nccl_comm = nullptr;
    std::vector<ibk_ctx*> peers; // loopback: the contexts of all ranks (this one included)
    cudaStream_t s_comm = nullptr;
    cudaEvent_t ev_packed[2] = { nullptr, nullptr }, ev_arrived[2] = { nullptr, nullptr };
    ibk_halo_plan* plan = nullptr;
    std::vector<Message> send[2], recv[2]; // per table
    // all messages of a table and direction share ONE buffer (a message is a slice of it) and ONE item list, so that an
    // exchange packs with one launch and unpacks with one launch however many neighbours there are ([table][0 send / 1 receive])
    double* all_buf[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } };
    struct Merged
    {
        std::vector<int> patch, axis, lo, hi;
        std::vector<long long> offs;
    } merged[2][2];
    bool posted[2] = { false, false };
    long long n_posted[2] = { 0, 0 }; // exchanges posted so far (loopback: a peer must not be behind when this rank finishes)
    // the global box list (migration)
    std::vector<int> g_lower, g_upper, g_rank;
    double* d_counts = nullptr; // [nranks * nranks] migration counts (as doubles: one NCCL datatype for everything)
};
std::map<ibk_ctx*, Comm*>& comms()
{
    static std::map<ibk_ctx*, Comm*> m;
    return m;
}
Comm* comm_of(ibk_ctx* ctx)
{
    auto it = comms().find(ctx);
    return it == comms().end() ? nullptr : it->second;
}
int cfail(ibk_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    return code;
}
#define CCK(call)                                                                                                  \
    do                                                                                                             \
    {                                                                                                              \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return cfail(ctx, IBK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define NCK(call)                                                                                                                   \
    do                                                                                                                              \
    {                                                                                                                               \
        int r_ = (call);                                                                                                            \
        if (r_ != 0) return cfail(ctx, IBK_ERR_CUDA, std::string(#call) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r_) : "NCCL error")); \
    } while (0)

void free_messages(Comm* c)
{
    for (int t = 0; t < 2; ++t)
        for (auto* v : { &c->send[t], &c->recv[t] })
        {
            for (Message& m : *v)
            {
                if (m.ev_ready) cudaEventDestroy(m.ev_ready);
                if (m.ev_consumed) cudaEventDestroy(m.ev_consumed);
            }
            v->clear();
        }
    for (int t = 0; t < 2; ++t)
        for (int dir = 0; dir < 2; ++dir)
        {
            if (c->all_buf[t][dir]) cudaFree(c->all_buf[t][dir]);
            c->all_buf[t][dir] = nullptr;
            c->merged[t][dir] = Comm::Merged();
        }
    if (c->plan) ibk_halo_plan_destroy(c->plan);
    c->plan = nullptr;
}
int comm_common_init(ibk_ctx* ctx, Comm* c)
{
    CCK(cudaSetDevice(ctx->device));
    CCK(cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking));
    for (int t = 0; t < 2; ++t)
    {
        CCK(cudaEventCreateWithFlags(&c->ev_packed[t], cudaEventDisableTiming));
        CCK(cudaEventCreateWithFlags(&c->ev_arrived[t], cudaEventDisableTiming));
    }
    return IBK_OK;
}
} // namespace

extern "C" int ibk_comm_destroy(ibk_ctx* ctx);

extern "C" int ibk_comm_unique_id(void* id128)
{
    if (!id128) return IBK_ERR_INVALID;
    NcclApi& n = nccl();
    if (!n.error.empty()) return IBK_ERR_STATE;
    return n.GetUniqueId(id128) == 0 ? IBK_OK : IBK_ERR_CUDA;
}

extern "C" int ibk_comm_init(ibk_ctx* ctx, const void* id128, int rank, int nranks)
{
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return IBK_ERR_INVALID;
    NcclApi& n = nccl();
    if (!n.error.empty()) return cfail(ctx, IBK_ERR_STATE, n.error);
    ibk_comm_destroy(ctx);
    Comm* c = new Comm();
    c->rank = rank;
    c->nranks = nranks;
    c->transport = 1;
    comms()[ctx] = c;
    if (int rc = comm_common_init(ctx, c)) return rc;
    NcclApi::Uid id;
    std::memcpy(id.internal, id128, 128);
    NCK(n.CommInitRank(&c->nccl_comm, nranks, id, rank));
    {
        // the NCCL kernels of an exchange need somewhere to run while the spread's persistent CTAs occupy the SMs
        // (IBK_COMM_RESERVE_SMS overrides; measured on 2 GPUs: see DESIGN.md)
        const char* env = getenv("IBK_COMM_RESERVE_SMS");
        ctx->L.reserve_sms = env ? atoi(env) : 8;
    }
    CCK(cudaMalloc(&c->d_counts, sizeof(double) * (size_t)nranks * nranks));
    return IBK_OK;
}

// Several contexts of ONE process form the ranks 0..nranks-1 (on the same or on different devices): messages are moved
// by device copies on the receiver's stream, ordered by events.  Collective calls (ibk_halo_*_post on every rank before
// any *_finish) are the caller's responsibility, as with any communicator.
extern "C" int ibk_comm_init_loopback(ibk_ctx** ctxs, int nranks)
{
    if (!ctxs || nranks < 1) return IBK_ERR_INVALID;
    for (int r = 0; r < nranks; ++r)
        if (!ctxs[r]) return IBK_ERR_INVALID;
    for (int r = 0; r < nranks; ++r)
    {
        ibk_ctx* ctx = ctxs[r];
        ibk_comm_destroy(ctx);
        Comm* c = new Comm();
        c->rank = r;
        c->nranks = nranks;
        c->transport = 2;
        c->peers.assign(ctxs, ctxs + nranks);
        comms()[ctx] = c;
        if (int rc = comm_common_init(ctx, c)) return rc;
    }
    return IBK_OK;
}

extern "C" int ibk_comm_destroy(ibk_ctx* ctx)
{
    Comm* c = comm_of(ctx);
    if (!c) return IBK_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->L.stream);
    if (c->s_comm) cudaStreamSynchronize(c->s_comm);
    free_messages(c);
    if (c->nccl_comm && nccl().CommDestroy) nccl().CommDestroy(c->nccl_comm);
    if (c->d_counts) cudaFree(c->d_counts);
    for (int t = 0; t < 2; ++t)
    {
        if (c->ev_packed[t]) cudaEventDestroy(c->ev_packed[t]);
        if (c->ev_arrived[t]) cudaEventDestroy(c->ev_arrived[t]);
    }
    if (c->s_comm) cudaStreamDestroy(c->s_comm);
    delete c;
    comms().erase(ctx);
    return IBK_OK;
}

// The global patch list of the level (every rank passes the same): builds the plan for this rank from the geometry of
// the context's level and allocates the message buffers.  The k-th patch of `my rank` in the list must be the k-th
// patch the level was created with.
extern "C" int ibk_comm_set_patches(ibk_ctx* ctx, int n_patches, const int* lower, const int* upper, const int* rank)
{
    Comm* c = comm_of(ctx);
    if (!c) return cfail(ctx, IBK_ERR_STATE, "no communicator: call ibk_comm_init first");
    if (!ctx->lv.valid) return cfail(ctx, IBK_ERR_STATE, "no level: call ibk_level_create first");
    const ibk::LevelState& lv = ctx->lv;
    const int ndim = lv.ndim;
    if (n_patches <= 0 || !lower || !upper || !rank) return cfail(ctx, IBK_ERR_INVALID, "patch list missing");
    int k = 0;
    for (int p = 0; p < n_patches; ++p)
    {
        if (rank[p] < 0 || rank[p] >= c->nranks) return cfail(ctx, IBK_ERR_INVALID, "patch rank out of range");
        if (rank[p] != c->rank) continue;
        if (k >= (int)lv.patches.size()) return cfail(ctx, IBK_ERR_INVALID, "the list holds more patches of this rank than the level");
        for (int d = 0; d < ndim; ++d)
            if (lower[p * ndim + d] != lv.patches[k].lower[d] || upper[p * ndim + d] != lv.patches[k].upper[d])
                return cfail(ctx, IBK_ERR_INVALID, "the patches of this rank in the list differ from the level's (box or order)");
        ++k;
    }
    if (k != (int)lv.patches.size()) return cfail(ctx, IBK_ERR_INVALID, "the list holds fewer patches of this rank than the level");
    CCK(cudaSetDevice(ctx->device));
    cudaStreamSynchronize(ctx->L.stream);
    free_messages(c);
    int ncells[3] = { 1, 1, 1 };
    for (int d = 0; d < ndim; ++d) ncells[d] = lv.domain_upper[d] - lv.domain_lower[d] + 1;
    if (int rc = ibk_halo_plan_create(ndim, n_patches, lower, upper, rank, ncells, lv.periodic, lv.gcw, c->rank, &c->plan))
        return cfail(ctx, rc, "ibk_halo_plan_create failed");
    c->g_lower.assign(lower, lower + (size_t)n_patches * ndim);
    c->g_upper.assign(upper, upper + (size_t)n_patches * ndim);
    c->g_rank.assign(rank, rank + n_patches);
    for (int t = 0; t < 2; ++t)
        for (const ibk_halo_message& hm : c->plan->table[t])
        {
            const bool sending = hm.src_rank == c->rank;
            Message m;
            m.peer = sending ? hm.dst_rank : hm.src_rank;
            m.count = hm.count;
            long long off = 0;
            for (const ibk_halo_item& it : hm.items)
            {
                m.patch.push_back(c->plan->local_id[sending ? it.src : it.dst]);
                m.axis.push_back(it.axis);
                for (int d = 0; d < ndim; ++d)
                {
                    m.lo.push_back(sending ? it.src_lo[d] : it.dst_lo[d]);
                    m.hi.push_back(sending ? it.src_hi[d] : it.dst_hi[d]);
                }
                m.offs.push_back(off);
                off += it.count;
            }
            if (c->transport == 2)
            {
                CCK(cudaEventCreateWithFlags(&m.ev_ready, cudaEventDisableTiming));
                CCK(cudaEventCreateWithFlags(&m.ev_consumed, cudaEventDisableTiming));
            }
            (sending ? c->send[t] : c->recv[t]).push_back(m);
        }
    for (int t = 0; t < 2; ++t)
        for (int dir = 0; dir < 2; ++dir)
        {
            std::vector<Message>& v = dir == 0 ? c->send[t] : c->recv[t];
            long long total = 0;
            for (Message& m : v) total += (m.count + 15) & ~15ll; // (slices start on 128-byte boundaries)
            CCK(cudaMalloc(&c->all_buf[t][dir], sizeof(double) * (size_t)std::max<long long>(total, 1)));
            Comm::Merged& mg = c->merged[t][dir];
            long long base = 0;
            for (Message& m : v) // (receive side: ascending source rank = the order of the additions)
            {
                m.d_buf = c->all_buf[t][dir] + base;
                mg.patch.insert(mg.patch.end(), m.patch.begin(), m.patch.end());
                mg.axis.insert(mg.axis.end(), m.axis.begin(), m.axis.end());
                mg.lo.insert(mg.lo.end(), m.lo.begin(), m.lo.end());
                mg.hi.insert(mg.hi.end(), m.hi.begin(), m.hi.end());
                for (long long o : m.offs) mg.offs.push_back(base + o);
                base += (m.count + 15) & ~15ll;
            }
        }
    return IBK_OK;
}

namespace
{
// which: 0 = u (fill), 1 = f (accumulate); the table index is the same number
int halo_post(ibk_ctx* ctx, int t)
{
    Comm* c = comm_of(ctx);
    if (!c || !c->plan) return cfail(ctx, IBK_ERR_STATE, "no halo plan: call ibk_comm_init and ibk_comm_set_patches first");
    if (c->posted[t]) return cfail(ctx, IBK_ERR_STATE, "the previous exchange of this kind was posted but not finished");
    CCK(cudaSetDevice(ctx->device));
    if (t == 1) // what lies beyond a physical boundary is folded back before any ghost value leaves
        if (int rc = ibk_spread_fold_walls(ctx)) return rc;
    for (Message& m : c->send[t])
        if (c->transport == 2 && m.consumed_pending) // the peer must have copied the previous content
        {
            CCK(cudaStreamWaitEvent(ctx->L.stream, m.ev_consumed, 0));
            m.consumed_pending = false;
        }
    {
        const Comm::Merged& mg = c->merged[t][0]; // every message of the exchange in one launch
        if (!mg.patch.empty())
            if (int rc = ibk_halo_pack_many(ctx, t, (int)mg.patch.size(), mg.patch.data(), mg.axis.data(), mg.lo.data(), mg.hi.data(),
                                            mg.offs.data(), c->all_buf[t][0]))
                return rc;
    }
    if (c->transport == 2)
        for (Message& m : c->send[t]) CCK(cudaEventRecord(m.ev_ready, ctx->L.stream));
    if (c->transport == 1)
    {
        // the messages start when the packing is done and run on the communication stream
        NcclApi& n = nccl();
        CCK(cudaEventRecord(c->ev_packed[t], ctx->L.stream));
        CCK(cudaStreamWaitEvent(c->s_comm, c->ev_packed[t], 0));
        NCK(n.GroupStart());
        for (Message& m : c->recv[t]) NCK(n.Recv(m.d_buf, (size_t)m.count, NCCL_FLOAT64, m.peer, c->nccl_comm, c->s_comm));
        for (Message& m : c->send[t]) NCK(n.Send(m.d_buf, (size_t)m.count, NCCL_FLOAT64, m.peer, c->nccl_comm, c->s_comm));
        NCK(n.GroupEnd());
        CCK(cudaEventRecord(c->ev_arrived[t], c->s_comm));
    }
    c->posted[t] = true;
    c->n_posted[t]++;
    return IBK_OK;
}
int halo_finish(ibk_ctx* ctx, int t)
{
    Comm* c = comm_of(ctx);
    if (!c || !c->plan) return cfail(ctx, IBK_ERR_STATE, "no halo plan: call ibk_comm_init and ibk_comm_set_patches first");
    if (!c->posted[t]) return cfail(ctx, IBK_ERR_STATE, "finish without post");
    CCK(cudaSetDevice(ctx->device));
    if (c->transport == 1) CCK(cudaStreamWaitEvent(ctx->L.stream, c->ev_arrived[t], 0));
    for (Message& m : c->recv[t]) // ascending source rank: fixed order of the additions
    {
        if (c->transport == 2)
        {
            ibk_ctx* pctx = c->peers[m.peer];
            Comm* pc = comm_of(pctx);
            if (!pc || !pc->plan) return cfail(ctx, IBK_ERR_STATE, "loopback peer has no plan");
            Message* src = nullptr;
            for (Message& s : pc->send[t])
                if (s.peer == c->rank) src = &s;
            if (!src || src->count != m.count) return cfail(ctx, IBK_ERR_STATE, "loopback: the peer's plan does not match");
            if (pc->n_posted[t] < c->n_posted[t]) return cfail(ctx, IBK_ERR_STATE, "loopback: every rank must post before any rank finishes");
            CCK(cudaStreamWaitEvent(ctx->L.stream, src->ev_ready, 0));
            CCK(cudaMemcpyPeerAsync(m.d_buf, ctx->device, src->d_buf, pctx->device, sizeof(double) * (size_t)m.count, ctx->L.stream));
            CCK(cudaEventRecord(src->ev_consumed, ctx->L.stream));
            src->consumed_pending = true;
        }
    }
    {
        const Comm::Merged& mg = c->merged[t][1]; // all received messages in one launch, items in ascending source-rank order
        if (!mg.patch.empty())
            if (int rc = ibk_halo_unpack_many(ctx, t, (int)mg.patch.size(), mg.patch.data(), mg.axis.data(), mg.lo.data(), mg.hi.data(),
                                              mg.offs.data(), c->all_buf[t][1], t == 0 ? 0 : 1))
                return rc;
    }
    c->posted[t] = false;
    return IBK_OK;
}
} // namespace

extern "C" int ibk_halo_fill_post(ibk_ctx* ctx)
{
    return ctx ? halo_post(ctx, 0) : IBK_ERR_INVALID;
}
extern "C" int ibk_halo_fill_finish(ibk_ctx* ctx)
{
    return ctx ? halo_finish(ctx, 0) : IBK_ERR_INVALID;
}
extern "C" int ibk_halo_accumulate_post(ibk_ctx* ctx)
{
    return ctx ? halo_post(ctx, 1) : IBK_ERR_INVALID;
}
extern "C" int ibk_halo_accumulate_finish(ibk_ctx* ctx)
{
    return ctx ? halo_finish(ctx, 1) : IBK_ERR_INVALID;
}
// SMs the persistent spread kernel leaves free for the message kernels (default with the NCCL transport: 8, for the
// sequence in which the accumulate messages travel WHILE the interior tiles are spread; a sequence that does not overlap
// messages with the spread, e.g. bench.py's, sets 0)
extern "C" int ibk_comm_set_reserved_sms(ibk_ctx* ctx, int n_sms)
{
    if (!ctx) return IBK_ERR_INVALID;
    if (n_sms < 0 || n_sms > 64) return cfail(ctx, IBK_ERR_INVALID, "reserved SMs must be in 0..64");
    ctx->L.reserve_sms = n_sms;
    return IBK_OK;
}
extern "C" long long ibk_halo_bytes(ibk_ctx* ctx, int which)
{
    Comm* c = comm_of(ctx);
    if (!c || which < 0 || which > 1) return -1;
    long long n = 0;
    for (const Message& m : c->send[which]) n += m.count;
    return 8 * n;
}

// Marker migration over the communicator (NCCL transport): plan on the device, counts by an all-gather, rows by one
// group of sends and receives, unpack; ibk_rebin must precede and follow (see ibk_migrate_plan).
extern "C" int ibk_migrate(ibk_ctx* ctx, unsigned id_bound, int* n_sent, int* n_received)
{
    Comm* c = comm_of(ctx);
    if (!c || c->g_rank.empty()) return cfail(ctx, IBK_ERR_STATE, "no communicator / patch list: call ibk_comm_init and ibk_comm_set_patches first");
    if (c->transport != 1) return cfail(ctx, IBK_ERR_STATE, "ibk_migrate needs the NCCL transport (loopback: ibk_migrate_loopback)");
    NcclApi& n = nccl();
    const int R = c->nranks, ndim = ctx->lv.ndim, width = 3 * ndim + 1;
    std::vector<int> send_counts(R, 0);
    if (int rc = ibk_migrate_plan(ctx, (int)c->g_rank.size(), c->g_lower.data(), c->g_upper.data(), c->g_rank.data(), R, c->rank, send_counts.data()))
        return rc;
    CCK(cudaSetDevice(ctx->device));
    // counts: row `rank` of an R x R matrix, all-gathered
    std::vector<double> row(R), all((size_t)R * R);
    for (int r = 0; r < R; ++r) row[r] = (double)send_counts[r];
    CCK(cudaMemcpyAsync(c->d_counts + (size_t)c->rank * R, row.data(), sizeof(double) * R, cudaMemcpyHostToDevice, ctx->L.stream));
    NCK(n.AllGather(c->d_counts + (size_t)c->rank * R, c->d_counts, (size_t)R, NCCL_FLOAT64, c->nccl_comm, ctx->L.stream));
    CCK(cudaMemcpyAsync(all.data(), c->d_counts, sizeof(double) * (size_t)R * R, cudaMemcpyDeviceToHost, ctx->L.stream));
    CCK(cudaStreamSynchronize(ctx->L.stream));
    std::vector<int> recv_counts(R, 0);
    long long ns = 0, nr = 0;
    for (int r = 0; r < R; ++r)
    {
        recv_counts[r] = r == c->rank ? 0 : (int)(all[(size_t)r * R + c->rank] + 0.5);
        if (r == c->rank) send_counts[r] = 0;
        ns += send_counts[r];
        nr += recv_counts[r];
    }
    ibk::DevBuf& sb = ctx->b_mig[8];
    ibk::DevBuf& rb = ctx->b_mig[9];
    CCK(sb.reserve(sizeof(double) * (size_t)std::max<long long>(ns, 1) * width));
    CCK(rb.reserve(sizeof(double) * (size_t)std::max<long long>(nr, 1) * width));
    if (int rc = ibk_migrate_pack(ctx, sb.as<double>())) return rc;
    NCK(n.GroupStart());
    long long so = 0, ro = 0;
    for (int r = 0; r < R; ++r)
    {
        if (recv_counts[r]) NCK(n.Recv(rb.as<double>() + ro * width, (size_t)recv_counts[r] * width, NCCL_FLOAT64, r, c->nccl_comm, ctx->L.stream));
        ro += recv_counts[r];
    }
    for (int r = 0; r < R; ++r)
    {
        if (send_counts[r]) NCK(n.Send(sb.as<double>() + so * width, (size_t)send_counts[r] * width, NCCL_FLOAT64, r, c->nccl_comm, ctx->L.stream));
        so += send_counts[r];
    }
    NCK(n.GroupEnd());
    if (int rc = ibk_migrate_unpack(ctx, rb.as<double>(), (int)nr, id_bound)) return rc;
    if (n_sent) *n_sent = (int)ns;
    if (n_received) *n_received = (int)nr;
    return IBK_OK;
}

// The same for the contexts of a loopback communicator, all ranks in one call.
extern "C" int ibk_migrate_loopback(ibk_ctx** ctxs, int nranks, unsigned id_bound, int* n_moved)
{
    if (!ctxs || nranks < 1) return IBK_ERR_INVALID;
    std::vector<std::vector<int>> counts(nranks, std::vector<int>(nranks, 0));
    std::vector<Comm*> cs(nranks);
    for (int r = 0; r < nranks; ++r)
    {
        ibk_ctx* ctx = ctxs[r];
        cs[r] = comm_of(ctx);
        if (!cs[r] || cs[r]->transport != 2 || cs[r]->g_rank.empty()) return cfail(ctx, IBK_ERR_STATE, "not a loopback communicator with a patch list");
        if (int rc = ibk_migrate_plan(ctx, (int)cs[r]->g_rank.size(), cs[r]->g_lower.data(), cs[r]->g_upper.data(), cs[r]->g_rank.data(), nranks, r,
                                      counts[r].data()))
            return rc;
        counts[r][r] = 0;
    }
    int moved = 0;
    std::vector<std::vector<long long>> soff(nranks, std::vector<long long>(nranks + 1, 0));
    for (int r = 0; r < nranks; ++r)
    {
        ibk_ctx* ctx = ctxs[r];
        const int width = 3 * ctx->lv.ndim + 1;
        for (int q = 0; q < nranks; ++q) soff[r][q + 1] = soff[r][q] + counts[r][q];
        CCK(cudaSetDevice(ctx->device));
        CCK(ctx->b_mig[8].reserve(sizeof(double) * (size_t)std::max<long long>(soff[r][nranks], 1) * width));
        if (int rc = ibk_migrate_pack(ctx, ctx->b_mig[8].as<double>())) return rc;
        CCK(cudaStreamSynchronize(ctx->L.stream));
        moved += (int)soff[r][nranks];
    }
    for (int r = 0; r < nranks; ++r)
    {
        ibk_ctx* ctx = ctxs[r];
        const int width = 3 * ctx->lv.ndim + 1;
        long long nr = 0;
        for (int q = 0; q < nranks; ++q) nr += counts[q][r];
        CCK(cudaSetDevice(ctx->device));
        CCK(ctx->b_mig[9].reserve(sizeof(double) * (size_t)std::max<long long>(nr, 1) * width));
        long long ro = 0;
        for (int q = 0; q < nranks; ++q) // ascending source rank
        {
            if (!counts[q][r]) continue;
            CCK(cudaMemcpyPeerAsync(ctx->b_mig[9].as<double>() + ro * width, ctx->device, ctxs[q]->b_mig[8].as<double>() + soff[q][r] * width,
                                    ctxs[q]->device, sizeof(double) * (size_t)counts[q][r] * width, ctx->L.stream));
            ro += counts[q][r];
        }
        if (int rc = ibk_migrate_unpack(ctx, ctx->b_mig[9].as<double>(), (int)nr, id_bound)) return rc;
        CCK(cudaStreamSynchronize(ctx->L.stream));
    }
    if (n_moved) *n_moved = moved;
    return IBK_OK;
}
