// Test driver for the C++ host mirror (ibamr_b200/host): reads one case from a binary file, runs it
// through IBTK_B200::LEInteractor (seam B3) and IBAMR_B200::IBMethodB200 (seam B1), writes the results.
//   driver --static                 : LEInteractor static queries + "no GPU -> refuse" check, prints lines
//   driver --structure base ndim    : IBStandardInitializerB200 on <base>.vertex/.spring/.beam/.target/.anchor (no GPU)
//   driver case.bin out.bin         : GPU run
// case.bin: int32 n (cells per dim), int32 g (ghost width), int32 N (markers), char[32] kernel,
//           double X[N][3], double F[N][3], then per axis the side array of u (Fortran order, ghosts incl.)
// out.bin : double Q[N][3] (LEInteractor::interpolate), per-axis f arrays (LEInteractor::spread into zero),
//           double U[N][3] (IBMethodB200::interpolateVelocity), per-axis f arrays (IBMethodB200::spreadForce),
//           double Fl[N][3] (computeLagrangianForce for a closed ring of springs i -> i+1, kappa 1.5, rest 0.01, and
//           target points on every 7th marker), double Xnew[N][3] (forwardEulerStep with dt = 0.01), double ok (1.0 if the
//           LDataB200 host mirror of the F column behaved: read, modify + restore, refetch after a kernel)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <iostream>
#include <vector>

#define NDIM 3
#include "../../ibamr_b200/host/IBMethodB200.h"
#include "../../ibamr_b200/host/IBStandardInitializerB200.h"
#include "../../ibamr_b200/host/LDataB200.h"
#include "../../ibamr_b200/host/LDataManagerB200.h"
#include "../../ibamr_b200/host/LEInteractorB200.h"

using namespace SAMRAI_standin;
using IBTK_B200::LEInteractor;

static int run_static()
{
    const char* names[] = { "IB_4", "IB_6", "BSPLINE_3", "BSPLINE_4", "PIECEWISE_LINEAR", "IB_3", "BSPLINE_5", "BSPLINE_6", "PIECEWISE_CUBIC", "IB_5", "PIECEWISE_CONSTANT",
                            "COMPOSITE_BSPLINE_32", "DISCONTINUOUS_LINEAR", "IB_4_W8" };
    for (const char* n : names)
        std::printf("%s known=%d stencil=%d ghosts=%d\n", n, (int)LEInteractor::isKnownKernel(n), LEInteractor::getStencilSize(n),
                    LEInteractor::getMinimumGhostWidth(n));
    std::printf("IB_7 known=%d\n", (int)LEInteractor::isKnownKernel("IB_7"));
    try
    {
        LEInteractor::getStencilSize("IB_7");
        std::printf("unknown kernel: no error\n");
    }
    catch (const std::exception& e)
    {
        std::printf("unknown kernel: error raised\n");
    }
    ibk_ctx* c = nullptr;
    const int rc = ibk_ctx_create(0, &c);
    std::printf("ctx_create rc=%d\n", rc);
    if (c) ibk_ctx_destroy(c);
    return 0;
}

static int run_structure(const char* base, int ndim)
{
    try
    {
        IBAMR_B200::IBStandardInitializerB200 init(ndim, { std::string(base) });
        std::printf("vertices=%d springs=%zu beams=%zu targets=%zu anchors=%zu\n", init.num_vertices, init.spring_master.size(),
                    init.beam_curr.size(), init.target_idx.size(), init.anchor_idx.size());
        std::printf("X0=%.17g,%.17g\n", init.X[0], init.X[1]);
        if (!init.spring_master.empty())
            std::printf("last spring=%d,%d,%.17g,%.17g\n", init.spring_master.back(), init.spring_slave.back(), init.spring_kappa.back(),
                        init.spring_rest.back());
    }
    catch (const std::exception& e)
    {
        std::printf("error: %s\n", e.what());
        return 1;
    }
    return 0;
}

// Two ranks in one process (loopback communicator of libibk.so): the case's periodic n^3 level cut in two along x, each
// half a rank with its own IBMethodB200; the reference's signatures all the way (IBStrategy::spreadForce /
// interpolateVelocity with data indices bound to host SideData; rank 1 goes through LDataManagerB200::spread / interp).
// out: per rank  int64 count, int64 ids[count], double U[count][3], then the three f arrays of the rank's patch.
static int run_two_ranks(const char* case_file, const char* out_file)
{
    FILE* fi = std::fopen(case_file, "rb");
    if (!fi) return 3;
    int n, g, N;
    char kernel[32];
    if (std::fread(&n, 4, 1, fi) != 1 || std::fread(&g, 4, 1, fi) != 1 || std::fread(&N, 4, 1, fi) != 1 || std::fread(kernel, 1, 32, fi) != 32)
        return 4;
    std::vector<double> X((size_t)N * 3), F((size_t)N * 3);
    if (std::fread(X.data(), 8, X.size(), fi) != X.size() || std::fread(F.data(), 8, F.size(), fi) != F.size()) return 4;
    Box dom(Index(0), Index(n - 1));
    SideData ug(dom, 1, IntVector(g)); // the global field, ghosts analytic (periodic)
    for (int a = 0; a < 3; ++a)
        if (std::fread(ug.getPointer(a), 8, ug.data[a].size(), fi) != ug.data[a].size()) return 4;
    std::fclose(fi);
    FILE* fo = std::fopen(out_file, "wb");
    try
    {
        std::vector<Box> boxes(2, dom);
        boxes[0].hi(0) = n / 2 - 1;
        boxes[1].lo(0) = n / 2;
        std::vector<std::unique_ptr<IBAMR_B200::IBMethodB200>> ib;
        std::vector<std::unique_ptr<SideData>> u, f;
        std::vector<std::vector<long long>> ids(2);
        for (int r = 0; r < 2; ++r)
        {
            IBAMR_B200::IBMethodB200::LevelSpec lv;
            lv.domain_box = dom;
            for (int d = 0; d < 3; ++d)
            {
                lv.x_lower[d] = 0.0;
                lv.x_upper[d] = 1.0;
                lv.periodic[d] = 1;
            }
            lv.patch_boxes.push_back(boxes[r]);
            ib.emplace_back(new IBAMR_B200::IBMethodB200(lv, kernel, 0, g));
            u.emplace_back(new SideData(boxes[r], 1, IntVector(g)));
            f.emplace_back(new SideData(boxes[r], 1, IntVector(g)));
            f[r]->fillAll(0.25);
            // the rank's slice of the global field; its ghost values are poisoned: they must come from the exchange
            for (int a = 0; a < 3; ++a)
            {
                int ng[3], nl[3];
                for (int d = 0; d < 3; ++d)
                {
                    ng[d] = n + (d == a ? 1 : 0) + 2 * g;
                    nl[d] = boxes[r].hi(d) - boxes[r].lo(d) + 1 + (d == a ? 1 : 0) + 2 * g;
                }
                for (int k = 0; k < nl[2]; ++k)
                    for (int j = 0; j < nl[1]; ++j)
                        for (int i = 0; i < nl[0]; ++i)
                        {
                            const bool ghost = i < g || i >= nl[0] - g || j < g || j >= nl[1] - g || k < g || k >= nl[2] - g;
                            const size_t gi = ((size_t)(k + boxes[r].lo(2)) * ng[1] + (j + boxes[r].lo(1))) * ng[0] + (i + boxes[r].lo(0));
                            u[r]->getPointer(a)[((size_t)k * nl[1] + j) * nl[0] + i] = ghost ? 1e30 : ug.getPointer(a)[gi];
                        }
            }
            ib[r]->registerPatchData(/*u_data_idx*/ 0, 0, u[r].get());
            ib[r]->registerPatchData(/*f_data_idx*/ 1, 0, f[r].get());
            // the markers whose cell lies in the rank's patch
            std::vector<double> Xr, Fr;
            for (int i = 0; i < N; ++i)
            {
                int c = (int)std::floor(X[3 * i] * n);
                c = c < 0 ? 0 : (c >= n ? n - 1 : c);
                if ((c < n / 2) != (r == 0)) continue;
                ids[r].push_back(i);
                for (int d = 0; d < 3; ++d)
                {
                    Xr.push_back(X[3 * i + d]);
                    Fr.push_back(F[3 * i + d]);
                }
            }
            ib[r]->setPositions(Xr);
            ib[r]->setForce(Fr);
            ib[r]->beginDataRedistribution();
            ib[r]->endDataRedistribution();
        }
        IBAMR_B200::IBMethodB200::initLoopbackCommunicator({ ib[0].get(), ib[1].get() });
        for (int r = 0; r < 2; ++r) ib[r]->setGlobalPatches(boxes, { 0, 1 });
        // rank 0: IBStrategy calls; rank 1: the same two halves reached through the LDataManager-shaped adapter's columns
        for (int r = 0; r < 2; ++r) ib[r]->beginSpreadForce(1);
        for (int r = 0; r < 2; ++r) ib[r]->finishSpreadForce(1);
        for (int r = 0; r < 2; ++r) ib[r]->beginInterpolateVelocity(0);
        for (int r = 0; r < 2; ++r) ib[r]->finishInterpolateVelocity();
        for (int r = 0; r < 2; ++r)
        {
            std::vector<double> U;
            ib[r]->getVelocity(U);
            const long long cnt = (long long)ids[r].size();
            std::fwrite(&cnt, 8, 1, fo);
            std::fwrite(ids[r].data(), 8, ids[r].size(), fo);
            std::fwrite(U.data(), 8, U.size(), fo);
            for (int a = 0; a < 3; ++a) std::fwrite(f[r]->getPointer(a), 8, f[r]->data[a].size(), fo);
        }
        // seam B2 on one rank alone (no exchange involved): LDataManager::spread with node weights, then interp into an aux LData
        {
            IBAMR_B200::IBMethodB200::LevelSpec lv;
            lv.domain_box = dom;
            for (int d = 0; d < 3; ++d)
            {
                lv.x_lower[d] = 0.0;
                lv.x_upper[d] = 1.0;
                lv.periodic[d] = 1;
            }
            lv.patch_boxes.push_back(dom);
            IBAMR_B200::IBMethodB200 one(lv, kernel, 0, g);
            one.setPositions(X);
            one.setForce(F);
            one.beginDataRedistribution();
            SideData fh(dom, 1, IntVector(g));
            one.registerPatchData(0, 0, &ug);
            one.registerPatchData(1, 0, &fh);
            IBTK_B200::LDataManagerB200 mgr(one);
            std::vector<Pointer<IBTK_B200::LDataB200>> Fd{ std::make_shared<IBTK_B200::LDataB200>("F", one.ctx(), IBK_COL_F, 3) };
            std::vector<Pointer<IBTK_B200::LDataB200>> Xd{ std::make_shared<IBTK_B200::LDataB200>("X", one.ctx(), IBK_COL_X, 3) };
            std::vector<Pointer<IBTK_B200::LDataB200>> Ud{ std::make_shared<IBTK_B200::LDataB200>("U_aux", one.ctx(), IBK_COL_AUX, 3) };
            std::vector<double> ds(N);
            for (int i = 0; i < N; ++i) ds[i] = 0.5 + 0.001 * (i % 100);
            mgr.spread(1, Fd, Xd, ds, kernel, nullptr, {}, 0.0);
            for (int a = 0; a < 3; ++a) std::fwrite(fh.getPointer(a), 8, fh.data[a].size(), fo);
            mgr.interp(0, Ud, Xd, {}, {}, 0.0);
            const double* Ua = static_cast<const IBTK_B200::LDataB200&>(*Ud[0]).getLocalFormVecArray();
            std::fwrite(Ua, 8, (size_t)N * 3, fo);
        }
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "driver: %s\n", e.what());
        std::fclose(fo);
        return 5;
    }
    std::fclose(fo);
    return 0;
}

// LData restart round trip (LData.cpp:99-130, 186-209): a level with X, U, F on the device is written to a database file,
// a SECOND level is rebuilt from the file alone; the columns must come back bit for bit and the next spread must produce
// the very same f (the spread is deterministic).  Prints "restart ok" on success.
static int run_restart(const char* tmp_path)
{
    try
    {
        const int n = 32, N = 5000, g = 3;
        Box dom(Index(0), Index(n - 1));
        IBAMR_B200::IBMethodB200::LevelSpec lv;
        lv.domain_box = dom;
        for (int d = 0; d < 3; ++d)
        {
            lv.x_lower[d] = 0.0;
            lv.x_upper[d] = 1.0;
            lv.periodic[d] = 1;
        }
        lv.patch_boxes.push_back(dom);
        std::vector<double> X((size_t)N * 3), F((size_t)N * 3);
        unsigned long long sd = 12345;
        auto rnd = [&]() {
            sd = sd * 6364136223846793005ull + 1442695040888963407ull;
            return (double)(sd >> 11) * (1.0 / 9007199254740992.0);
        };
        for (auto& v : X) v = rnd();
        for (auto& v : F) v = 2.0 * rnd() - 1.0;
        std::vector<double> f_first[3], Xa, Ua, Fa;
        {
            IBAMR_B200::IBMethodB200 ib(lv, "IB_4", 0, g);
            ib.setPositions(X);
            ib.setForce(F);
            ib.beginDataRedistribution();
            SideData u(dom, 1, IntVector(g)), f(dom, 1, IntVector(g));
            for (int a = 0; a < 3; ++a)
                for (size_t k = 0; k < u.data[a].size(); ++k) u.data[a][k] = std::sin(0.001 * (double)k);
            ib.registerPatchData(0, 0, &u);
            ib.registerPatchData(1, 0, &f);
            ib.interpolateVelocity(0, {}, {}, 0.0);
            ib.spreadForce(1, nullptr, {}, 0.0);
            for (int a = 0; a < 3; ++a) f_first[a] = f.data[a];
            ib.getColumn(IBK_COL_X, Xa);
            ib.getColumn(IBK_COL_U, Ua);
            ib.getColumn(IBK_COL_F, Fa);
            const char* names[3] = { "X", "U", "F" };
            const int cols[3] = { IBK_COL_X, IBK_COL_U, IBK_COL_F };
            for (int c = 0; c < 3; ++c)
            {
                auto db = std::make_shared<Database>();
                IBTK_B200::LDataB200 data(names[c], ib.ctx(), cols[c], 3);
                data.putToDatabase(db);
                db->writeToFile(std::string(tmp_path) + "." + names[c]);
            }
        }
        {
            IBAMR_B200::IBMethodB200 ib(lv, "IB_4", 0, g);
            auto dbX = Database::readFromFile(std::string(tmp_path) + ".X");
            std::vector<double> Xr((size_t)dbX->getInteger("num_local_nodes") * 3);
            dbX->getDoubleArray("vals", Xr.data(), (int)Xr.size());
            ib.setPositions(Xr);
            IBTK_B200::LDataB200 Ud(Database::readFromFile(std::string(tmp_path) + ".U"), ib.ctx(), IBK_COL_U);
            IBTK_B200::LDataB200 Fd(Database::readFromFile(std::string(tmp_path) + ".F"), ib.ctx(), IBK_COL_F);
            ib.beginDataRedistribution();
            std::vector<double> Xb, Ub, Fb;
            ib.getColumn(IBK_COL_X, Xb);
            ib.getColumn(IBK_COL_U, Ub);
            ib.getColumn(IBK_COL_F, Fb);
            if (Xb != Xa || Ub != Ua || Fb != Fa || Ud.getName() != "U" || Fd.getDepth() != 3)
            {
                std::printf("restart: columns differ\n");
                return 1;
            }
            SideData f(dom, 1, IntVector(g));
            ib.registerPatchData(1, 0, &f);
            ib.spreadForce(1, nullptr, {}, 0.0);
            for (int a = 0; a < 3; ++a)
                if (f.data[a] != f_first[a])
                {
                    std::printf("restart: the spread after the restart differs\n");
                    return 1;
                }
        }
        std::printf("restart ok\n");
    }
    catch (const std::exception& e)
    {
        std::printf("restart error: %s\n", e.what());
        return 1;
    }
    return 0;
}

// N3 through the C++ seam: two levels resident on one device, LDataManagerB200 of the finer one with setCoarserLevel.
// spread: the coarse level's f is a constant, the fine markers carry no force, the prolongation schedule of level 1 is set
// -> the fine f must be that constant on every patch interior (a constant has no slope).  interp: the fine u is linear in x,
// the synchronisation schedule of level 1 is set -> the coarse u under the fine patch must be the same linear field (the
// area mean of a linear field is its value at the centre).
static int run_amr()
{
    try
    {
        const int nc = 16, ratio = 2, g = 3, N = 2000;
        IBAMR_B200::IBMethodB200::LevelSpec lc, lf;
        lc.domain_box = Box(Index(0), Index(nc - 1));
        lf.domain_box = Box(Index(0), Index(nc * ratio - 1));
        for (int d = 0; d < 3; ++d)
        {
            lc.x_lower[d] = lf.x_lower[d] = 0.0;
            lc.x_upper[d] = lf.x_upper[d] = 1.0;
            lc.periodic[d] = lf.periodic[d] = 1;
        }
        lc.patch_boxes.push_back(lc.domain_box);
        const Box fine_box(Index(8), Index(23));
        lf.patch_boxes.push_back(fine_box);
        IBAMR_B200::IBMethodB200 ibc(lc, "IB_4", 0, g), ibf(lf, "IB_4", 0, g);
        std::vector<double> X((size_t)N * 3), F((size_t)N * 3, 0.0);
        unsigned long long sd = 99;
        for (auto& v : X)
        {
            sd = sd * 6364136223846793005ull + 1442695040888963407ull;
            v = 0.4 + 0.2 * (double)(sd >> 11) * (1.0 / 9007199254740992.0); // well inside the fine patch
        }
        ibf.setPositions(X);
        ibf.setForce(F);
        ibf.beginDataRedistribution();
        const int r3[3] = { ratio, ratio, ratio };
        IBTK_B200::LDataManagerB200 mgr(ibf, 1);
        mgr.setCoarserLevel(&ibc, r3);
        std::vector<Pointer<IBTK_B200::LDataB200>> Xd(2), Fd(2), Ud(2);
        Xd[1] = std::make_shared<IBTK_B200::LDataB200>("X", ibf.ctx(), IBK_COL_X, 3);
        Fd[1] = std::make_shared<IBTK_B200::LDataB200>("F", ibf.ctx(), IBK_COL_F, 3);
        Ud[1] = std::make_shared<IBTK_B200::LDataB200>("U", ibf.ctx(), IBK_COL_U, 3);
        // ---- spread with prolongation
        if (ibk_grid_fill(ibc.ctx(), 1, 2.5) != IBK_OK || ibk_grid_fill(ibf.ctx(), 1, -1.0) != IBK_OK) return 1;
        std::vector<Pointer<RefineSchedule>> prolong(2);
        prolong[1] = std::make_shared<RefineSchedule>();
        mgr.spread(-1, Fd, Xd, "IB_4", nullptr, prolong, 0.0);
        const int nf = 16 + 2 * g;
        for (int a = 0; a < 3; ++a)
        {
            const int n0 = nf + (a == 0), n1 = nf + (a == 1), n2 = nf + (a == 2);
            std::vector<double> f((size_t)n0 * n1 * n2);
            if (ibk_grid_download(ibf.ctx(), 1, 0, a, f.data()) != IBK_OK) return 1;
            for (int k = g; k < n2 - g; ++k)
                for (int j = g; j < n1 - g; ++j)
                    for (int i = g; i < n0 - g; ++i)
                        if (f[((size_t)k * n1 + j) * n0 + i] != 2.5)
                        {
                            std::printf("amr: prolonged f is %.17g at (%d,%d,%d) of axis %d\n", f[((size_t)k * n1 + j) * n0 + i], i, j, k, a);
                            return 1;
                        }
        }
        // ---- interp with synchronisation
        const double hf = 1.0 / (nc * ratio), hc = 1.0 / nc;
        for (int a = 0; a < 3; ++a)
        {
            const int n0 = nf + (a == 0), n1 = nf + (a == 1), n2 = nf + (a == 2);
            std::vector<double> u((size_t)n0 * n1 * n2);
            for (int k = 0; k < n2; ++k)
                for (int j = 0; j < n1; ++j)
                    for (int i = 0; i < n0; ++i) u[((size_t)k * n1 + j) * n0 + i] = 1.0 + 2.0 * ((8 - g + i) + (a == 0 ? 0.0 : 0.5)) * hf;
            if (ibk_grid_upload(ibf.ctx(), 0, 0, a, u.data()) != IBK_OK) return 1;
        }
        if (ibk_grid_fill(ibc.ctx(), 0, 0.0) != IBK_OK) return 1;
        std::vector<Pointer<CoarsenSchedule>> synch(2);
        synch[1] = std::make_shared<CoarsenSchedule>();
        mgr.interp(-1, Ud, Xd, synch, {}, 0.0);
        const int ncg = nc + 2 * g;
        double worst = 0.0;
        long long covered = 0;
        for (int a = 0; a < 3; ++a)
        {
            const int n0 = ncg + (a == 0), n1 = ncg + (a == 1), n2 = ncg + (a == 2);
            std::vector<double> u((size_t)n0 * n1 * n2);
            if (ibk_grid_download(ibc.ctx(), 0, 0, a, u.data()) != IBK_OK) return 1;
            const int hi[3] = { 11 + (a == 0), 11 + (a == 1), 11 + (a == 2) }; // coarse sides tiled by the fine patch [8, 23]
            for (int k = 4; k <= hi[2]; ++k)
                for (int j = 4; j <= hi[1]; ++j)
                    for (int i = 4; i <= hi[0]; ++i)
                    {
                        const double want = 1.0 + 2.0 * (i + (a == 0 ? 0.0 : 0.5)) * hc;
                        worst = std::max(worst, std::fabs(u[((size_t)(k + g) * n1 + (j + g)) * n0 + (i + g)] - want));
                        ++covered;
                    }
        }
        if (worst > 1e-13 || covered == 0)
        {
            std::printf("amr: coarsened u off by %.3e\n", worst);
            return 1;
        }
        std::printf("amr ok\n");
    }
    catch (const std::exception& e)
    {
        std::printf("amr error: %s\n", e.what());
        return 1;
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 2 && !std::strcmp(argv[1], "--amr")) return run_amr();
    if (argc >= 3 && !std::strcmp(argv[1], "--restart")) return run_restart(argv[2]);
    if (argc >= 2 && !std::strcmp(argv[1], "--static")) return run_static();
    if (argc >= 4 && !std::strcmp(argv[1], "--two-ranks")) return run_two_ranks(argv[2], argv[3]);
    if (argc >= 4 && !std::strcmp(argv[1], "--structure")) return run_structure(argv[2], std::atoi(argv[3]));
    if (argc < 3) return 2;
    FILE* fi = std::fopen(argv[1], "rb");
    if (!fi) return 3;
    int n, g, N;
    char kernel[32];
    if (std::fread(&n, 4, 1, fi) != 1 || std::fread(&g, 4, 1, fi) != 1 || std::fread(&N, 4, 1, fi) != 1 ||
        std::fread(kernel, 1, 32, fi) != 32)
        return 4;
    std::vector<double> X((size_t)N * 3), F((size_t)N * 3);
    if (std::fread(X.data(), 8, X.size(), fi) != X.size() || std::fread(F.data(), 8, F.size(), fi) != F.size()) return 4;
    Box box(Index(0), Index(n - 1));
    auto u = std::make_shared<SideData>(box, 1, IntVector(g));
    for (int a = 0; a < 3; ++a)
        if (std::fread(u->getPointer(a), 8, u->data[a].size(), fi) != u->data[a].size()) return 4;
    std::fclose(fi);

    auto geom = std::make_shared<CartesianPatchGeometry>();
    for (int d = 0; d < 3; ++d)
    {
        geom->x_lower[d] = 0.0;
        geom->x_upper[d] = 1.0;
        geom->dx[d] = 1.0 / n;
    }
    auto patch = std::make_shared<Patch>();
    patch->box = box;
    patch->geom = geom;

    FILE* fo = std::fopen(argv[2], "wb");
    try
    {
        // seam B3
        std::vector<double> Q((size_t)N * 3, 0.0);
        LEInteractor::interpolate(Q, 3, X, 3, u, patch, box, kernel);
        std::fwrite(Q.data(), 8, Q.size(), fo);
        auto f = std::make_shared<SideData>(box, 1, IntVector(g));
        LEInteractor::spread(f, F, 3, X, 3, patch, box, kernel);
        for (int a = 0; a < 3; ++a) std::fwrite(f->getPointer(a), 8, f->data[a].size(), fo);
        // seam B1
        IBAMR_B200::IBMethodB200::LevelSpec lv;
        lv.domain_box = box;
        for (int d = 0; d < 3; ++d)
        {
            lv.x_lower[d] = 0.0;
            lv.x_upper[d] = 1.0;
            lv.periodic[d] = 1;
        }
        lv.patch_boxes.push_back(box);
        IBAMR_B200::IBMethodB200 ib(lv, kernel, 0, g);
        ib.setPositions(X);
        ib.setForce(F);
        ib.setEulerianVelocity(0, *u);
        ib.beginDataRedistribution();
        ib.endDataRedistribution();
        ib.interpolateVelocity();
        std::vector<double> U;
        ib.getVelocity(U);
        std::fwrite(U.data(), 8, U.size(), fo);
        ib.spreadForce();
        SideData f2(box, 1, IntVector(g));
        ib.getEulerianForce(0, f2);
        for (int a = 0; a < 3; ++a) std::fwrite(f2.getPointer(a), 8, f2.data[a].size(), fo);
        // N1: force generation and a forward-Euler step on the device
        std::vector<int> m(N), sl(N), ti;
        std::vector<double> kap(N, 1.5), rest(N, 0.01), tk, te, tx0;
        for (int i = 0; i < N; ++i)
        {
            m[i] = i;
            sl[i] = (i + 1) % N;
            if (i % 7 == 0)
            {
                ti.push_back(i);
                tk.push_back(2.0);
                te.push_back(0.25);
                for (int d = 0; d < 3; ++d) tx0.push_back(0.5);
            }
        }
        ib.registerSprings(m, sl, kap, rest);
        ib.registerTargetPoints(ti, tk, te, tx0);
        ib.computeLagrangianForce();
        std::vector<double> Fl, Xnew;
        ib.getColumn(IBK_COL_F, Fl);
        std::fwrite(Fl.data(), 8, Fl.size(), fo);
        ib.preprocessIntegrateData();
        ib.forwardEulerStep(0.0, 0.01);
        ib.getColumn(IBK_COL_X_NEW, Xnew);
        std::fwrite(Xnew.data(), 8, Xnew.size(), fo);
        // seam B2: LData-shaped lazy host mirror of the F column: read, modify on the host, restore, recompute
        IBTK_B200::LDataB200 Fdata("F", ib.ctx(), IBK_COL_F, 3);
        const IBTK_B200::LDataB200& Fconst = Fdata;
        const double* f_ro = Fconst.getLocalFormVecArray();
        int ok = 1;
        for (size_t k = 0; k < Fl.size(); ++k) ok &= (f_ro[k] == Fl[k]);
        ok &= Fdata.hostCopyIsCurrent() ? 1 : 0;
        double* f_rw = Fdata.getLocalFormVecArray();
        for (size_t k = 0; k < Fl.size(); ++k) f_rw[k] = 2.0 * f_rw[k];
        Fdata.restoreArrays();
        std::vector<double> F2;
        ib.getColumn(IBK_COL_F, F2);
        for (size_t k = 0; k < Fl.size(); ++k) ok &= (F2[k] == 2.0 * Fl[k]);
        ib.computeLagrangianForce(); // a kernel rewrites the column: the mirror must refetch
        Fdata.markDeviceModified();
        ok &= Fdata.hostCopyIsCurrent() ? 0 : 1;
        const double* f_again = Fconst.getLocalFormVecArray();
        std::vector<double> F3; // X is the midpoint position by now, so this is a new force, not Fl
        ib.getColumn(IBK_COL_F, F3);
        int changed = 0;
        for (size_t k = 0; k < Fl.size(); ++k)
        {
            ok &= (f_again[k] == F3[k]);
            changed |= (F3[k] != F2[k]);
        }
        ok &= changed;
        const double okd = (double)ok;
        std::fwrite(&okd, 8, 1, fo);
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "driver: %s\n", e.what());
        std::fclose(fo);
        return 5;
    }
    std::fclose(fo);
    return 0;
}
