// capi_host.cpp -- C entry points of libnsem_host.so for the Python test/bench harness and other FFI users.
// Wraps nsemh::EulerSolver (euler_app.cpp); every call returns 0 on success, the message is in nsemh_error().
#include <stdexcept>
#include <cstdio>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <chrono>
#include <cstdlib>

#include "nsem_host.h"

using namespace nsemh;

struct nsemh_solver {
    std::unique_ptr<EulerSolver> sp{new EulerSolver()};    // replaced by the solver on the regridded mesh at an AMR regrid
    std::string err;
};
static std::string g_err;

namespace {
struct HillArg { double H, xc, hw, Lz; };
void hill_map(Vec3& v, const void* a) {
    const HillArg* h = (const HillArg*)a;
    const double PI = 3.14159265358979323846264;
    const double x = v[0];
    double ht = 0;
    if (std::fabs(x - h->xc) < h->hw) {
        const double c = std::cos(0.5 * PI * (x - h->xc) / h->hw);
        ht = h->H * (c * c);
    }
    v[2] = ht + v[2] * (h->Lz - ht) / h->Lz;
}
FieldFile mkfield(int comps, const std::string& kind, std::vector<double> a, const std::vector<BCond>& bcs) {
    FieldFile f;
    f.comps = comps;
    f.inits.push_back({kind, std::move(a)});
    f.bcs = bcs;
    return f;
}
std::vector<BCond> all(const std::vector<std::string>& patches, const std::string& type) {
    std::vector<BCond> v;
    for (auto& p : patches) { BCond b; b.patch = p; b.type = type; v.push_back(b); }
    return v;
}
}  // namespace

extern "C" {

const char* nsemh_error(const nsemh_solver* h) { return h ? h->err.c_str() : g_err.c_str(); }

void nsemh_close(nsemh_solver* h) { delete h; }

// Full pre-loop sequence of the euler app on a case directory (controls, <mesh>_<step>, rho/U/T/p<step>).
nsemh_solver* nsemh_open_case(const char* dir, int step) {
    nsemh_solver* h = new nsemh_solver();
    try {
        (*h->sp).read_controls(dir);
        (*h->sp).load_mesh(step);
        (*h->sp).read_fields(step);
        (*h->sp).setup();
        return h;
    } catch (const std::exception& e) {
        g_err = e.what();
        delete h;
        return nullptr;
    }
}

// The same for partition `rank` of `nranks`: every rank decomposes the case's grid the way the controls say (decomposition{type, n}) and
// keeps its part, fields localized through the cell map (what lib/euler does under a multi-process launch).
nsemh_solver* nsemh_open_case_part(const char* dir, int step, int rank, int nranks) {
    nsemh_solver* h = new nsemh_solver();
    try {
        (*h->sp).rank = rank; (*h->sp).nranks = nranks;
        (*h->sp).read_controls(dir);
        (*h->sp).load_mesh(step);
        (*h->sp).read_fields(step);
        (*h->sp).setup();
        return h;
    } catch (const std::exception& e) {
        g_err = e.what();
        delete h;
        return nullptr;
    }
}

// In-memory synthetic cases (same definitions as oracle/cases.py): kind in {bubble2d, bubble3d, vortex, hill3d}.
// nranks > 1: the global grid is decomposed (decomp = METIS | XYZ | CELLID; XYZ uses px*py*pz = nranks) and only
// partition `rank` is kept (Prepare::decomposeMesh, field.cpp:1086-1257).
nsemh_solver* nsemh_synthetic_part(const char* kind_, int nx, int ny, int nz, int order, int rank, int nranks,
                                   const char* decomp, int px, int py, int pz) {
    nsemh_solver* h = new nsemh_solver();
    EulerSolver& s = (*h->sp);
    const std::string kind = kind_;
    const int pxyz[3] = {px, py, pz};
    const bool verbose = std::getenv("NSEM_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto t1 = std::chrono::steady_clock::now();
        if (verbose) std::printf("synthetic[%d]: %-30s %.3f s\n", rank, what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    auto mesh = [&](const Grid& g) {
        lap("box_grid");
        if (nranks > 1) s.set_mesh_partition(g, rank, nranks, decomp ? decomp : "METIS", pxyz);
        else s.set_mesh(g);
        lap("set_mesh");
    };
    try {
        s.time_scheme = "BDF1";
        s.viscosity = 1.5;
        if (kind == "bubble2d") {
            const int n[3] = {nx, 1, nz};
            const double lo[3] = {0, 0, 0}, hi[3] = {1000, 100, 1000};
            s.nop[0] = order; s.nop[1] = 0; s.nop[2] = order;
            s.dt = 0.005; s.gravity = Vec3{0, 0, -9.80606};
            mesh(box_grid(n, lo, hi, {"sides", "sides", "delete", "delete", "bottom", "top"}));
            const std::vector<std::string> pt = {"top", "bottom", "sides"};
            s.set_fields(mkfield(1, "uniform", {0}, all(pt, "NEUMANN")), mkfield(3, "uniform", {0, 0, 0}, all(pt, "SYMMETRY")),
                         mkfield(1, "cosine", {0, 0.5, 500, 50, 350, 250, 1000, 250}, all(pt, "NEUMANN")),
                         mkfield(1, "uniform", {0}, all(pt, "NEUMANN")));
        } else if (kind.rfind("bubble3d", 0) == 0) {
            // "bubble3d" = the 1000 m cube of examples/atmo/srtb-3d; "bubble3d:sx,sy,sz" stretches the domain (not the
            // elements) by integer factors, so that a weak-scaling run keeps the element size -- and with it the acoustic
            // Courant number of the explicit step -- of the one-partition case.
            // An optional fourth number replaces the example's dt = 0.00125, which belongs to ITS 6^3 (+AMR) mesh: the step is
            // one forward-Euler stage, so a finer mesh needs a proportionally smaller dt to stay at the example's Courant number.
            double sc[3] = {1, 1, 1}, dt_case = 0.00125;
            if (kind.size() > 8) {
                const int got = (kind[8] == ':') ? std::sscanf(kind.c_str() + 9, "%lf,%lf,%lf,%lf", &sc[0], &sc[1], &sc[2], &dt_case) : 0;
                if (got < 3 || !(dt_case > 0)) throw std::runtime_error("synthetic: expected bubble3d or bubble3d:sx,sy,sz[,dt]");
            }
            const int n[3] = {nx, ny, nz};
            const double lo[3] = {0, 0, 0}, hi[3] = {1000 * sc[0], 1000 * sc[1], 1000 * sc[2]};
            s.nop[0] = s.nop[1] = s.nop[2] = order;
            s.dt = dt_case; s.gravity = Vec3{0, -9.80606, 0}; s.time_scheme = "AB1";
            mesh(box_grid(n, lo, hi, {"sides", "sides", "bottom", "top", "sides", "sides"}));
            const std::vector<std::string> pt = {"top", "bottom", "sides"};
            s.set_fields(mkfield(1, "uniform", {0}, all(pt, "NEUMANN")), mkfield(3, "uniform", {0, 0, 0}, all(pt, "SYMMETRY")),
                         mkfield(1, "cosine", {0, 0.5, 500 * sc[0], 350, 500 * sc[2], 250, 250, 250}, all(pt, "NEUMANN")),
                         mkfield(1, "uniform", {0}, all(pt, "NEUMANN")));
        } else if (kind == "vortex") {
            const int n[3] = {nx, ny, 1};
            const double lo[3] = {-5, -5, -0.5}, hi[3] = {5, 5, 0.5};
            s.nop[0] = s.nop[1] = order; s.nop[2] = 0;
            s.T0 = 1; s.P0 = 1; s.cp = 3.5; s.cv = 2.5; s.viscosity = 0; s.dt = 0.0005;
            s.gravity = Vec3{0, -9.80606, 0}; s.diffusion = false; s.buoyancy = false; s.problem_init = "ISENTROPIC_VORTEX";
            s.cyclic_patches = {{"inx", "outx"}, {"iny", "outy"}};       // a decomposition keeps the owner cells of paired periodic faces together
            mesh(box_grid(n, lo, hi, {"inx", "outx", "iny", "outy", "delete", "delete"}));
            std::vector<BCond> cyc;
            const char* pr[4][2] = {{"inx", "outx"}, {"outx", "inx"}, {"iny", "outy"}, {"outy", "iny"}};
            for (auto& q : pr) { BCond b; b.patch = q[0]; b.type = "CYCLIC"; b.neighbor = q[1]; cyc.push_back(b); }
            s.set_fields(mkfield(1, "uniform", {0}, cyc), mkfield(3, "uniform", {1, 1, 0}, cyc), mkfield(1, "uniform", {0}, cyc),
                         mkfield(1, "uniform", {0}, cyc));
        } else if (kind.rfind("hill3d", 0) == 0) {
            // "hill3d" or "hill3d:dt" (a finer mesh needs a smaller forward-Euler step, see bubble3d)
            double dt_case = 0.001;
            if (kind.size() > 6 && (kind[6] != ':' || std::sscanf(kind.c_str() + 7, "%lf", &dt_case) != 1 || !(dt_case > 0)))
                throw std::runtime_error("synthetic: expected hill3d or hill3d:dt");
            const int n[3] = {nx, ny, nz};
            static HillArg ha{200.0, 1000.0, 400.0, 1400.0};
            const double lo[3] = {0, 0, 0}, hi[3] = {3400, 100.0 * ny, ha.Lz};
            s.nop[0] = s.nop[1] = s.nop[2] = order;
            s.dt = dt_case; s.gravity = Vec3{0, 0, -9.80606};
            mesh(box_grid(n, lo, hi, {"inlet", "outlet", "sides", "sides", "WALLS", "top"}, hill_map, &ha));
            const std::vector<std::string> pt = {"inlet", "outlet", "WALLS", "top", "sides"};
            std::vector<BCond> ub = all(pt, "SYMMETRY");
            ub[0].type = "DIRICHLET"; ub[0].value[0] = 10;
            ub[1].type = "NEUMANN";
            s.set_fields(mkfield(1, "uniform", {0}, all(pt, "NEUMANN")), mkfield(3, "uniform", {10, 0, 0}, ub),
                         mkfield(1, "cosine", {0, 0.5, 1700, 100, 700, 400, 1000, 300}, all(pt, "NEUMANN")),
                         mkfield(1, "uniform", {0}, all(pt, "NEUMANN")));
        } else {
            throw Error("unknown synthetic case " + kind);
        }
        lap("set_fields");
        s.setup();
        lap("setup");
        return h;
    } catch (const std::exception& e) {
        g_err = e.what();
        delete h;
        return nullptr;
    }
}

nsemh_solver* nsemh_synthetic(const char* kind, int nx, int ny, int nz, int order) {
    return nsemh_synthetic_part(kind, nx, ny, nz, order, 0, 1, "METIS", 1, 1, 1);
}

#define GUARD(...)                          \
    try { __VA_ARGS__; return 0; }          \
    catch (const std::exception& e) { h->err = e.what(); return 1; }

int nsemh_attach(nsemh_solver* h, int device, int rank, int nranks, const void* uid) { GUARD((*h->sp).attach_device(device, rank, nranks, uid)) }
int nsemh_step(nsemh_solver* h, int n) { GUARD((*h->sp).step(n)) }
int nsemh_upload(nsemh_solver* h) { GUARD((*h->sp).upload_state()) }
int nsemh_download(nsemh_solver* h) { GUARD((*h->sp).download()) }
int nsemh_upload_async(nsemh_solver* h) { GUARD((*h->sp).upload_state_async()) }
int nsemh_download_async(nsemh_solver* h) { GUARD((*h->sp).download_async()) }
// AMR field transfer on the device: solver `h` (regridded mesh) takes the state of `old` (EulerSolver::adopt_refined_state)
int nsemh_adopt_refined_state(nsemh_solver* h, nsemh_solver* old, const uint32_t* refineMap, uint32_t nr, const uint32_t* coarseMap, uint32_t nc,
                              const uint32_t* cellMap, uint32_t nm, int restart) {
    GUARD((*h->sp).adopt_refined_state((*old->sp), std::vector<u32>(refineMap, refineMap + nr), std::vector<u32>(coarseMap, coarseMap + nc),
                                   std::vector<u32>(cellMap, cellMap + nm), restart != 0))
}
// AMR regrid in memory (amr.cpp): the handle continues as the solver on the regridded mesh, state transferred on the device when attached.
// refine / coarsen: one flag per current cell; NULL, NULL = tag by the refinement{} indicator of the controls.
int nsemh_regrid(nsemh_solver* h, const uint8_t* refine, const uint8_t* coarsen, uint32_t n) {
    GUARD(std::unique_ptr<EulerSolver> nw = (refine && coarsen)
              ? (*h->sp).regridded(std::vector<uint8_t>(refine, refine + n), std::vector<uint8_t>(coarsen, coarsen + n))
              : (*h->sp).regridded_by_indicator();
          h->sp = std::move(nw))
}
int nsemh_enable_amr(nsemh_solver* h, double dx, double dy, double dz, const char* field, double fmin, double fmax, int max_level, int buffer_zone) {
    GUARD(EulerSolver& s = *h->sp;
          if (!s.forest) throw Error("nsemh_enable_amr: the solver has no AMR forest (create it with NSEM_AMR=1 or amr_step in the controls)");
          s.refine_params.dir = Vec3{dx, dy, dz}; s.forest->dir = s.refine_params.dir;
          if (field && *field) s.refine_params.field = field;
          s.refine_params.field_min = fmin; s.refine_params.field_max = fmax; s.refine_params.max_level = max_level;
          s.refine_params.buffer_zone = buffer_zone)
}
int nsemh_write_amr_grid(nsemh_solver* h, int dump) { GUARD((*h->sp).write_amr_grid(dump)) }
int nsemh_cell_levels(nsemh_solver* h, int32_t* out, uint32_t n) {
    GUARD(if (!(*h->sp).forest) throw Error("nsemh_cell_levels: no AMR forest");
          const std::vector<int> l = (*h->sp).forest->levels();
          if (l.size() != n) throw Error("nsemh_cell_levels: one entry per real cell expected");
          std::copy(l.begin(), l.end(), out))
}
int nsemh_restart_state(nsemh_solver* h) { GUARD((*h->sp).restart_state()) }
int nsemh_write(nsemh_solver* h, int index) { GUARD((*h->sp).write_fields(index)) }
int nsemh_write_vtk(nsemh_solver* h, int index) { GUARD((*h->sp).write_vtk(index)) }
int nsemh_run(nsemh_solver* h) { GUARD(run_case(h->sp)) }
int nsemh_sync(nsemh_solver* h) { GUARD(if (nsem_sync((*h->sp).ctx)) throw Error(nsem_last_error((*h->sp).ctx))) }
int nsemh_time(nsemh_solver* h, int nsteps, double* ms, double* per_kernel) {
    GUARD(if (!(*h->sp).ctx) throw Error("no device attached"); if (nsem_time_steps((*h->sp).ctx, nsteps, ms, per_kernel)) throw Error(nsem_last_error((*h->sp).ctx)))
}
// {courant max, min, avg, mass, energy, volume} of the CURRENT device state + the initial totals mass0/energy0/volume0
int nsemh_diagnostics(nsemh_solver* h, double out[9]) {
    GUARD(if (!(*h->sp).ctx) throw Error("no device attached"); if (nsem_diagnostics((*h->sp).ctx, out)) throw Error(nsem_last_error((*h->sp).ctx));
          out[6] = (*h->sp).mass0; out[7] = (*h->sp).energy0; out[8] = (*h->sp).volume0)
}
uint64_t nsemh_launch_count(nsemh_solver* h) { return (*h->sp).ctx ? nsem_launch_count((*h->sp).ctx) : 0; }
const char* nsemh_kernel_info(nsemh_solver* h) { return (*h->sp).ctx ? nsem_kernel_info((*h->sp).ctx) : ""; }
const char* nsemh_halo_info(nsemh_solver* h) { return (*h->sp).ctx ? nsem_halo_info((*h->sp).ctx) : ""; }
int nsemh_halo_wait_ms(nsemh_solver* h, double out[3]) { GUARD(if (nsem_halo_wait_ms((*h->sp).ctx, out)) throw Error(nsem_last_error((*h->sp).ctx))) }
int nsemh_set_schedule(nsemh_solver* h, const uint32_t* order, uint32_t n) {
    GUARD(if (nsem_set_schedule((*h->sp).ctx, order, n)) throw Error(nsem_last_error((*h->sp).ctx)))
}

// out = {NPX, NPY, NPZ, NP, NPF, nBCS, nCells, nFacets, gBCSfield, gALL}
int nsemh_dims(nsemh_solver* h, uint64_t out[10]) {
    Basis b((*h->sp).nop);
    const Geometry& g = (*h->sp).geo;
    const uint64_t v[10] = {(uint64_t)b.NPX, (uint64_t)b.NPY, (uint64_t)b.NPZ, (uint64_t)b.NP, (uint64_t)b.NPF,
                            g.nBCS, g.nCells, g.nFacets, g.gBCSfield, g.gALL};
    std::memcpy(out, v, sizeof v);
    return 0;
}
int nsemh_params(nsemh_solver* h, double out[12]) {
    const EulerSolver& s = (*h->sp);
    const double v[12] = {s.P0, s.T0, s.cp, s.cv, s.viscosity, s.Pr, s.gravity[0], s.gravity[1], s.gravity[2], s.dt,
                          (double)s.buoyancy, (double)s.diffusion};
    std::memcpy(out, v, sizeof v);
    return 0;
}

const double* nsemh_f64(nsemh_solver* h, const char* name, uint64_t* n) {
    EulerSolver& s = (*h->sp);
    const std::string k = name;
    const std::vector<double>* v = nullptr;
    if (k == "gCC") { *n = s.topo.CC.size() * 3; return s.topo.CC.empty() ? nullptr : s.topo.CC.data()->data(); }   // cell centroids (mesh.cpp:450-577)
    if (k == "gCV") v = &s.topo.CV; else
    if (k == "cC") v = &s.geo.cC; else if (k == "cV") v = &s.geo.cV; else if (k == "Jinv") v = &s.geo.Jinv;
    else if (k == "fN") v = &s.geo.fN; else if (k == "fC") v = &s.geo.fC; else if (k == "fI") v = &s.geo.fI;
    else if (k == "faceNormal") v = &s.geo.faceNormal; else if (k == "faceCenter") v = &s.geo.faceCenter;
    else if (k.size() == 7 && k.compare(0, 6, "psiRef") == 0 && k[6] >= '0' && k[6] <= '5') v = &s.geo.psiRef[k[6] - '0'];
    else if (k.size() == 7 && k.compare(0, 6, "psiCor") == 0 && k[6] >= '0' && k[6] <= '5') v = &s.geo.psiCor[k[6] - '0'];
    else if (k == "rho") v = &s.rho; else if (k == "U") v = &s.U; else if (k == "T") v = &s.T; else if (k == "p") v = &s.p;
    else if (k == "rho_ref") v = &s.rho_ref; else if (k == "p_ref") v = &s.p_ref; else if (k == "g") v = &s.gvec;
    else if (k == "out_rho") v = &s.out_rho; else if (k == "out_U") v = &s.out_U; else if (k == "out_T") v = &s.out_T; else if (k == "out_p") v = &s.out_p;
    if (!v) { *n = 0; return nullptr; }
    *n = v->size();
    return v->data();
}
const uint32_t* nsemh_u32(nsemh_solver* h, const char* name, uint64_t* n) {
    EulerSolver& s = (*h->sp);
    const std::string k = name;
    const std::vector<u32>* v = nullptr;
    if (k == "FO") v = &s.geo.FO; else if (k == "FN") v = &s.geo.FN; else if (k == "faceBegin") v = &s.geo.faceBegin;
    else if (k == "faceEnd") v = &s.geo.faceEnd; else if (k == "allFaces") v = &s.geo.allFaces; else if (k == "faceID") v = &s.geo.faceID;
    else if (k == "cellGlobal") v = &s.cellGlobal;
    else if (k == "refineMap") v = &s.last_maps.refineMap; else if (k == "coarseMap") v = &s.last_maps.coarseMap; else if (k == "cellMap") v = &s.last_maps.cellMap;
    else if (k == "faceOwner") v = &s.geo.faceOwner; else if (k == "faceNeigh") v = &s.geo.faceNeigh; else if (k == "faceMortar") v = &s.geo.faceMortar;
    if (!v) { *n = 0; return nullptr; }
    *n = v->size();
    static const uint32_t none = 0;                    // a known but empty array is not an unknown name
    return v->empty() ? &none : v->data();
}
double* nsemh_state_ptr(nsemh_solver* h, const char* name) {
    EulerSolver& s = (*h->sp);
    const std::string k = name;
    if (k == "rho") return s.rho.data();
    if (k == "U") return s.U.data();
    if (k == "T") return s.T.data();
    if (k == "p") return s.p.data();
    return nullptr;
}
// faces of boundary patch `name` (local ids); returns the count, copies at most cap entries
uint64_t nsemh_patch_faces(nsemh_solver* h, const char* name, uint32_t* out, uint64_t cap) {
    auto it = (*h->sp).topo.boundaries.find(name);
    if (it == (*h->sp).topo.boundaries.end()) return 0;
    for (uint64_t q = 0; q < it->second.size() && q < cap; q++) out[q] = it->second[q];
    return it->second.size();
}
// number of neighbouring ranks; copies them to out (ascending)
int nsemh_peers(nsemh_solver* h, int* out, int cap) {
    for (int q = 0; q < (int)(*h->sp).peers.size() && q < cap; q++) out[q] = (*h->sp).peers[q];
    return (int)(*h->sp).peers.size();
}
// Decompose the grid file <grid_noext>.{txt,bin} into nparts (Prepare::decomposeMesh, field.cpp:1086-1257) without
// building a solver: part_out[cell] = rank, mortar_out[face] = gFMC in the grid's own face numbering (either may be
// NULL).  Returns the number of cells, or -1 (message via nsemh_error) -- e.g. when a non-conforming face would be cut.
int64_t nsemh_partition_grid(const char* grid_noext, int nparts, const char* method, int px, int py, int pz,
                             uint32_t* part_out, uint32_t* mortar_out) {
    try {
        const Grid g = read_grid(grid_noext);
        const std::vector<u32> fmc = mortar_flags(g);
        const bool amr = std::any_of(fmc.begin(), fmc.end(), [](u32 v) { return v != 0; });
        const int pxyz[3] = {px, py, pz};
        const std::vector<u32> part = partition_cells(g, nparts, method ? method : "METIS", pxyz, amr ? &fmc : nullptr);
        if (part_out) std::copy(part.begin(), part.end(), part_out);
        if (mortar_out) std::copy(fmc.begin(), fmc.end(), mortar_out);
        return (int64_t)g.nCells();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

void nsemh_totals(nsemh_solver* h, double out[3]) { out[0] = (*h->sp).mass0; out[1] = (*h->sp).energy0; out[2] = (*h->sp).volume0; }

}  // extern "C"
