// euler_app.cpp -- host mirror of apps/euler/euler.cpp around the GPU time loop.
//
//   parameters          apps/euler/euler.cpp:17-51, apps/utils/properties.cpp:14-34, src/field/field.cpp:496-552
//   field read + ICs    src/field/field.h:1412-1586 (analytic initialisers run over ALL nodes, ghosts included)
//   applyExplicitBCs    src/field/field.h:2586-2727 (host version, used during set-up only)
//   set-up              apps/euler/euler.cpp:58-176 (isentropic vortex, gravity, hydrostatic reference state,
//                       rho from p, scaleBCs/fixedBCs)
//   time loop           apps/euler/euler.cpp:179-287 -> nsem_euler_step (CUDA)
//   Iteration::next     src/solvers/iteration.h:62-84 (dump every write_interval steps)
#include <algorithm>
#include <cmath>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>

#include "nsem_host.h"

namespace nsemh {

namespace {
constexpr double PI = 3.14159265358979323846264;
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline double mag3(const double* a) { return std::sqrt(dot3(a, a)); }
// equal(p,q) with EqualEpsilon (tensor.h:469-474)
inline bool equal(double p, double q, double tol = 1e-7) {
    const double d = std::fabs(p - q);
    return d <= tol || d <= tol * std::fabs(p) || d <= tol * std::fabs(q);
}
}  // namespace

EulerSolver::~EulerSolver() {
    if (ctx) nsem_destroy(ctx);
}

void EulerSolver::read_controls(const std::string& case_dir) {
    dir = case_dir;
    ctl = Controls::read(dir + "/controls");
    const std::string solver = ctl.str("general", "solver", "euler");
    if (solver != "euler" && solver != "convection") throw Error("Incorrect solver euler used instead of " + solver + ".");   // wrapper.cpp:32-38
    convection = (solver == "convection");
    if (convection) {
        conv_init = ctl.str("convection", "problem_init", "NONE");          // convection.cpp:26-29
        if (conv_init != "NONE" && conv_init != "LEVEQUE" && conv_init != "LAURITZEN_0" && conv_init != "LAURITZEN_1")
            throw Error("convection{problem_init " + conv_init + "} is not one of NONE, LEVEQUE, LAURITZEN_0, LAURITZEN_1");
        const bool sph = ctl.yes("general", "is_spherical", false);
        // init_wind_field leaves u, v unset in the other combinations (convection.cpp:55-82)
        if ((conv_init == "LEVEQUE" && sph) || (conv_init.rfind("LAURITZEN", 0) == 0 && !sph))
            throw Error("convection{problem_init " + conv_init + "} needs general{is_spherical " + (sph ? "NO" : "YES") + "}");
    }
    meshName = ctl.str("general", "mesh", "grid");
    nop[0] = (int)ctl.integer("general", "npx", 0);
    nop[1] = (int)ctl.integer("general", "npy", 0);
    nop[2] = (int)ctl.integer("general", "npz", 0);
    viscosity = ctl.num("general", "viscosity", viscosity);
    Pr = ctl.num("general", "Pr", Pr);
    T0 = ctl.num("general", "T0", T0);
    P0 = ctl.num("general", "P0", P0);
    cp = ctl.num("general", "cp", cp);
    cv = ctl.num("general", "cv", cv);
    dt = ctl.num("general", "dt", dt);
    gravity = ctl.vec("general", "gravity", gravity);
    time_scheme = ctl.str("general", "time_scheme", time_scheme);
    end_step = ctl.integer("general", "end_step", end_step);
    write_interval = std::max(1L, ctl.integer("general", "write_interval", write_interval));
    // the controls count start_step in time steps; the run starts from dump start_step / write_interval (AmrIteration, iteration.h:102),
    // which is what the member holds from here on
    start_step = ctl.integer("general", "start_step", 0) / write_interval;
    binary_out = ctl.str("general", "write_format", "BINARY") != "TEXT";
    buoyancy = ctl.yes("euler", "buoyancy", buoyancy);
    diffusion = ctl.yes("euler", "diffusion", diffusion);
    problem_init = ctl.str("euler", "problem_init", problem_init);
    // prepare{fields N { ... }} and vtk{} as apps/prepare reads them (prepareApp.cpp:63-102); NSEM_VTK=1 (or vtk{on_dump YES}, an addition) also
    // writes <mesh><k>.vtk next to every dump, straight from the downloaded state
    if (ctl.has("prepare", "fields")) {
        vtk_fields.clear();
        const auto& v = ctl.blocks.at("prepare").at("fields");
        for (size_t i = 1; i < v.size(); i++) if (v[i] != "{" && v[i] != "}") vtk_fields.push_back(v[i]);
    }
    vtk_cell_value = ctl.yes("vtk", "write_cell_value", vtk_cell_value);
    vtk_polyhedral = ctl.yes("vtk", "write_polyhedral", vtk_polyhedral);
    vtk_on_dump = ctl.yes("vtk", "on_dump", false) || (std::getenv("NSEM_VTK") && std::atoi(std::getenv("NSEM_VTK")) > 0);
    decomp_type = ctl.str("decomposition", "type", decomp_type);
    {
        const Vec3 n = ctl.vec("decomposition", "n", Vec3{1, 1, 1});
        for (int d = 0; d < 3; d++) decomp_n[d] = std::max(1, (int)n[d]);
    }
    conv_scheme = ctl.str("general", "convection_scheme", "RUSANOV");
    blend_factor = ctl.num("general", "blend_factor", blend_factor);
    // the euler app hands divf its lambdaMax only under RUSANOV; the convection app's scalar also takes the plain face values
    if (conv_scheme != "RUSANOV" && !(convection && (conv_scheme == "CDS" || conv_scheme == "UDS" || conv_scheme == "BLENDED")))
        throw Error("convection_scheme " + conv_scheme + " is not implemented on the GPU path (RUSANOV; CDS, UDS, BLENDED for solver convection)");
    // BDF1, AB1 and RK1..RK4 are the same single forward-Euler stage on this path (SURVEY finding 1)
    const std::string& ts = time_scheme;
    const bool ab_multi = (ts == "AB2" || ts == "AB3" || ts == "AB4" || ts == "AB5");
    // the scalar of the convection app also takes AB2..AB5 (examples/atmo/advection-leveque ships AB2): residual history on the device
    if (!(ts == "BDF1" || ts == "AB1" || ts == "RK1" || ts == "RK2" || ts == "RK3" || ts == "RK4" || (convection && ab_multi)))
        throw Error("time_scheme " + ts + " is not implemented on the GPU path (BDF1, AB1, RK1-RK4 are; AB2-AB5 for solver convection)");
    // cubed-sphere shells (field.cpp:516-519): the mesh loader projects the grid onto the sphere, set-up uses radial gravity
    topo.spherical = ctl.yes("general", "is_spherical", false);
    topo.sphere_radius = ctl.num("general", "sphere_radius", topo.sphere_radius);
    topo.sphere_height = ctl.num("general", "sphere_height", topo.sphere_height);
    // AmrIteration (iteration.h:94-147): a regrid before step 1 and after every amr_step dumps; here in memory (amr.cpp), the state stays
    // on the device (nsem_refine_state).  NSEM_IGNORE_AMR_STEP=1 runs on the grid as it is.
    amr_step = std::getenv("NSEM_IGNORE_AMR_STEP") ? 0 : ctl.integer("general", "amr_step", 0);
    refine_params.dir = ctl.vec("refinement", "direction", refine_params.dir);
    refine_params.field = ctl.str("refinement", "field", refine_params.field);
    refine_params.field_max = ctl.num("refinement", "field_max", refine_params.field_max);
    refine_params.field_min = ctl.num("refinement", "field_min", refine_params.field_min);
    refine_params.max_level = (int)ctl.integer("refinement", "max_level", refine_params.max_level);
    refine_params.buffer_zone = (int)ctl.integer("refinement", "buffer_zone", refine_params.buffer_zone);
    refine_params.limit = ctl.integer("refinement", "limit", refine_params.limit);
    if (ctl.str("general", "state", "STEADY") != "TRANSIENT") throw Error("state must be TRANSIENT");
}

void EulerSolver::set_mesh(const Grid& g) {
    const bool verbose = std::getenv("NSEM_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto t1 = std::chrono::steady_clock::now();
        if (verbose) std::printf("set_mesh[%d]: %-32s %.3f s\n", rank, what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    // Prepare::convertVTK loads with remove_empty = false (prepare.cpp:12): a 2-D cell keeps its two empty faces, the corners of its node
    // placement are then taken from sides 0/1 instead of the first remaining pair -- the same nodes on a flat mesh, but on the sphere the
    // radial rescale of the placement depends on that choice (the files of `prepare -vtk` differ from the solver's geometry by a metre)
    topo.keep_empty = vtk_mode && topo.spherical;
    topo.load(g);
    lap("topology (MeshTopo::load)");
    Basis b(nop);
    geo.build(topo, b);
    lap("node geometry (Geometry::build)");
    // the AMR forest starts from the grid as loaded (conforming hexahedra); only built when a regrid can follow
    forest.reset();
    if ((amr_step != 0 || std::getenv("NSEM_AMR")) && nranks == 1) {       // a partition has no forest: the whole-domain solver regrids (run_case)
        forest = std::make_shared<AmrForest>();
        struct stat st;
        bool loaded = false;
        if (!forest_file.empty() && ::stat(forest_file.c_str(), &st) == 0) {
            // restart of an AMR run: the forest that emitted this grid was saved next to it.  A forest that does not match (the case's own
            // conforming <mesh>_0.txt read again next to the <mesh>_0.bin + .forest an earlier run's initial regrid left) is not this grid's
            forest->load(forest_file);
            loaded = forest->leaves.size() == g.nCells();
        }
        if (!loaded) {
            forest = std::make_shared<AmrForest>();
            try { forest->init(g, refine_params.dir); }
            catch (const Error&) { forest.reset(); }  // a grid that is already non-conforming cannot seed the forest: regridded() says so
        }
    }
    if (forest) forest->spherical = topo.spherical;
    forest_file.clear();
}

void EulerSolver::set_mesh_partition(const Grid& global, int rank_, int nranks_, const std::string& type, const int nxyz[3]) {
    rank = rank_; nranks = nranks_;
    nGlobalCells = global.nCells();
    const bool verbose = std::getenv("NSEM_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto t1 = std::chrono::steady_clock::now();
        if (verbose) std::printf("decompose[%d]: %-34s %.3f s\n", rank, what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    const std::vector<u32> fmc = mortar_flags(global);
    lap("mortar flags of the global grid");
    const bool amr = std::any_of(fmc.begin(), fmc.end(), [](u32 v) { return v != 0; });
    // face j of a CYCLIC patch is paired with face j of its neighbor patch (field.h:2662-2664, 2683-2687): their owner cells stay together
    std::vector<std::array<u32, 2>> together;
    for (const auto& pr : cyclic_patches) {
        auto a = global.boundaries.find(pr[0]), b = global.boundaries.find(pr[1]);
        if (a == global.boundaries.end() || b == global.boundaries.end() || a->second.size() != b->second.size())
            throw Error("CYCLIC patches " + pr[0] + "/" + pr[1] + " missing or of different size");
        for (size_t j = 0; j < a->second.size(); j++) together.push_back({a->second[j], b->second[j]});
    }
    const std::vector<u32> part = partition_cells(global, nranks, type, nxyz, amr ? &fmc : nullptr, &together);
    lap("partition_cells");
    if (verbose && rank == 0) {
        // quality of the partition: cells and cut faces per part (the step time follows the largest of both)
        std::vector<uint64_t> cells(nranks, 0), cut(nranks, 0);
        for (u32 c = 0; c < global.nCells(); c++) cells[part[c]]++;
        std::vector<u32> owner(global.nFacets(), MAX_INT);
        for (u32 c = 0; c < global.nCells(); c++)
            for (u32 q = global.cellStart[c]; q < global.cellStart[c + 1]; q++) {
                const u32 f = global.cellFaces[q];
                if (owner[f] == MAX_INT) owner[f] = c;
                else if (part[owner[f]] != part[c]) { cut[part[owner[f]]]++; cut[part[c]]++; }
            }
        for (int r = 0; r < nranks; r++) std::printf("decompose: part %d: %llu cells, %llu cut faces\n", r, (unsigned long long)cells[r], (unsigned long long)cut[r]);
    }
    Partition P = extract_partition(global, part, rank, nranks);
    lap("extract_partition");
    if (P.cellGlobal.empty()) throw Error("partition " + std::to_string(rank) + " is empty");
    cellGlobal = P.cellGlobal;
    peers = P.peers;
    if (topo.spherical) {
        // ExtrudeMesh scales by the extremes of the vertices it sees (mesh.cpp:735-741); a part need not reach both shells, so every part
        // projects with the whole grid's extremes and lands on the single-partition mesh bit for bit
        double minh = 1e30, maxh = 0;
        for (const Vec3& v : global.V) {
            const double h = std::max(std::max(std::fabs(v[0]), std::fabs(v[1])), std::fabs(v[2]));
            maxh = std::max(maxh, h); minh = std::min(minh, h);
        }
        topo.shell_h[0] = minh; topo.shell_h[1] = maxh;
    }
    set_mesh(P.grid);
    lap("local topology + node geometry");
}

static bool grid_exists(const std::string& path_noext) {
    struct stat st;
    return ::stat((path_noext + ".txt").c_str(), &st) == 0 || ::stat((path_noext + ".bin").c_str(), &st) == 0;
}
void EulerSolver::load_mesh(int step_) {
    // findLastRefinedGrid (field.cpp:79-91): the newest <mesh>_<k>, k <= step; a run restarted from dump k of a fixed mesh reads <mesh>_0
    int step = step_;
    while (step >= 0 && !grid_exists(dir + "/" + meshName + "_" + std::to_string(step))) step--;
    if (step < 0) step = step_;
    const Grid g = read_grid(dir + "/" + meshName + "_" + std::to_string(step));
    forest_file = dir + "/" + meshName + "_" + std::to_string(step) + ".forest";
    // one process per partition: every rank decomposes the same global grid the same way (Prepare::decomposeMesh does it
    // once on rank 0 and hands the parts over through grid<r>/ files, field.cpp:1086-1443) and keeps its own part
    if (nranks > 1) {
        // the CYCLIC pairs are in the field files, which are read after the mesh: look at their boundary sections first
        cyclic_patches.clear();
        const std::string s = std::to_string(step_);
        for (const char* f : {"U", "T", "p", "rho"}) {
            struct stat st;
            if (::stat((dir + "/" + f + s + ".txt").c_str(), &st) != 0 && ::stat((dir + "/" + f + s + ".bin").c_str(), &st) != 0) continue;
            const FieldFile ff = read_field(dir + "/" + f + s, std::string(f) == "U" ? 3 : 1);
            for (const BCond& b : ff.bcs) {
                if (b.type != "CYCLIC") continue;
                bool seen = false;
                for (const auto& pr : cyclic_patches) seen |= (pr[0] == b.patch || pr[1] == b.patch);
                if (!seen) cyclic_patches.push_back({b.patch, b.neighbor});
            }
        }
        set_mesh_partition(g, rank, nranks, decomp_type, decomp_n);
    } else { nGlobalCells = g.nCells(); set_mesh(g); }
}

// ---------------------------------------------------------------------------------------------------------
std::vector<double> init_field(const FieldFile& ff, const Geometry& g, const Vec3& gravity) {
    const int c = ff.comps;
    const uint64_t gA = g.gALL;
    std::vector<double> out(gA * c, 0.0);
    if (!ff.values.empty()) {
        std::copy(ff.values.begin(), ff.values.begin() + std::min(ff.values.size(), out.size()), out.begin());
        return out;
    }
    for (const auto& in : ff.inits) {
        const double* a = in.a.data();
        if (in.kind == "uniform") {
#pragma omp parallel for schedule(static)
            for (uint64_t i = 0; i < gA; i++) for (int d = 0; d < c; d++) out[i * c + d] += a[d];
        } else if (in.kind == "cosine" || in.kind == "cosine2" || in.kind == "linear" || in.kind == "gaussian") {
            const double *value = a, *pert = a + c, *center = a + 2 * c, *radius = a + 2 * c + 3;
            const double pw = (in.kind == "cosine2") ? 2.0 : 1.0;
#pragma omp parallel for schedule(static)
            for (uint64_t i = 0; i < gA; i++) {
                double q[3];
                for (int d = 0; d < 3; d++) q[d] = (g.cC[i * 3 + d] - center[d]) / radius[d];
                double R = mag3(q);
                if (g.spherical) {
                    // centre given as (radius, latitude, longitude); distance along the great circle (field.h:1441-1444)
                    const double* x = &g.cC[i * 3];
                    const double sr = mag3(x), lat = std::atan2(x[2], std::sqrt(x[0] * x[0] + x[1] * x[1])), lon = std::atan2(x[1], x[0]);
                    double dd = (center[0] + sr) / 2;
                    dd *= std::acos(std::sin(center[1]) * std::sin(lat) + std::cos(center[1]) * std::cos(lat) * std::cos(center[2] - lon));
                    R = dd / mag3(radius);
                }
                for (int d = 0; d < c; d++) {
                    double val = value[d];
                    if (in.kind == "gaussian") {
                        double v = std::exp(-R * R);
                        if (equal(v, 0.0)) v = 0;
                        val += pert[d] * v;
                    } else {
                        const double Rc = (1.0 <= R) ? 1.0 : R;
                        if (in.kind == "linear") val += pert[d] * (1.0 - Rc);
                        else val += (pert[d] / 2) * std::pow(1.0 + std::cos(Rc * PI), pw);
                    }
                    out[i * c + d] += val;
                }
            }
        } else if (in.kind == "gaussian-outside") {
            const double *value = a, *pert = a + c, *center = a + 2 * c;
            const double radius = a[2 * c + 3], radius2 = a[2 * c + 4];
            for (uint64_t i = 0; i < gA; i++) {
                double q[3];
                for (int d = 0; d < 3; d++) q[d] = g.cC[i * 3 + d] - center[d];
                const double R = (mag3(q) - radius) / radius2;
                for (int d = 0; d < c; d++) {
                    double val = value[d];
                    if (R <= 0) val += pert[d];
                    else {
                        double v = std::exp(-R * R);
                        if (equal(v, 0.0)) v = 0;
                        val += pert[d] * v;
                    }
                    out[i * c + d] += val;
                }
            }
        } else if (in.kind == "hydrostatic") {
            const double *p0 = a, scale = a[c], expon = a[c + 1];
            for (uint64_t i = 0; i < gA; i++) {
                const double gh = g.spherical ? -(mag3(&g.cC[i * 3]) - g.sphere_radius) * mag3(gravity.data()) : dot3(&g.cC[i * 3], gravity.data());
                for (int d = 0; d < c; d++) out[i * c + d] += p0[d] * std::pow(1.0 + scale * gh, expon);
            }
        }
    }
    return out;
}

// applyExplicitBCs on host arrays (set-up only; in the time loop the CUDA bc_kernel does this)
void EulerSolver::apply_bcs(std::vector<double>& F, int comps, std::vector<BCond>& bcs) {
    const int NPF = Basis(nop).NPF;
    const uint64_t gA = geo.gALL;
    for (auto& bc : bcs) {
        if (bc.type == "GHOST" || bc.type == "UNLISTED") continue;
        auto it = topo.boundaries.find(bc.patch);
        if (it == topo.boundaries.end() || it->second.empty()) continue;
        const std::vector<u32>& faces = it->second;
        const std::vector<u32>* nb = nullptr;
        if (!bc.neighbor.empty()) {
            auto jt = topo.boundaries.find(bc.neighbor);
            if (jt == topo.boundaries.end()) throw Error("CYCLIC neighbor patch " + bc.neighbor + " not found");
            nb = &jt->second;
        }
        const bool fresh_fixed = (bc.type == "CALC_DIRICHLET" && bc.fixed.empty());
        if (fresh_fixed) bc.fixed.assign(faces.size() * (size_t)NPF * comps, 0.0);
        for (size_t j = 0; j < faces.size(); j++)
            for (int n = 0; n < NPF; n++) {
                const size_t k = (size_t)faces[j] * NPF + n;
                const u32 c1 = geo.FO[k], c2 = geo.FN[k];
                if (c2 >= gA) continue;
                double* gph = &F[(size_t)c2 * comps];
                const double* own = &F[(size_t)c1 * comps];
                if (bc.type == "NEUMANN") {
                    double dv[3];
                    for (int d = 0; d < 3; d++) dv[d] = geo.cC[(size_t)c2 * 3 + d] - geo.cC[(size_t)c1 * 3 + d];
                    const double m = mag3(dv);
                    for (int d = 0; d < comps; d++) gph[d] = own[d] + bc.value[d] * m;
                } else if (bc.type == "ROBIN") {
                    for (int d = 0; d < comps; d++) gph[d] = bc.shape * bc.value[d] + (1 - bc.shape) * own[d];
                } else if (bc.type == "SYMMETRY") {
                    if (comps == 1) gph[0] = own[0];
                    else {
                        // sym(Vector,Vector), tensor.h:486-494
                        const double* N = &geo.fN[k * 3];
                        const double mg = mag3(N);
                        const double en[3] = {N[0] / mg, N[1] / mg, N[2] / mg};
                        const double Axx = 1.0 - en[0] * en[0], Ayy = 1.0 - en[1] * en[1], Azz = 1.0 - en[2] * en[2];
                        const double Axy = 0.0 - en[0] * en[1], Ayz = 0.0 - en[1] * en[2], Axz = 0.0 - en[0] * en[2];
                        const double r[3] = {Axx * own[0] + Axy * own[1] + Axz * own[2], Axy * own[0] + Ayy * own[1] + Ayz * own[2],
                                             Axz * own[0] + Ayz * own[1] + Azz * own[2]};
                        const double magR = mag3(r);
                        if (equal(magR, 0.0)) { gph[0] = r[0]; gph[1] = r[1]; gph[2] = r[2]; }
                        else {
                            const double f = mag3(own) / magR;
                            gph[0] = r[0] * f; gph[1] = r[1] * f; gph[2] = r[2] * f;
                        }
                    }
                } else if (bc.type == "CYCLIC") {
                    if (!nb || nb->size() != faces.size()) throw Error("CYCLIC patches " + bc.patch + "/" + bc.neighbor + " differ in size");
                    const u32 c11 = geo.FO[(size_t)(*nb)[j] * NPF + n];
                    for (int d = 0; d < comps; d++) gph[d] = F[(size_t)c11 * comps + d];
                } else if (bc.type == "DIRICHLET") {
                    for (int d = 0; d < comps; d++) gph[d] = bc.value[d];
                } else if (bc.type == "CALC_DIRICHLET") {
                    double* fx = &bc.fixed[(j * NPF + n) * comps];
                    if (fresh_fixed) for (int d = 0; d < comps; d++) fx[d] = own[d];
                    for (int d = 0; d < comps; d++) gph[d] = fx[d];
                } else {
                    throw Error("boundary condition type " + bc.type + " is not implemented on the GPU path");
                }
            }
    }
}

void EulerSolver::set_fields(const FieldFile& frho, const FieldFile& fU, const FieldFile& fT, const FieldFile& fp) {
    p = init_field(fp, geo, gravity);
    U = init_field(fU, geo, gravity);
    T = init_field(fT, geo, gravity);
    rho = init_field(frho, geo, gravity);
    bc_p = fp.bcs; bc_U = fU.bcs; bc_T = fT.bcs; bc_rho = frho.bcs;
    file_bc_p = fp.bcs; file_bc_U = fU.bcs; file_bc_T = fT.bcs; file_bc_rho = frho.bcs;
    for (const auto& kv : topo.boundaries)
        if (kv.first.find("interMesh") != std::string::npos)
            for (auto* list : {&bc_p, &bc_U, &bc_T, &bc_rho}) {
                BCond b;
                b.patch = kv.first;
                b.type = "GHOST";
                list->insert(list->begin(), b);
            }
    // MeshField::read_ applies the BCs right after reading (field.h:1562-1565)
    apply_bcs(p, 1, bc_p);
    apply_bcs(U, 3, bc_U);
    apply_bcs(T, 1, bc_T);
    apply_bcs(rho, 1, bc_rho);
}

// raw node values of a field file are in the GLOBAL real-cell order: keep this partition's cells
static FieldFile localize(FieldFile ff, const std::vector<u32>& cellGlobal, u32 nGlobalCells, int NP) {
    if (ff.values.empty()) return ff;
    const size_t per = (size_t)NP * ff.comps;
    if (ff.values.size() < (size_t)nGlobalCells * per) throw Error("field file holds fewer node values than the global grid has nodes");
    std::vector<double> v(cellGlobal.size() * per);
    for (size_t l = 0; l < cellGlobal.size(); l++)
        std::copy(ff.values.begin() + (size_t)cellGlobal[l] * per, ff.values.begin() + ((size_t)cellGlobal[l] + 1) * per, v.begin() + l * per);
    ff.values.swap(v);
    return ff;
}

void EulerSolver::read_fields(int step) {
    const std::string s = std::to_string(step);
    if (convection) {
        // VectorCellField U("U"), ScalarCellField T("T") (convection.cpp:39-40); the scalar takes the rho slot, the other two stay zero
        FieldFile fU = read_field(dir + "/U" + s, 3), fT = read_field(dir + "/T" + s, 1);
        if (nranks > 1) {
            const int NP = Basis(nop).NP;
            fU = localize(fU, cellGlobal, nGlobalCells, NP); fT = localize(fT, cellGlobal, nGlobalCells, NP);
        }
        FieldFile zero;
        zero.comps = 1;
        zero.inits.push_back({"uniform", std::vector<double>(1, 0.0)});
        zero.bcs = fT.bcs;
        set_fields(fT, fU, zero, zero);
        return;
    }
    FieldFile frho;
    struct stat st;
    if (step != 0 || ::stat((dir + "/rho" + s + ".txt").c_str(), &st) == 0 || ::stat((dir + "/rho" + s + ".bin").c_str(), &st) == 0) {
        frho = read_field(dir + "/rho" + s, 1);
    } else {
        // MeshField::read skips a file that is not there (field.h:1579-1585; examples/atmo/hydro-sphere ships no rho0): the start branch
        // forms rho from p and T on every entry, boundary cells included, and with no conditions to apply they keep that value for the run.
        // Only for the initial state: a dump that is not there stays an error
        frho.comps = 1;
        frho.inits.push_back({"uniform", std::vector<double>(1, 0.0)});
        rho_file_missing = true;
    }
    FieldFile fU = read_field(dir + "/U" + s, 3), fT = read_field(dir + "/T" + s, 1), fp = read_field(dir + "/p" + s, 1);
    if (nranks > 1) {
        const int NP = Basis(nop).NP;
        frho = localize(frho, cellGlobal, nGlobalCells, NP); fU = localize(fU, cellGlobal, nGlobalCells, NP);
        fT = localize(fT, cellGlobal, nGlobalCells, NP); fp = localize(fp, cellGlobal, nGlobalCells, NP);
    }
    set_fields(frho, fU, fT, fp);
}

void EulerSolver::mark_unlisted_patches() {
    // A boundary patch the field file has no condition for: nothing overwrites what SolveTexplicit leaves in its boundary cells
    // (solve.cpp:563-581), the owner's residual (fillBCs, field.h:2731-2743) over the boundary cell's own volume.  examples/atmo/hydro-sphere
    // ships no rho0 file, so rho is such a field there; the device continues those cells the same way (NSEM_BC_UNLISTED, volume ratios in
    // `fixed`).  For U, T and p the boundary cells would also pick up source terms and the equation of state: refused.
    const int NPF = Basis(nop).NPF;
    struct L { const char* name; std::vector<BCond>* bcs; } lists[4] = {{"rho", &bc_rho}, {"U", &bc_U}, {"T", &bc_T}, {"p", &bc_p}};
    for (const auto& kv : topo.boundaries) {
        if (kv.second.empty() || kv.first.find("interMesh") != std::string::npos) continue;
        for (auto& l : lists) {
            bool listed = false;
            for (const auto& b : *l.bcs) listed = listed || b.patch == kv.first;
            if (listed) continue;
            if (l.bcs != &bc_rho)
                throw Error(std::string("field ") + l.name + " has no boundary condition on patch " + kv.first + " (only rho may go without one on the GPU path)");
            BCond b;
            b.patch = kv.first;
            b.type = "UNLISTED";
            b.held = true;
            b.fixed.assign(kv.second.size() * (size_t)NPF, 0.0);
            for (size_t j = 0; j < kv.second.size(); j++)
                for (int n = 0; n < NPF; n++) {
                    const size_t k = (size_t)kv.second[j] * NPF + n;
                    const u32 c1 = geo.FO[k], c2 = geo.FN[k];
                    if (c2 >= geo.gALL) continue;
                    b.fixed[j * NPF + n] = geo.cV[c1] / geo.cV[c2];
                }
            l.bcs->push_back(std::move(b));
        }
    }
}

static std::vector<BCond> scale_bcs(const std::vector<BCond>& src) {
    // Mesh::scaleBCs with psi = 1 (field.h:2797-2826): NEUMANN/ROBIN/SYMMETRY/CYCLIC stay, the rest freeze
    std::vector<BCond> out = src;
    for (auto& b : out)
        if (!(b.type == "NEUMANN" || b.type == "ROBIN" || b.type == "SYMMETRY" || b.type == "CYCLIC")) {
            b.type = "CALC_DIRICHLET";
            b.fixed.clear();
        }
    return out;
}

void EulerSolver::setup() {
    if (convection) {
        // nothing of the euler set-up applies (convection.cpp:89-111): no reference state, no gravity; total scalar for the loss line
        const uint64_t gA = geo.gALL, gB = geo.gBCSfield;
        gvec.assign(gA * 3, 0.0); gh.assign(gA, 0.0); rho_ref.assign(gA, 0.0); p_ref.assign(gA, 0.0);
        buoyancy = false; diffusion = false;
        scalar0 = 0; volume0 = 0;
        for (uint64_t i = 0; i < gB; i++) { scalar0 += rho[i] * geo.cV[i]; volume0 += geo.cV[i]; }
        return;
    }
    const uint64_t gA = geo.gALL, gB = geo.gBCSfield;
    const double R = cp - cv, gamma = cp / cv;
    if (problem_init == "ISENTROPIC_VORTEX") {        // euler.cpp:80-95
        const double beta = 5;
        for (uint64_t i = 0; i < gA; i++) {
            const double r = mag3(&geo.cC[i * 3]);
            T[i] = (-((gamma - 1) * beta * beta) / (8 * gamma * PI * PI)) * std::exp(1 - r * r);
        }
        for (uint64_t i = 0; i < gB; i++) {
            const double r = mag3(&geo.cC[i * 3]);
            U[i * 3 + 0] += (beta / (2 * PI)) * std::exp((1 - r * r) / 2.0) * -geo.cC[i * 3 + 1];
            U[i * 3 + 1] += (beta / (2 * PI)) * std::exp((1 - r * r) / 2.0) * geo.cC[i * 3 + 0];
        }
        for (uint64_t i = 0; i < gA; i++) {
            p[i] = std::pow(T[i] + T0, gamma / (gamma - 1)) - P0;
            rho[i] = (P0 / (R * (T[i] + T0))) * std::pow((p[i] + P0) / P0, 1 / gamma) - (P0 / (R * T0));
        }
    } else if (problem_init != "NONE") {
        throw Error("unknown problem_init " + problem_init);
    }
    gvec.assign(gA * 3, 0.0);
    gh.assign(gA, 0.0);
    p_ref.assign(gA, P0);
    rho_ref.assign(gA, P0 / (R * T0));
    if (buoyancy) {                                    // euler.cpp:105-123
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < gA; i++) {
            if (geo.spherical) {                       // gravity points to the centre of the sphere, euler.cpp:109-111
                const double r = mag3(&geo.cC[i * 3]), mg = mag3(gravity.data());
                for (int d = 0; d < 3; d++) gvec[i * 3 + d] = -(geo.cC[i * 3 + d] / r) * mg;
                gh[i] = -(r - geo.sphere_radius) * mg;
                continue;
            }
            for (int d = 0; d < 3; d++) gvec[i * 3 + d] = gravity[d];
            gh[i] = dot3(&gvec[i * 3], &geo.cC[i * 3]);
        }
        bc_g = bc_U;                                   // Mesh::fixedBCs<Vector>(U,g), field.h:2779-2795
        for (auto& b : bc_g) { b.type = "CALC_DIRICHLET"; b.fixed.clear(); }
        apply_bcs(gvec, 3, bc_g);
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < gA; i++) {
            p_ref[i] = P0 * std::pow(1.0 + gh[i] / (cp * T0), cp / R);
            rho_ref[i] = (P0 / (R * T0)) * std::pow(p_ref[i] / P0, 1 / gamma);
        }
    }
    // ait.start() branch, euler.cpp:133-146
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < gA; i++) p[i] += p_ref[i];
    bc_p_ref = scale_bcs(bc_p);
    apply_bcs(p_ref, 1, bc_p_ref);
    apply_bcs(p, 1, bc_p);
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < gA; i++) rho[i] = (P0 / (R * (T[i] + T0))) * std::pow(p[i] / P0, 1 / gamma);
    apply_bcs(rho, 1, bc_rho);
    bc_rho_ref = scale_bcs(bc_rho);
    apply_bcs(rho_ref, 1, bc_rho_ref);
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < gA; i++) p[i] -= p_ref[i];
    mark_unlisted_patches();
    // totals (euler.cpp:164-176): one pow per node -- fixed chunks summed in parallel, the chunk sums added in order, so the result does not
    // depend on the number of threads
    mass0 = energy0 = volume0 = 0;
    {
        const uint64_t chunk = 1u << 16, nchunks = (gB + chunk - 1) / chunk;
        std::vector<double> part(nchunks * 3, 0.0);
#pragma omp parallel for schedule(static)
        for (int64_t q = 0; q < (int64_t)nchunks; q++) {
            double m = 0, e_ = 0, v = 0;
            const uint64_t i1 = std::min<uint64_t>(gB, (uint64_t)(q + 1) * chunk);
            for (uint64_t i = (uint64_t)q * chunk; i < i1; i++) {
                const double sf = rho[i] * geo.cV[i];
                m += sf;
                const double e = gh[i] + 0.5 * dot3(&U[i * 3], &U[i * 3]) + std::pow((p[i] + p_ref[i]) / P0, R / cp) * (T[i] + T0) * cv;
                e_ += rho[i] * geo.cV[i] * e;
                v += geo.cV[i];
            }
            part[q * 3] = m; part[q * 3 + 1] = e_; part[q * 3 + 2] = v;
        }
        for (uint64_t q = 0; q < nchunks; q++) { mass0 += part[q * 3]; energy0 += part[q * 3 + 1]; volume0 += part[q * 3 + 2]; }
    }
}

// ---------------------------------------------------------------------------------------------------------
static int kind_of(const std::string& t) {
    if (t == "NEUMANN") return NSEM_BC_NEUMANN;
    if (t == "DIRICHLET") return NSEM_BC_DIRICHLET;
    if (t == "SYMMETRY") return NSEM_BC_SYMMETRY;
    if (t == "CYCLIC") return NSEM_BC_CYCLIC;
    if (t == "GHOST") return NSEM_BC_GHOST;
    if (t == "CALC_DIRICHLET") return NSEM_BC_FIXED;
    if (t == "ROBIN") return NSEM_BC_ROBIN;
    if (t == "UNLISTED") return NSEM_BC_UNLISTED;
    throw Error("boundary condition type " + t + " is not implemented on the GPU path");
}

void EulerSolver::build_c_bcs() {
    c_bcs_.clear();
    keep_faces_.clear();
    struct Src { int field; std::vector<BCond>* list; };
    Src srcs[4] = {{NSEM_F_RHO, &bc_rho}, {NSEM_F_P, &bc_p}, {NSEM_F_U, &bc_U}, {NSEM_F_T, &bc_T}};
    for (auto& s : srcs)
        for (auto& b : *s.list) {
            auto it = topo.boundaries.find(b.patch);
            if (it == topo.boundaries.end() || it->second.empty()) continue;
            nsem_bc c;
            std::memset(&c, 0, sizeof c);
            c.field = s.field;
            c.kind = kind_of(b.type);
            c.n_faces = (u32)it->second.size();
            c.faces = it->second.data();
            if (c.kind == NSEM_BC_CYCLIC) {
                auto jt = topo.boundaries.find(b.neighbor);
                if (jt == topo.boundaries.end() || jt->second.size() != it->second.size())
                    throw Error("CYCLIC neighbor patch of " + b.patch + " missing or of different size");
                c.peer_faces = jt->second.data();
            }
            for (int d = 0; d < 3; d++) { c.value[d] = b.value[d]; c.tvalue[d] = b.tvalue[d]; }
            c.shape = b.shape; c.tshape = b.tshape; c.zMin = b.zMin;
            if (c.kind == NSEM_BC_FIXED || c.kind == NSEM_BC_UNLISTED) {
                if (b.fixed.empty()) throw Error("CALC_DIRICHLET on " + b.patch + " has no frozen values yet");
                c.fixed = b.fixed.data();
            }
            c_bcs_.push_back(c);
        }
}

// Processing order of the elements on the GPU: a Z-curve over the cell centroids, so that elements launched together
// are neighbours in space and their face-trace gathers hit L2 (the mesh order of the block mesher is x-major: an
// x-neighbour is ny*nz elements away).  Only the order of execution changes, never the result.
static std::vector<u32> morton_schedule(const MeshTopo& t) {
    const u32 n = t.nBCS;
    Vec3 lo{1e300, 1e300, 1e300}, hi{-1e300, -1e300, -1e300};
    for (u32 c = 0; c < n; c++)
        for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], t.CC[c][d]); hi[d] = std::max(hi[d], t.CC[c][d]); }
    auto spread = [](uint64_t v) {   // 21 bits -> every third bit
        v &= 0x1fffff;
        v = (v | v << 32) & 0x1f00000000ffffull;
        v = (v | v << 16) & 0x1f0000ff0000ffull;
        v = (v | v << 8) & 0x100f00f00f00f00full;
        v = (v | v << 4) & 0x10c30c30c30c30c3ull;
        v = (v | v << 2) & 0x1249249249249249ull;
        return v;
    };
    std::vector<std::pair<uint64_t, u32>> key(n);
    for (u32 c = 0; c < n; c++) {
        uint64_t q[3];
        for (int d = 0; d < 3; d++) {
            const double w = hi[d] - lo[d];
            q[d] = w > 0 ? (uint64_t)std::min(1023.0, std::floor((t.CC[c][d] - lo[d]) / w * 1024.0)) : 0;
        }
        key[c] = {spread(q[0]) << 2 | spread(q[1]) << 1 | spread(q[2]), c};
    }
    std::sort(key.begin(), key.end());
    std::vector<u32> order(n);
    for (u32 c = 0; c < n; c++) order[c] = key[c].second;
    return order;
}

void EulerSolver::attach_device(int device, int rank, int nranks, const void* uid) {
    if (ctx) { nsem_destroy(ctx); ctx = nullptr; }
    if (nsem_create(device, rank, nranks, uid, &ctx)) throw Error(nsem_last_error(nullptr));
    device_id = nsem_device(ctx);
    auto ck = [&](int rc) { if (rc) throw Error(nsem_last_error(ctx)); };
    Basis b(nop);
    ck(nsem_set_order(ctx, b.NPX, b.NPY, b.NPZ));
    const double* dp[3] = {b.dpsi[0].data(), b.dpsi[1].data(), b.dpsi[2].data()};
    const double* wp[3] = {b.wgl[0].data(), b.wgl[1].data(), b.wgl[2].data()};
    ck(nsem_set_basis(ctx, dp, wp));
    nsem_mesh m = geo.as_c();
    ck(nsem_upload_mesh(ctx, &m));
    {
        const char* sc = std::getenv("NSEM_SCHEDULE");
        if (sc && std::string(sc) == "morton") {    // measured: no gain at 100^3 (the sweeps are latency-, not traffic-bound); mesh order is the default
            const std::vector<u32> order = morton_schedule(topo);
            ck(nsem_set_schedule(ctx, order.data(), (u32)order.size()));
        }
    }
    build_c_bcs();
    ck(nsem_set_bcs(ctx, c_bcs_.data(), (u32)c_bcs_.size()));
    nsem_params q;
    std::memset(&q, 0, sizeof q);
    q.P0 = P0; q.T0 = T0; q.cp = cp; q.cv = cv; q.viscosity = viscosity; q.Pr = Pr; q.dt = dt;
    for (int d = 0; d < 3; d++) q.gravity[d] = gravity[d];
    q.buoyancy = buoyancy; q.diffusion = diffusion;
    ck(nsem_set_params(ctx, &q));
    ck(nsem_upload_ref(ctx, rho_ref.data(), p_ref.data(), (geo.spherical && buoyancy) ? gvec.data() : nullptr));
    ck(nsem_upload_geopotential(ctx, gh.data()));
    // the field storage lives as long as the solver: page-lock it once for the per-dump transfers
    for (std::vector<double>* v : {&rho, &U, &T, &p}) nsem_pin_host(ctx, v->data(), v->size() * sizeof(double));
    upload_state();
    if (nranks > 1) {
        // gInterMesh (mesh.h:180-196): one entry per interMesh_<me>_<peer> patch
        std::vector<nsem_halo_peer> hp;
        for (const auto& kv : topo.boundaries) {
            if (kv.first.find("interMesh") == std::string::npos) continue;
            const size_t us = kv.first.rfind('_');
            nsem_halo_peer h;
            h.peer_rank = std::stoi(kv.first.substr(us + 1));
            h.n_faces = (u32)kv.second.size();
            h.faces = kv.second.data();
            hp.push_back(h);
        }
        ck(nsem_set_halo(ctx, hp.data(), (u32)hp.size()));
        exchange_setup_halos();
    }
    if (convection) {
        ck(nsem_upload_coords(ctx, geo.cC.data()));
        if (geo.spherical) ck(nsem_set_sphere(ctx, geo.sphere_radius));
        if (time_scheme.size() == 3 && time_scheme.compare(0, 2, "AB") == 0) ck(nsem_set_ab_order(ctx, time_scheme[2] - '0'));
        ck(nsem_set_convection_scheme(ctx, conv_scheme == "CDS" ? 1 : (conv_scheme == "UDS" ? 2 : (conv_scheme == "BLENDED" ? 3 : 0)), blend_factor));
        arm_wind(write_interval * start_step + 1);
    }
}

// the analytic wind is a function of the step number (convection.cpp:114-121): (re)set the step the next call starts with
void EulerSolver::arm_wind(long first_step) {
    if (!convection || !ctx) return;
    const int kind = conv_init == "LEVEQUE" ? 1 : (conv_init == "LAURITZEN_0" ? 2 : (conv_init == "LAURITZEN_1" ? 3 : 0));
    const long total = conv_end_step > 0 ? conv_end_step : end_step;       // the period is the WHOLE run's end_step * dt, also inside an AMR cycle
    if (nsem_set_convection(ctx, kind, (double)total * dt, first_step)) throw Error(nsem_last_error(ctx));
}

void EulerSolver::exchange_setup_halos() {
    if (nsem_exchange_state_halos(ctx)) throw Error(nsem_last_error(ctx));
}

void EulerSolver::upload_state() {
    if (nsem_upload_state(ctx, rho.data(), U.data(), T.data(), p.data())) throw Error(nsem_last_error(ctx));
}
void EulerSolver::upload_state_async() {
    if (!ctx) throw Error("EulerSolver::upload_state_async: no device attached");
    if (nsem_upload_state_async(ctx, rho.data(), U.data(), T.data(), p.data())) throw Error(nsem_last_error(ctx));
}
void EulerSolver::download_async() {
    if (!ctx) throw Error("EulerSolver::download_async: no device attached");
    if (out_rho.size() != rho.size()) {
        out_rho.assign(rho.size(), 0.0); out_U.assign(U.size(), 0.0); out_T.assign(T.size(), 0.0); out_p.assign(p.size(), 0.0);
        for (std::vector<double>* v : {&out_rho, &out_U, &out_T, &out_p}) nsem_pin_host(ctx, v->data(), v->size() * sizeof(double));
    }
    if (nsem_download_state_async(ctx, out_rho.data(), out_U.data(), out_T.data(), out_p.data())) throw Error(nsem_last_error(ctx));
}
void EulerSolver::sync() {
    if (ctx && nsem_sync(ctx)) throw Error(nsem_last_error(ctx));
}
void EulerSolver::adopt_refined_state(EulerSolver& old, const std::vector<u32>& refineMap, const std::vector<u32>& coarseMap,
                                      const std::vector<u32>& cellMap, bool restart) {
    if (!ctx || !old.ctx) throw Error("EulerSolver::adopt_refined_state: both solvers must be attached to the device (there is no CPU fallback)");
    for (int d = 0; d < 3; d++)
        if (nop[d] != old.nop[d]) throw Error("EulerSolver::adopt_refined_state: polynomial orders differ");
    Basis b(nop);
    // gCV / gCC as std::vector<Vec3> are contiguous triples
    nsem_regrid g;
    std::memset(&g, 0, sizeof g);
    g.n_cells_new = geo.nBCS;
    g.refine_map = refineMap.data(); g.n_refine_map = (u32)refineMap.size();
    g.coarse_map = coarseMap.data(); g.n_coarse_map = (u32)coarseMap.size();
    g.cell_map = cellMap.data(); g.n_cell_map = (u32)cellMap.size();
    g.old_cV = old.topo.CV.data(); g.old_cC = old.topo.CC.data()->data();
    g.new_cV = topo.CV.data(); g.new_cC = topo.CC.data()->data();
    g.old_node_cC = old.geo.cC.data();
    for (int q = 0; q < 6; q++) { g.psi_ref[q] = b.psiRef[q].data(); g.psi_cor[q] = b.psiCor[q].data(); }
    if (nsem_refine_state(old.ctx, &g, ctx)) throw Error(nsem_last_error(ctx));
    if (restart && nsem_restart_state(ctx)) throw Error(nsem_last_error(ctx));
}
void EulerSolver::restart_state() {
    if (!ctx) throw Error("EulerSolver::restart_state: no device attached (there is no CPU fallback)");
    if (nsem_restart_state(ctx)) throw Error(nsem_last_error(ctx));
}
void EulerSolver::step(int n) {
    if (!ctx) throw Error("EulerSolver::step: no device attached (there is no CPU fallback)");
    if (convection ? nsem_convection_step(ctx, n) : nsem_euler_step(ctx, n)) throw Error(nsem_last_error(ctx));
}
void EulerSolver::download() {
    if (nsem_download_state(ctx, rho.data(), U.data(), T.data(), p.data())) throw Error(nsem_last_error(ctx));
}

void EulerSolver::write_fields(int index) {
    const std::string s = std::to_string(index);
    const uint64_t n = geo.gBCSfield;
    std::string out = dir;
    if (nranks > 1) {
        out = dir + "/grid" + std::to_string(rank);
        ::mkdir(out.c_str(), 0777);
        ::unlink((out + "/cells" + s).c_str());          // a marker of an earlier run must not vouch for files that are being rewritten
    }
    if (convection) {
        write_field(out + "/T" + s, binary_out, 1, rho.data(), n, bc_rho);          // the scalar lives in the rho slot
        write_field(out + "/U" + s, binary_out, 3, U.data(), n, bc_U);
    } else {
        write_field(out + "/rho" + s, binary_out, 1, rho.data(), n, bc_rho);
        write_field(out + "/U" + s, binary_out, 3, U.data(), n, bc_U);
        write_field(out + "/T" + s, binary_out, 1, T.data(), n, bc_T);
        write_field(out + "/p" + s, binary_out, 1, p.data(), n, bc_p);
    }
    if (nranks > 1) {
        // local real cell -> global cell, then the marker rank 0's merge waits for (written last, renamed into place)
        FILE* f = std::fopen((out + "/cells" + s + ".tmp").c_str(), "wb");
        if (!f) throw Error("cannot write " + out + "/cells" + s);
        const u32 nc = (u32)cellGlobal.size();
        std::fwrite(&launch_nonce, sizeof launch_nonce, 1, f);      // this launch's stamp (euler_main.cpp: share_launch_blob)
        std::fwrite(&nc, sizeof nc, 1, f);
        std::fwrite(cellGlobal.data(), sizeof(u32), nc, f);
        std::fclose(f);
        if (std::rename((out + "/cells" + s + ".tmp").c_str(), (out + "/cells" + s).c_str()) != 0) throw Error("cannot rename the cell map in " + out);
    }
}

void EulerSolver::write_vtk(int index) const {
    std::string out = dir;
    if (nranks > 1) {
        out = dir + "/grid" + std::to_string(rank);      // prepare -vtk works inside grid<rank>/ too (prepareApp.cpp:104-109)
        ::mkdir(out.c_str(), 0777);
    }
    std::vector<VtkField> fl;
    for (const auto& name : vtk_fields) {
        if (convection) {                                  // the scalar T lives in the rho slot; rho and p are not fields of that app
            if (name == "T") fl.push_back({name, 1, rho.data()});
            else if (name == "U") fl.push_back({name, 3, U.data()});
            continue;
        }
        if (name == "rho") { if (!rho_file_missing) fl.push_back({name, 1, rho.data()}); }
        else if (name == "U") fl.push_back({name, 3, U.data()});
        else if (name == "T") fl.push_back({name, 1, T.data()});
        else if (name == "p") fl.push_back({name, 1, p.data()});
        // a name without a field is skipped like a name without a file is (Prepare::createFields, field.cpp:556-590)
    }
    nsemh::write_vtk(out + "/" + meshName + std::to_string(index) + ".vtk", Basis(nop), geo.cC.data(), geo.nBCS, fl, vtk_cell_value, vtk_polyhedral);
}

void EulerSolver::merge_fields(int index) {
    if (nranks <= 1 || rank != 0) return;
    const std::string s = std::to_string(index);
    const int NP = Basis(nop).NP;
    // the convection app dumps its scalar T (kept in the rho slot, conditions under bc_rho) and the wind
    const char* names[4] = {convection ? "T" : "rho", "U", "T", "p"};
    const int comps[4] = {1, 3, 1, 1};
    const std::vector<BCond>* bcs[4] = {&bc_rho, &bc_U, &bc_T, &bc_p};
    const int nfields = convection ? 2 : 4;
    std::vector<std::vector<u32>> maps(nranks);
    for (int r = 0; r < nranks; r++) {
        const std::string path = dir + "/grid" + std::to_string(r) + "/cells" + s;
        // the marker is written last and carries this launch's nonce: a marker an earlier run left behind is not this dump
        FILE* f = nullptr;
        for (int tries = 0; tries < 6000; tries++) {                                                             // <= 10 min
            if ((f = std::fopen(path.c_str(), "rb"))) {
                uint64_t stamp = 0;
                if (std::fread(&stamp, sizeof stamp, 1, f) == 1 && stamp == launch_nonce) break;
                std::fclose(f);
                f = nullptr;
            }
            ::usleep(100000);
        }
        if (!f) throw Error("merge: rank " + std::to_string(r) + " never wrote " + path + " in this launch");
        u32 nc = 0;
        if (std::fread(&nc, sizeof nc, 1, f) != 1) { std::fclose(f); throw Error("merge: short read of " + path); }
        maps[r].resize(nc);
        const size_t got = std::fread(maps[r].data(), sizeof(u32), nc, f);
        std::fclose(f);
        if (got != nc) throw Error("merge: short read of " + path);
    }
    for (int q = 0; q < nfields; q++) {
        const size_t per = (size_t)NP * comps[q];
        std::vector<double> all((size_t)nGlobalCells * per, 0.0);
        for (int r = 0; r < nranks; r++) {
            const FieldFile ff = read_field(dir + "/grid" + std::to_string(r) + "/" + names[q] + s, comps[q]);
            if (ff.values.size() < maps[r].size() * per) throw Error(std::string("merge: ") + names[q] + s + " of rank " + std::to_string(r) + " is too short");
            for (size_t l = 0; l < maps[r].size(); l++)
                std::copy(ff.values.begin() + l * per, ff.values.begin() + (l + 1) * per, all.begin() + (size_t)maps[r][l] * per);
        }
        std::vector<BCond> gb;      // the physical patches; the interMesh_* GHOST conditions exist only inside a partition
        for (const auto& b : *bcs[q]) if (b.type != "GHOST") gb.push_back(b);
        write_field(dir + "/" + names[q] + s, binary_out, comps[q], all.data(), (uint64_t)nGlobalCells * NP, gb);
    }
}

void EulerSolver::run() {
    // Iteration (iteration.h:18-84): steps start_step*write_interval+1 .. end_step, dump when i % write_interval == 0
    long i = write_interval * start_step + 1;
    arm_wind(i);
    const bool diag = !convection && (nranks == 1 || std::getenv("NSEM_DIAGNOSTICS"));
    double m0 = mass0, e0 = energy0, v0 = volume0;
    if (diag && nranks > 1 && i <= end_step) {          // the set-up's totals are those of this partition: take the global ones from the device
        double dg[6];
        if (nsem_diagnostics(ctx, dg)) throw Error(nsem_last_error(ctx));
        m0 = dg[3]; e0 = dg[4]; v0 = dg[5];
    }
    while (i <= end_step) {
        long next_dump = ((i + write_interval - 1) / write_interval) * write_interval;
        long upto = std::min(next_dump, end_step);
        step((int)(upto - i + 1));
        i = upto + 1;
        // the lines the reference prints when its print timer fires and on the last step (euler.cpp:260-283, iteration.h:68-70): here at every
        // dump and at the end.  On several partitions the sums are an ncclAllReduce every rank must join: opt-in (NSEM_DIAGNOSTICS=1)
        if ((upto % write_interval == 0 || upto == end_step) && diag) {
            double dg[6];
            if (nsem_diagnostics(ctx, dg)) throw Error(nsem_last_error(ctx));
            if (rank == 0) {
                std::printf("Time %f\n", upto * dt);
                std::printf("Courant number: Max: %g Min: %g Avg: %g\n", dg[0], dg[1], dg[2]);
                std::printf("Mass loss: %.12g Energy loss %.12g Volume loss %.12g\n", (m0 - dg[3]) / m0, (e0 - dg[4]) / e0, (v0 - dg[5]) / v0);
            }
        }
        if (upto % write_interval == 0) {
            download();
            if (convection && nranks == 1) {             // convection.cpp:139-149
                double sc = 0;
                for (uint64_t q = 0; q < geo.gBCSfield; q++) sc += rho[q] * geo.cV[q];
                std::printf("Time %f\nScalar loss: %.12g Volume loss: %.12g\n", upto * dt, (scalar0 - sc) / scalar0, 0.0);
            }
            write_fields((int)(upto / write_interval));
            merge_fields((int)(upto / write_interval));
            if (vtk_on_dump) write_vtk((int)(upto / write_interval));
            if (rank == 0) std::printf("Time %f : wrote fields %ld\n", upto * dt, upto / write_interval);
        }
    }
    if (nsem_sync(ctx)) throw Error(nsem_last_error(ctx));
}


// AmrIteration (iteration.h:94-147) around Iteration: with amr_step != 0 a regrid before the first step and after every amr_step dumps.
// The solver object is replaced by the one on the regridded mesh; the grid of every regrid is written as <mesh>_<dump>.txt.
// The AMR cycle on several partitions.  The reference merges the fields on rank 0, refines there and decomposes again (Prepare::mergeFields,
// refineMesh, decomposeMesh: field.cpp:1086-1496).  Here every rank keeps, next to its partition, a whole-domain solver on its own GPU that
// exists for the regrids only: at a regrid the parts' states are summed into it (nsem_allreduce_host: every rank fills its own cells), it
// regrids exactly like a one-partition run (same tags, same forest, field transfer + restart branch on the device), and every rank cuts its
// new part from the emitted grid (mortar-weighted METIS: non-conforming faces are never cut) and takes its cells' values.  All ranks compute
// the same regrid, so nothing but the state and a fresh communicator id crosses between them.
static void run_case_partitioned(std::unique_ptr<EulerSolver>& s) {
    const int rank = s->rank, nranks = s->nranks, device = s->device_id;
    const bool verbose = std::getenv("NSEM_VERBOSE") != nullptr;
    if (!s->ctx) throw Error("run_case: the partition is not attached to a device");
    const int NP = Basis(s->nop).NP;
    // the whole-domain solver: the case as one partition would set it up (no dumps, no prints)
    std::unique_ptr<EulerSolver> G(new EulerSolver());
    G->rank = 0; G->nranks = 1;
    G->read_controls(s->dir);
    G->launch_nonce = s->launch_nonce;
    G->keep_regrid_grid = true;
    G->load_mesh((int)G->start_step);
    G->read_fields((int)G->start_step);
    G->setup();
    if (!G->forest) throw Error("run_case: the grid of dump " + std::to_string(G->start_step) + " cannot seed the AMR forest (it must be conforming, or have its .forest file)");
    G->attach_device(device);
    auto regrid = [&](long dump) {
        // 1. whole-domain state from the parts
        s->download();
        const uint64_t nG = (uint64_t)s->nGlobalCells * NP;
        if (nG != G->geo.gBCSfield) throw Error("run_case: the partitions and the whole-domain solver are not on the same grid");
        std::vector<double> all(nG * 6, 0.0);
        for (size_t l = 0; l < s->cellGlobal.size(); l++)
            for (int q = 0; q < NP; q++) {
                const uint64_t src = (uint64_t)l * NP + q, dst = ((uint64_t)s->cellGlobal[l] * NP + q) * 6;
                all[dst] = s->rho[src];
                for (int d = 0; d < 3; d++) all[dst + 1 + d] = s->U[src * 3 + d];
                all[dst + 4] = s->T[src];
                all[dst + 5] = s->p[src];
            }
        if (nsem_allreduce_host(s->ctx, all.data(), all.size(), 0)) throw Error(nsem_last_error(s->ctx));
        for (uint64_t i = 0; i < nG; i++) {
            G->rho[i] = all[i * 6];
            for (int d = 0; d < 3; d++) G->U[i * 3 + d] = all[i * 6 + 1 + d];
            G->T[i] = all[i * 6 + 4];
            G->p[i] = all[i * 6 + 5];
        }
        G->apply_bcs(G->rho, 1, G->bc_rho); G->apply_bcs(G->U, 3, G->bc_U); G->apply_bcs(G->T, 1, G->bc_T); G->apply_bcs(G->p, 1, G->bc_p);
        G->upload_state();
        // 2. the one-partition regrid: tags, forest, new mesh, device transfer + restart branch
        std::unique_ptr<EulerSolver> Gn = G->regridded_by_indicator();
        Gn->download();
        if (rank == 0) {
            Gn->write_amr_grid(dump);
            std::printf("Regrid at dump %ld: %u -> %u cells\n", dump, G->geo.nBCS, Gn->geo.nBCS);
        }
        // 3. this rank's part of the new grid with its cells' values
        std::unique_ptr<EulerSolver> n(new EulerSolver());
        s->copy_run_parameters(*n);
        n->set_mesh_partition(*Gn->regrid_grid, rank, nranks, n->decomp_type, n->decomp_n);
        auto blank = [](int comps, const std::vector<BCond>& bcs) {
            FieldFile f;
            f.comps = comps;
            f.inits.push_back({"uniform", std::vector<double>(comps, 0.0)});
            f.bcs = bcs;
            return f;
        };
        n->set_fields(blank(1, s->file_bc_rho), blank(3, s->file_bc_U), blank(1, s->file_bc_T), blank(1, s->file_bc_p));
        n->setup();
        n->mass0 = s->mass0; n->energy0 = s->energy0; n->volume0 = s->volume0;
        for (size_t l = 0; l < n->cellGlobal.size(); l++)
            for (int q = 0; q < NP; q++) {
                const uint64_t dst = (uint64_t)l * NP + q, src = (uint64_t)n->cellGlobal[l] * NP + q;
                n->rho[dst] = Gn->rho[src];
                for (int d = 0; d < 3; d++) n->U[dst * 3 + d] = Gn->U[src * 3 + d];
                n->T[dst] = Gn->T[src];
                n->p[dst] = Gn->p[src];
            }
        n->apply_bcs(n->rho, 1, n->bc_rho); n->apply_bcs(n->U, 3, n->bc_U); n->apply_bcs(n->T, 1, n->bc_T); n->apply_bcs(n->p, 1, n->bc_p);
        // 4. a communicator for the new parts: rank 0 draws the id, the old communicator carries it
        unsigned char id[128];
        std::memset(id, 0, sizeof id);
        if (rank == 0 && nsem_get_unique_id(id)) throw Error(nsem_last_error(nullptr));
        if (nsem_allreduce_host(s->ctx, id, sizeof id, 1)) throw Error(nsem_last_error(s->ctx));
        n->attach_device(device, rank, nranks, id);
        if (verbose) std::printf("regrid[%d]: part of %zu cells, %zu neighbours\n", rank, n->cellGlobal.size(), n->peers.size());
        s = std::move(n);
        G = std::move(Gn);
        if (std::getenv("NSEM_DEBUG_REGRID")) { s->write_fields(100 + (int)dump); s->merge_fields(100 + (int)dump); }
    };
    const long last = s->end_step;
    s->conv_end_step = last; G->conv_end_step = last;
    long dump = s->start_step;
    regrid(dump);
    while (dump * s->write_interval < last) {
        const long upto = std::min(last, (dump + s->amr_step) * s->write_interval);
        s->start_step = dump;
        s->end_step = upto;
        s->conv_end_step = last;
        s->run();
        s->end_step = last;
        dump = upto / s->write_interval;
        if (upto < last) regrid(dump);
    }
}

void run_case(std::unique_ptr<EulerSolver>& s) {
    if (s->amr_step == 0) { s->run(); return; }
    if (s->nranks > 1) { run_case_partitioned(s); return; }
    auto regrid = [&](long dump) {
        std::unique_ptr<EulerSolver> n = s->regridded_by_indicator();
        const u32 before = s->geo.nBCS;
        s = std::move(n);
        s->write_amr_grid(dump);
        std::printf("Regrid at dump %ld: %u -> %u cells\n", dump, before, s->geo.nBCS);
        if (std::getenv("NSEM_DEBUG_REGRID")) { s->download(); s->write_fields(100 + (int)dump); }
    };
    const long last = s->end_step;
    s->conv_end_step = last;
    long dump = s->start_step;
    regrid(dump);
    while (dump * s->write_interval < last) {
        const long upto = std::min(last, (dump + s->amr_step) * s->write_interval);
        s->start_step = dump;
        s->end_step = upto;
        s->conv_end_step = last;
        s->run();
        s->end_step = last;
        dump = upto / s->write_interval;
        if (upto < last) regrid(dump);
    }
}

}  // namespace nsemh
