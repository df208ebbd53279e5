// nsem_host.h -- host side of the drop-in: what NebulaSEM's `euler` app does around its time loop.
//
// Mirrors the reference's host surface for the explicit Euler path (names follow the reference):
//   Util::read_params / ParamList      src/util/util.cpp:32-79, util.h:343-393   -> Controls
//   MeshObject::readTextMesh           src/mesh/mesh.h:158-200                   -> Grid, read_grid
//   Mesh::LoadMesh                     src/field/field.cpp:95-167                -> MeshTopo::load
//   DG::init_poly/init_basis/init_geom src/field/dg.cpp:147-590                  -> Basis, Geometry
//   MeshField::read/write, BCondition  src/field/field.h:144-267,1412-1677       -> FieldFile, read_field, write_field
//   euler() set-up + time loop         apps/euler/euler.cpp:17-293               -> EulerSolver
//   Iteration / AmrIteration           src/solvers/iteration.h                   -> EulerSolver::run
// The compute bodies are NOT here: the time-loop body is nsem_euler_step() of include/nsem_c.h (CUDA).
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/nsem_c.h"

namespace nsemh {

using u32 = uint32_t;
using Vec3 = std::array<double, 3>;
constexpr u32 MAX_INT = 1u << 31;        // Constants::MAX_INT (tensor.h:455)

struct Error : std::exception {
    std::string msg;
    explicit Error(std::string m) : msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

// ---- token streams (text and binary share one grammar, util.h:191-256) --------------------------------
class Tokens {
public:
    static Tokens open(const std::string& path_noext);            // .txt first, then .bin (field.cpp:108-117)
    static Tokens from_text(const std::string& text);
    static Tokens from_binary(std::vector<unsigned char> bytes);
    std::string word();
    u32 uint();
    double real();
    void sym() { word(); }
    bool eof() const;
    bool binary = false;

private:
    std::vector<std::string> tok_;
    size_t pos_ = 0;
    std::vector<unsigned char> bin_;
};

// ---- controls -------------------------------------------------------------------------------------------
struct Controls {
    std::map<std::string, std::map<std::string, std::vector<std::string>>> blocks;
    static Controls read(const std::string& path);
    bool has(const std::string& blk, const std::string& key) const;
    std::string str(const std::string& blk, const std::string& key, const std::string& def) const;
    double num(const std::string& blk, const std::string& key, double def) const;
    long integer(const std::string& blk, const std::string& key, long def) const;
    bool yes(const std::string& blk, const std::string& key, bool def) const;
    Vec3 vec(const std::string& blk, const std::string& key, Vec3 def) const;
};

// ---- grid -------------------------------------------------------------------------------------------------
struct Grid {
    std::vector<Vec3> V;
    std::vector<u32> facetStart{0}, facetVerts;      // CSR
    std::vector<u32> cellStart{0}, cellFaces;        // CSR, real cells
    std::map<std::string, std::vector<u32>> boundaries;   // std::map: iteration in name order like the reference
    u32 nFacets() const { return (u32)facetStart.size() - 1; }
    u32 nCells() const { return (u32)cellStart.size() - 1; }
};
Grid read_grid(const std::string& path_noext);
void write_grid_text(const std::string& path, const Grid& g);
void write_grid_binary(const std::string& path, const Grid& g);
// structured box of n cells on [lo,hi]; sides x-,x+,y-,y+,z-,z+ -> patch names; optional terrain map
Grid box_grid(const int n[3], const double lo[3], const double hi[3], const std::array<std::string, 6>& patches,
              void (*vertex_map)(Vec3&, const void*) = nullptr, const void* map_arg = nullptr);

// ---- domain decomposition (Prepare::decomposeMesh, field.cpp:1086-1257) --------------------------------------
// face_mortar (gFMC in the grid's own face numbering, or nullptr): METIS edge weight 1000 on non-conforming faces so that
// AMR families stay together (field.cpp:1037-1047); any method that still cuts one is refused (field.cpp:1215-1220)
std::vector<u32> partition_cells(const Grid& g, int nparts, const std::string& method, const int nxyz[3],
                                 const std::vector<u32>* face_mortar = nullptr, const std::vector<std::array<u32, 2>>* together = nullptr);
// gFMC of a grid in its own face numbering (addBoundaryCells + fixHexCells only); all zero on a conforming grid
std::vector<u32> mortar_flags(const Grid& g);
struct Partition {
    Grid grid;                      // local grid with interMesh_<me>_<peer> patches
    std::vector<u32> cellGlobal;    // local real cell -> global cell
    std::vector<int> peers;         // neighbouring ranks, ascending
};
Partition extract_partition(const Grid& g, const std::vector<u32>& part, int rank, int nparts);

// ---- topology + element geometry (mesh.cpp:55-109, 113-157, 161-446, 450-577, 581-669) ------------------
struct MeshTopo {
    std::vector<Vec3> V;
    std::vector<u32> facetStart, facetVerts;
    std::vector<u32> cellStart, cellFaces, cellFaceID;   // all cells (real + boundary); faceID aligned with cellFaces
    std::map<std::string, std::vector<u32>> boundaries;
    u32 nBCS = 0;
    std::vector<u32> FOC, FNC, FMC;
    std::vector<Vec3> FC, FN, CC;
    std::vector<double> CV;
    // Mesh::is_spherical / sphere_radius / sphere_height (mesh.cpp:31-33): set before load()
    bool spherical = false;
    double sphere_radius = 6371220.0, sphere_height = 10000.0;
    // the two cubes of the WHOLE grid (smallest and largest max|coordinate| over its vertices); when set (shell_h[1] > 0) extrude() uses
    // them instead of this mesh's own extremes, so that a partition that does not reach both shells is projected like the whole mesh
    double shell_h[2] = {0, 0};
    bool keep_empty = false;             // LoadMesh(step, remove_empty = false), what Prepare::convertVTK loads: the `delete` patch stays
    bool no_extrude = false;             // LoadMesh(step, first, extrude = false): the cube shell as the grid file has it (the AMR tagging's volumes)
    u32 nFacets() const { return (u32)facetStart.size() - 1; }
    u32 nCells() const { return (u32)cellStart.size() - 1; }
    void load(const Grid& g);            // Mesh::LoadMesh
    void load_flags_only(const Grid& g); // addBoundaryCells + fixHexCells: FOC/FNC/FMC in the grid's own numbering
    void hex_corners(const u32* f1, const u32* f2, u32 out[8]) const;   // quads only
    void hex_corners_poly(const std::vector<u32>& f1, const std::vector<u32>& f2, u32 out[8]) const;   // merged sides
    std::vector<u32> merged_side(u32 cell, u32 id) const;               // polygon of all facets of `cell` with local id `id`

private:
    void add_boundary_cells();
    void fix_hex_cells();
    void fix_general_cell(u32 ci);       // non-conforming cell: coplanar sub-facets grouped and merged per side
    void calc_geometry();
    void extrude();                      // ExtrudeMesh: the cube shell of the grid file projected onto the sphere
    void sphere_geometry();              // curved-element corrections of calcGeometry
    void remove_boundary(const std::vector<u32>& faces);
};

// ---- DG basis + node geometry ---------------------------------------------------------------------------------
struct Basis {
    int NPX = 1, NPY = 1, NPZ = 1, NP = 1, NPF = 1;
    std::vector<double> xgl[3], wgl[3], psi[3], dpsi[3];
    // 2:1 mortar projections per direction d and half h (dg.cpp:550-590): psiRef[d*2+h][in*n+io] coarse -> fine trace,
    // psiCor[d*2+h][io*n+in] fine -> coarse flux
    std::vector<double> psiRef[6], psiCor[6];
    explicit Basis(const int nop[3]);     // DG::Nop = polynomial degree per direction
    int n(int d) const { return d == 0 ? NPX : (d == 1 ? NPY : NPZ); }
};
void legendre_gauss_lobatto(int N, double* xgl, double* wgl);                               // dg.cpp:53-99
void lagrange_basis(int N, const double* xgl, int Ns, const double* xs, double* psi);      // dg.cpp:103-118
void lagrange_basis_derivative(int N, const double* xgl, int Ns, const double* xs, double* dpsi);   // dg.cpp:122-143

struct Geometry {
    u32 nBCS = 0, nCells = 0, nFacets = 0;
    uint64_t gBCSfield = 0, gALL = 0;
    bool spherical = false;               // copied from the topology by build()
    double sphere_radius = 0;
    std::vector<double> cC, cV, Jinv, fN, fC, fI, faceNormal, faceCenter;   // AoS like the reference
    std::vector<double> psiRef[6], psiCor[6];                   // copies of the basis tables (only read on non-conforming meshes)
    std::vector<u32> FO, FN, faceBegin, faceEnd, allFaces, faceID, faceOwner, faceNeigh, faceMortar;
    void build(const MeshTopo& t, const Basis& b);   // initGeomMeshFields + init_geom (fI = 0.5 on interMesh_* faces)
    nsem_mesh as_c() const;
};

// ---- field files -------------------------------------------------------------------------------------------------
struct BCond {                       // BCondition<T>, field.h:144-173
    std::string patch, type, neighbor;
    double value[3] = {0, 0, 0}, tvalue[3] = {0, 0, 0};
    double shape = 0, tshape = 0, zMin = 0, zMax = 0;
    Vec3 dir{0, 0, 1};
    std::vector<double> fixed;       // frozen CALC_DIRICHLET values [nfaces*NPF*comps]
    bool held = false;               // not from the field file: added by mark_unlisted_patches for a patch rho has no condition for
};
struct FieldFile {
    int comps = 1;
    struct Init { std::string kind; std::vector<double> a; };
    std::vector<Init> inits;          // `internal N<=4` analytic initialisers (field.h:1424-1521)
    std::vector<double> values;       // otherwise raw node values
    std::vector<BCond> bcs;
};
FieldFile read_field(const std::string& path_noext, int comps_expected);
std::vector<double> init_field(const FieldFile& ff, const Geometry& g, const Vec3& gravity);
void write_field(const std::string& path_noext, bool binary, int comps, const double* v, uint64_t n_nodes,
                 const std::vector<BCond>& bcs);
// <mesh><step>.vtk as `prepare ./controls -vtk` writes it for a dGSEM case (Vtk::write_vtk, vtk.cpp:125-286): the LGL nodes as points, every
// element cut into (NPX-1)(NPY-1)(NPZ-1) sub-cells, cellID per sub-cell, the fields as point data (scalars first, then vectors)
struct VtkField { std::string name; int comps; const double* v; };
void write_vtk(const std::string& path, const Basis& b, const double* cC, u32 nBCS, const std::vector<VtkField>& fields,
               bool write_cell_value = true, bool write_polyhedral = false);

// ---- adaptive regrid in memory (amr.cpp): Prepare::refineMesh tagging + MeshObject::refineMesh -----------------------
struct RefineParams {                // refinement{} of the controls (Controls::enrollRefine, field.cpp:474-482; defaults field.h:18-40)
    Vec3 dir{0, 0, 0};
    std::string field = "U";
    double field_max = 0.6, field_min = 0.2;
    int max_level = 1, buffer_zone = 2;
    long limit = 100000;
};
struct AmrForest {
    struct Node {
        u32 v[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // corners c0..c7 <-> (xi,eta,zeta) bits 000 100 110 010 001 101 111 011
        int level = 0, parent = -1, child0 = -1, nchild = 0;
    };
    struct Maps { std::vector<u32> refineMap, coarseMap, cellMap; };    // as MeshObject::refineMesh returns them (mesh.cpp:2216-2748)
    std::vector<Vec3> V;
    std::vector<Node> nodes;
    std::vector<int> leaves;                   // node of every cell of the current grid, in cell order
    std::map<std::array<u32, 2>, u32> edgeMid;
    std::map<std::array<u32, 4>, u32> faceMid;
    std::map<std::array<u32, 4>, std::string> patchOf;    // boundary quads of every level -> patch name
    Vec3 dir{0, 0, 0};
    bool spherical = false;                    // cubed sphere: the axis a 2-D refinement never splits is the radial one of each cell
    void init(const Grid& conforming_hex_grid, const Vec3& direction);
    Maps regrid(const std::vector<uint8_t>& refine, const std::vector<uint8_t>& coarsen);   // one flag per current cell
    Grid grid() const;                         // the current grid in the reference's format
    std::vector<int> levels() const;
    std::vector<std::vector<u32>> families() const;
    // the forest next to the grid it emitted (<mesh>_<dump>.forest), so that an AMR run can be restarted from a dump; own binary format
    // (the reference keeps amrTree_<k>.bin, a different tree over a different cell order)
    void save(const std::string& path) const;
    void load(const std::string& path);
private:
    int split_mask(const Node& n) const;
    u32 mid_vertex(const std::vector<u32>& of);
    void split(int node);
};
struct EulerSolver;
// Prepare::refineMesh's tagging (field.cpp:606-620, 696-824): normalised indicator sqrt(|field|) (cV^0.125 / max), element means against
// field_max / field_min, buffer zone, limits, whole families only, 2:1 balance across faces.  `levels` = refinement level per cell.
void amr_tag_cells(const EulerSolver& s, const RefineParams& rp, const std::vector<int>& levels, const std::vector<std::vector<u32>>& families,
                   std::vector<uint8_t>& refine, std::vector<uint8_t>& coarsen, const std::vector<double>* node_volumes = nullptr);

// ---- the solver app ------------------------------------------------------------------------------------------------
struct EulerSolver {
    Controls ctl;
    std::string dir, meshName = "grid";
    int nop[3] = {0, 0, 0};
    // general{} / euler{} (properties.cpp:14-34, euler.cpp:19-48, field.cpp:56-74)
    double viscosity = 1.568e-5, Pr = 0.9, T0 = 300, P0 = 101325, cp = 1004.67, cv = 715.5, dt = 0.1;
    Vec3 gravity{0, 0, -9.860616};
    bool buoyancy = true, diffusion = true, binary_out = true;
    std::string time_scheme = "BDF1", problem_init = "NONE";
    long start_step = 0, end_step = 2, write_interval = 20;   // start_step: the DUMP the run starts from (controls' start_step / write_interval)
    // decomposition{type n} (field.cpp:486-492): METIS | XYZ (n = parts per axis) | CELLID
    std::string decomp_type = "METIS";
    int decomp_n[3] = {1, 1, 1};

    MeshTopo topo;
    Geometry geo;
    std::vector<double> rho, U, T, p, rho_ref, p_ref, gvec, gh;
    std::vector<BCond> bc_rho, bc_U, bc_T, bc_p, bc_rho_ref, bc_p_ref, bc_g;
    double mass0 = 0, energy0 = 0, volume0 = 0;
    // `solver convection` (apps/convection/convection.cpp: dT/dt + div(T U) = 0): the transported scalar is kept where the euler solver keeps
    // rho (members rho / bc_rho, file T<k>), the wind in U; problem_init LEVEQUE re-evaluates the wind on the device at every step
    bool convection = false;
    std::string conv_init = "NONE";
    double scalar0 = 0;
    uint64_t launch_nonce = 0;                  // stamp of this launch on the per-dump markers (euler_main.cpp: share_launch_blob)

    nsem_ctx* ctx = nullptr;
    ~EulerSolver();

    void read_controls(const std::string& case_dir);      // Solver::Initialize + euler{} params
    void load_mesh(int step);                             // Mesh::LoadMesh from <mesh>_<step>.{txt,bin}
    void set_mesh(const Grid& g);                         // same, from memory
    // decompose `global` into nranks parts (type METIS | XYZ | CELLID) and keep part `rank` (decomposeMesh)
    void set_mesh_partition(const Grid& global, int rank, int nranks, const std::string& type, const int nxyz[3]);
    int rank = 0, nranks = 1;
    u32 nGlobalCells = 0;                                 // real cells of the undecomposed grid
    std::vector<u32> cellGlobal;                          // local real cell -> global cell (identity on 1 rank)
    std::vector<int> peers;
    void exchange_setup_halos();                          // the applyExplicitBCs(...,true) halos of euler.cpp:105-146
    void read_fields(int step);                           // Mesh::read_fields
    void set_fields(const FieldFile& frho, const FieldFile& fU, const FieldFile& fT, const FieldFile& fp);
    void setup();                                         // euler.cpp:58-176 (reference state, rho from p, BCs)
    void attach_device(int device, int rank = 0, int nranks = 1, const void* uid = nullptr);   // C-ABI uploads
    void upload_state();
    // pipelined batches (nsem_upload_state_async / nsem_download_state_async): the download lands in out_* so that the next upload can
    // read rho/U/T/p while it is in flight; results are complete after sync()
    std::vector<double> out_rho, out_U, out_T, out_p;
    void upload_state_async();
    void download_async();
    void sync();
    void step(int n);                                     // time-loop body on the GPU
    // AMR regrid with the state resident on the device (the field transfer of Prepare::refineMesh, field.cpp:884-906): `old` holds the
    // mesh before the regrid and the state, this solver the regridded mesh (attached to the same device); the maps are those of
    // MeshObject::refineMesh.  nsem_refine_state, then (restart) the set-up's restart branch nsem_restart_state (euler.cpp:150-162).
    void adopt_refined_state(EulerSolver& old, const std::vector<u32>& refineMap, const std::vector<u32>& coarseMap,
                             const std::vector<u32>& cellMap, bool restart);
    void restart_state();                                 // nsem_restart_state: p from rho, ghost cells from the resident state
    // AMR in memory (AmrIteration::next, iteration.h:124-141, without the files): the forest persists across regrids
    std::shared_ptr<AmrForest> forest;
    RefineParams refine_params;
    long amr_step = 0;
    AmrForest::Maps last_maps;                            // maps of the regrid that produced this solver's mesh
    std::vector<BCond> file_bc_rho, file_bc_U, file_bc_T, file_bc_p;   // boundary conditions as the field files state them
    int device_id = -1;
    std::string forest_file;                              // set by load_mesh: <mesh>_<step>.forest, read by set_mesh when it exists
    // the solver on the regridded mesh: same controls and boundary conditions, mesh from the forest after regrid(refine, coarsen), set-up
    // done; when this solver is attached the new one is attached to the same device and takes the state (adopt_refined_state, restart)
    std::unique_ptr<EulerSolver> regridded(const std::vector<uint8_t>& refine, const std::vector<uint8_t>& coarsen);
    std::unique_ptr<EulerSolver> regridded_by_indicator();   // amr_tag_cells on the downloaded state, then regridded()
    void copy_run_parameters(EulerSolver& n) const;       // controls, physics, AMR and output settings (not the mesh, fields or device)
    // regrid on several partitions (run_case): the whole-domain solver keeps the grid its regrid emitted so that the parts can be cut from it
    bool keep_regrid_grid = false;
    std::shared_ptr<Grid> regrid_grid;
    void write_amr_grid(long dump) const;                 // <mesh>_<dump>.txt + <mesh>_<dump>.forest in the case directory
    void download();
    void write_fields(int index);                         // Mesh::write_fields; with nranks > 1 into <case>/grid<rank>/ like the
                                                          // reference's per-rank working directories (field.cpp:1436-1440)
    void merge_fields(int index);                         // rank 0: grid<r>/<field><index> of all ranks -> <case>/<field><index>
                                                          // in global node order (Prepare::mergeFields, field.cpp:1446-1496)
    // Prepare::convertVTK (prepare.cpp:9-17) without the round trip through the field files: <mesh><index>.vtk from the state on the host
    // (after download()), fields and order from prepare{fields} of the controls, options from vtk{}; per rank into grid<rank>/ like the dumps
    std::vector<std::string> vtk_fields{"U", "T", "p", "rho"};
    bool vtk_cell_value = true, vtk_polyhedral = false, vtk_on_dump = false;
    void write_vtk(int index) const;
    void run();                                           // Iteration loop: steps + dumps every write_interval

    void apply_bcs(std::vector<double>& f, int comps, std::vector<BCond>& bcs);   // applyExplicitBCs on the host
    std::string conv_scheme = "RUSANOV";  // Controls::convection_scheme / blend_factor (field.cpp:56, 520-527)
    double blend_factor = 0.2;
    // (patch, neighbor) of the CYCLIC conditions of the case: a decomposition keeps the owner cells of paired faces in one part
    std::vector<std::array<std::string, 2>> cyclic_patches;
    bool rho_file_missing = false;        // read_fields found no rho<step> file (a name without a file is not converted, field.cpp:556-590)
    bool vtk_mode = false;                // `-vtk`: load meshes the way Prepare::convertVTK does (prepare.cpp:9-17)
    long conv_end_step = 0;               // the whole run's end_step while run_case shortens end_step to the next regrid (the wind's period)
    void arm_wind(long first_step);       // convection: nsem_set_convection with the step the next call starts with
    void mark_unlisted_patches();         // patches rho has no condition for (NSEM_BC_UNLISTED)
private:
    std::vector<nsem_bc> c_bcs_;
    std::vector<std::vector<u32>> keep_faces_;
    void build_c_bcs();
};

void run_case(std::unique_ptr<EulerSolver>& s);          // EulerSolver::run, with the AMR cycle around it when amr_step != 0

}  // namespace nsemh
