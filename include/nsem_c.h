/* nsem_c.h -- C ABI of libnsem_cuda.so: the B200 (sm_100a) replacement for the compute bodies of
 * NebulaSEM's explicit dGSEM Euler path.
 *
 * The reference has no plugin/FFI boundary (its operator API is C++ templates in src/field/field.h used
 * directly by apps/euler/euler.cpp).  This header IS the boundary the drop-in introduces: the host side
 * (nebulasem_b200/csrc/host, the `euler` app mirror) keeps the reference's surface and calls these entry
 * points instead of running the OpenMP/OpenACC loops.  Each entry cites the reference code it replaces
 * (file:line relative to the NebulaSEM tree).
 *
 * Conventions: every function returns 0 on success, non-zero on error (message via nsem_last_error);
 * host arrays are borrowed for the duration of the call; device memory is owned by the context; one host
 * thread per context; array layouts are the reference's (AoS Vector = 3, Tensor = 9 doubles per node in
 * XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX order, node index = cell*NP + i*NPY*NPZ + j*NPZ + k, src/field/dg.h:43-44;
 * all indices uint32 like `Int`, src/tensor/types.h:9).
 */
#ifndef NSEM_C_H
#define NSEM_C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsem_ctx nsem_ctx;

/* Boundary-condition kinds that reach the explicit path (src/field/field.h:129-141, 2674-2719). */
enum nsem_bc_kind {
    NSEM_BC_NONE = 0,
    NSEM_BC_NEUMANN = 1,      /* ghost = owner + value*|dx|, |dx| = 0 for DG ghost nodes (field.h:2674-2676) */
    NSEM_BC_DIRICHLET = 2,    /* ghost = value (field.h:2688-2689) */
    NSEM_BC_SYMMETRY = 3,     /* ghost = sym(owner, fN) (field.h:2681-2682, tensor.h:483-494) */
    NSEM_BC_CYCLIC = 4,       /* ghost = owner value on the paired patch face (field.h:2683-2687) */
    NSEM_BC_GHOST = 5,        /* inter-partition face: filled by the halo exchange (field.h:2605-2606, 2255-2324) */
    NSEM_BC_FIXED = 6,        /* frozen per-node values: CALC_DIRICHLET/POWER/LOG/PARABOLIC/INVERSE after their
                                 first evaluation (field.h:2691-2718) */
    NSEM_BC_ROBIN = 7,        /* ghost = shape*value + (1-shape)*(owner + tvalue*|dx|) (field.h:2677-2680) */
    NSEM_BC_UNLISTED = 8      /* rho only: the field file names no condition for the patch, so nothing overwrites what
                                 SolveTexplicit leaves in the boundary cell (solve.cpp:563-565): the owner's residual (fillBCs,
                                 field.h:2731-2743) over the boundary cell's own volume, i.e. ghost += (owner_new - owner_old) *
                                 cV_owner / cV_ghost; `fixed` holds that volume ratio per face node (examples/atmo/hydro-sphere
                                 ships no rho0 file) */
};

/* Field ids for BC tables and op-level calls. */
enum nsem_field { NSEM_F_RHO = 0, NSEM_F_P = 1, NSEM_F_U = 2, NSEM_F_T = 3, NSEM_F_COUNT = 4 };

/* Geometry + connectivity exactly as Mesh::initGeomMeshFields / DG::init_geom leave them
 * (src/field/field.cpp:171-280, src/field/dg.cpp:167-477). */
typedef struct {
    uint32_t n_cells_real;        /* gBCS */
    uint32_t n_cells_all;         /* gNCells = real + boundary ("ghost") cells */
    uint32_t n_faces;             /* gNFacets */
    const double* cV;             /* [n_cells_all*NP] nodal mass (dg.cpp:315-318) */
    const double* Jinv;           /* [n_cells_real*NP*9] (dg.cpp:413-476) */
    const double* fN;             /* [n_faces*NPF*3] weighted area vectors (dg.cpp:359) */
    const double* fI;             /* [n_faces*NPF] owner weight: 0.5 interior/inter-rank, 0 physical boundary
                                     (field.cpp:257-270) */
    const double* face_normal;    /* [n_faces*3] un-weighted gFN (mesh.cpp:476-502); fN = face_normal * w_a*w_b/4 */
    const uint32_t* FO;           /* [n_faces*NPF] owner node of each face node, sentinel n_cells_all*NP */
    const uint32_t* FN;           /* [n_faces*NPF] neighbour node */
    const uint32_t* face_begin;   /* [n_cells_all] faceIndices[0] (field.cpp:178-193) */
    const uint32_t* face_end;     /* [n_cells_all] faceIndices[1] */
    const uint32_t* all_faces;    /* [face_end[n_cells_all-1]] */
    const uint32_t* face_id;      /* [same] gFaceID flattened: local face id 0..5 (mesh.cpp:161-446) */
    const uint32_t* face_owner;   /* [n_faces] gFOC */
    const uint32_t* face_neigh;   /* [n_faces] gFNC */
    const uint32_t* face_mortar;  /* [n_faces] gFMC: 0 = conforming, 1 / 2 = non-conforming (2:1) sub-facet whose owner /
                                     neighbour cell is the fine one (mesh.cpp:431-444, mesh.h:89-92) */
    /* Only read when face_mortar has non-zero entries (scatter/gather_non_conforming, field.h:2019-2248): */
    const double* cC;             /* [n_cells_all*NP*3] node coordinates (dg.cpp:176-325): picks the half of the coarse face */
    const double* face_center;    /* [n_faces*3] gFC (mesh.cpp:476-502) */
    const double* psi_ref[6];     /* DG::psiRef[d*2+half][in*n+io], coarse -> fine trace (dg.cpp:576-590) */
    const double* psi_cor[6];     /* DG::psiCor[d*2+half][io*n+in], fine -> coarse flux */
} nsem_mesh;

/* One boundary patch's condition for one field (BCondition<T>, field.h:144-173). */
typedef struct {
    int32_t field;                /* enum nsem_field */
    int32_t kind;                 /* enum nsem_bc_kind */
    uint32_t n_faces;
    const uint32_t* faces;        /* patch face list (gBoundaries[name]) */
    const uint32_t* peer_faces;   /* CYCLIC: faces of the `neighbor` patch, same length (field.h:2662-2664) */
    double value[3];              /* value (scalar in [0]) */
    double shape;
    double tvalue[3];
    double tshape;
    double zMin;
    const double* fixed;          /* NSEM_BC_FIXED: [n_faces*NPF*comps]; NSEM_BC_UNLISTED: volume ratios [n_faces*NPF] */
} nsem_bc;

/* Scalars of general{} / euler{} that the step needs (apps/utils/properties.cpp:14-34, euler.cpp:19-48,
 * src/field/field.cpp:496-552). */
typedef struct {
    double P0, T0, cp, cv, viscosity, Pr;
    double gravity[3];
    double dt;
    int32_t buoyancy;             /* euler{buoyancy} */
    int32_t diffusion;            /* euler{diffusion} */
} nsem_params;

/* Inter-partition neighbour (Mesh::interBoundary, src/mesh/mesh.h:32-37). */
typedef struct {
    int32_t peer_rank;
    uint32_t n_faces;
    const uint32_t* faces;        /* faces of patch interMesh_<me>_<peer>, ascending global order */
} nsem_halo_peer;

/* ---- life cycle ------------------------------------------------------------------------------------ */
/* MP::MP / MP::~MP (src/mp/mp.cpp:17-35): one context per rank/GPU. nccl_unique_id is the 128-byte
 * ncclUniqueId shared by all ranks (NULL when nranks == 1). device < 0 selects rank % (visible devices). */
int nsem_create(int device, int rank, int nranks, const void* nccl_unique_id, nsem_ctx** out);
void nsem_destroy(nsem_ctx* ctx);
const char* nsem_last_error(const nsem_ctx* ctx);          /* ctx may be NULL: error of the last failed create */
/* In-place sum of a HOST array over the ranks of the context's communicator (MP::allreduce, src/mp/mp.h:93-104); dtype 0 = double,
 * 1 = unsigned byte.  Collective.  The host side of a regrid on several partitions uses it to assemble the whole-domain state (every rank
 * fills its own cells, zeros elsewhere) and to pass a fresh communicator id from rank 0 to everybody (Prepare::mergeFields +
 * decomposeMesh around refineMesh, src/field/field.cpp:1086-1496). */
int nsem_allreduce_host(nsem_ctx* ctx, void* buf, uint64_t count, int dtype);
int nsem_device(const nsem_ctx* ctx);                      /* the CUDA device ordinal nsem_create resolved (device < 0: rank % visible devices) */
int nsem_get_unique_id(void* out128);                      /* ncclGetUniqueId for rank 0 to broadcast */

/* ---- set-up ---------------------------------------------------------------------------------------- */
/* DG::init_poly (dg.cpp:147-163): points per direction = order+1 (npx+1, npy+1, npz+1). */
int nsem_set_order(nsem_ctx* ctx, int NPX, int NPY, int NPZ);
/* DG::init_basis tables (dg.cpp:481-547): dpsi[d][s*n+i] = l_i'(x_s), wgl[d][n]. */
int nsem_set_basis(nsem_ctx* ctx, const double* const dpsi[3], const double* const wgl[3]);
int nsem_upload_mesh(nsem_ctx* ctx, const nsem_mesh* mesh);
int nsem_set_bcs(nsem_ctx* ctx, const nsem_bc* bcs, uint32_t n_bcs);           /* AllBConditions of rho,p,U,T */
int nsem_set_halo(nsem_ctx* ctx, const nsem_halo_peer* peers, uint32_t n_peers); /* gInterMesh */
int nsem_set_params(nsem_ctx* ctx, const nsem_params* p);
/* Process elements in this order (cache blocking); NULL = mesh order. Results do not depend on it. */
int nsem_set_schedule(nsem_ctx* ctx, const uint32_t* order, uint32_t n);

/* ---- state ------------------------------------------------------------------------------------------ */
/* Fields over ALL nodes (n_cells_all*NP), ghost nodes included, as euler.cpp holds them between steps:
 * rho, U (AoS 3), T (perturbation theta - T0, euler.cpp:181,286), p (perturbation p - p_ref). */
int nsem_upload_state(nsem_ctx* ctx, const double* rho, const double* U, const double* T, const double* p);
int nsem_download_state(nsem_ctx* ctx, double* rho, double* U, double* T, double* p);
/* Operator-level view for unit parity: the results of gradf<strong>(U) and gradf<strong>(T) + fillBCs(r, fIndex) (field.h:3328-3362,
 * 2731-2769) of the last nsem_euler_step, per unit volume, ghost nodes included: grad_U as Tensor (9 per node, XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX
 * with G[ab] = d_a U_b), grad_T as Vector.  Only with diffusion on (euler.cpp:189-190); NULL arrays are skipped. */
int nsem_download_gradients(nsem_ctx* ctx, double* grad_U, double* grad_T);

/* ---- operator-level entry points (unit parity).  Each evaluates ONE of the reference's operators on the state the context
 * holds and leaves that state untouched.  Element-face layout of the FACET outputs: value[(elem * 6 + local_face) * NPF + slot], the
 * face seen from each of its elements, in the face owner's frame (both elements of an interior face report the same number).
 *   nsem_op_cds          cds(cell)                 src/field/field.h:2881-2893
 *   nsem_op_rusanov      rusanov(rho U, rho, lambdaMax) . fN of the mass equation   field.h:2928-2943 (+ div_flux :3093-3114), euler.cpp:186,200-203
 *   nsem_op_gradf_strong gradf<strong>(U), gradf<strong>(theta) + fillBCs   field.h:3328-3362, 2731-2769 (Tensor / Vector AoS as nsem_download_gradients)
 *   nsem_op_divf_weak    divf<weak> residuals of the rho-, U- and theta-equations before src/addTemporal/Solve   field.h:3417-3478, euler.cpp:195-258
 *   nsem_op_apply_bcs    applyExplicitBCs of one field (NSEM_F_RHO, NSEM_F_U, NSEM_F_T)   field.h:2586-2727
 *   nsem_op_halo         ASYNC_COMM of the current state   field.h:2255-2324
 * (the update itself, SolveTexplicit solve.cpp:563-570, is nsem_euler_step: it has no form apart from the residual it divides) */
int nsem_op_cds(nsem_ctx* ctx, const double* cell_field, double* facet_out);
int nsem_op_rusanov(nsem_ctx* ctx, double* facet_flux);
int nsem_op_gradf_strong(nsem_ctx* ctx, double* grad_U, double* grad_T);
int nsem_op_divf_weak(nsem_ctx* ctx, double* r_rho, double* r_U, double* r_T);
int nsem_op_apply_bcs(nsem_ctx* ctx, int field, double* values);
int nsem_op_halo(nsem_ctx* ctx);

/* ---- explicit scalar advection (apps/convection/convection.cpp:113-137: dT/dt + div(T U) = 0 with RUSANOV and a one-stage scheme) ----
 * The transported scalar takes the place of rho in the context (nsem_upload_state(ctx, scalar, U, zeros, zeros); its boundary conditions
 * are set under NSEM_F_RHO) and one step is the mass-equation half of the euler step with lambdaMax = cds(mag(U)) / 2 (:105).
 * problem_init: 0 = NONE (the wind is the uploaded U), 1 = LEVEQUE (the deformational wind of convection.cpp:74-82, re-evaluated on the
 * device before every step at time step * dt with period etime = end_step * dt; needs the node coordinates Mesh::cC). */
int nsem_upload_coords(nsem_ctx* ctx, const double* cC);
/* Mesh::sphere_radius (src/mesh/mesh.cpp:32) of a cubed-sphere mesh (general{is_spherical YES}); read by problem_init 2 / 3 =
 * LAURITZEN_0 / LAURITZEN_1 (convection.cpp:55-72), the deformational winds on the sphere. */
int nsem_set_sphere(nsem_ctx* ctx, double radius);
/* Controls::time_scheme AB2..AB5 for the scalar (ddt + addTemporal, src/field/field.h:3789-3806, 3885-3905): order residuals are kept and
 * combined with the Adams-Bashforth weights, the first steps use the lower orders as the reference does; order 1 (default) is the single
 * forward-Euler stage.  Resets the history.  nsem_euler_step does not take it: no euler example asks for a multi-step scheme. */
int nsem_set_ab_order(nsem_ctx* ctx, int order);
/* Controls::convection_scheme as the convection app's divf reads it (src/field/field.h:3427-3437): 0 RUSANOV (default), 1 CDS, 2 UDS (the
 * upwind side by the sign of flx(U)), 3 BLENDED (blend_factor * CDS + (1 - blend_factor) * UDS), on conforming and on 2:1 faces; the
 * euler step is RUSANOV only. */
int nsem_set_convection_scheme(nsem_ctx* ctx, int scheme, double blend_factor);
int nsem_set_convection(nsem_ctx* ctx, int problem_init, double etime, long first_step);
int nsem_convection_step(nsem_ctx* ctx, int nsteps);
/* Pipelined variants for drivers that stream batches through the device: both only ENQUEUE and return.  The upload copies on a
 * copy-in stream into its own staging buffer and converts the layout on the compute stream, ordered after everything enqueued
 * before it; the download converts on the compute stream into a second staging buffer and copies out on a copy-out stream, so
 * the download of one batch overlaps the upload of the next (PCIe is full duplex).  Host arrays passed to the upload must stay
 * unchanged, and arrays passed to the download are complete, after nsem_sync().  NULL arrays are skipped.  Page-lock the arrays
 * (nsem_pin_host), pageable memory makes the copies synchronous. */
int nsem_upload_state_async(nsem_ctx* ctx, const double* rho, const double* U, const double* T, const double* p);
int nsem_download_state_async(nsem_ctx* ctx, double* rho, double* U, double* T, double* p);
/* Page-lock a host array that will be passed to upload/download repeatedly and outlives the context (the solver's
 * field storage); optional, transfers from pageable memory work too. Unregistered by nsem_destroy. */
int nsem_pin_host(nsem_ctx* ctx, const void* ptr, uint64_t bytes);
/* Hydrostatic reference state and gravity (euler.cpp:105-131); g may be NULL for uniform params.gravity. */
int nsem_upload_ref(nsem_ctx* ctx, const double* rho_ref, const double* p_ref, const double* g);
/* Geopotential gh = dot(g, cC) per node (euler.cpp:113), only needed by the energy diagnostic; optional. */
int nsem_upload_geopotential(nsem_ctx* ctx, const double* gh);

/* ---- the hot path ------------------------------------------------------------------------------------ */
/* nsteps iterations of the time-loop body apps/euler/euler.cpp:179-287 (steps 1-8 and 10 of SURVEY 3.2):
 * rho-, U- and T-equations with divf<weak>/gradf<strong>/rusanov/addTemporal<1>/Solve + BCs + halo. */
int nsem_euler_step(nsem_ctx* ctx, int nsteps);
/* On one partition of at most 32768 elements and nsteps >= 8 the call captures two steps into a CUDA graph, replays it and returns after
 * the replays have finished (launch-bound meshes; NSEM_GRAPH=0 keeps plain launches); otherwise it only enqueues. */
/* applyExplicitBCs(field, true) halos of the set-up phase (euler.cpp:105-146, field.h:2596-2599,2726): fills the
 * inter-partition ghost cells of rho, U, T, p, rho_ref and p_ref from the neighbouring ranks. Collective. */
int nsem_exchange_state_halos(nsem_ctx* ctx);
/* euler.cpp:261-283 + Mesh::calc_courant (field.cpp:440-448): out = {courant max, min, avg, mass, energy,
 * volume} over the real nodes, all-reduced over ranks (reduce_max/min/avg/sum, field.h:953-1006). Collective. */
int nsem_diagnostics(nsem_ctx* ctx, double out[6]);
/* cudaDeviceSynchronize on the context's streams. */
int nsem_sync(nsem_ctx* ctx);
/* Time `nsteps` steps with CUDA events on the launching stream; returns milliseconds in *ms and, when
 * per_kernel_ms != NULL, the accumulated event time of {sweepA, bcA, sweepB, bcB} (4 doubles). */
int nsem_time_steps(nsem_ctx* ctx, int nsteps, double* ms, double* per_kernel_ms);
/* Number of kernels of this library launched since the context was created. */
uint64_t nsem_launch_count(const nsem_ctx* ctx);
/* Which kernel generation and metric path the context runs after nsem_upload_mesh, e.g. "v4 persistent pipelined,
 * metrics on the fly (trilinear map verified)", "v4 persistent pipelined, stored metrics", "v2", "v1". */
const char* nsem_kernel_info(const nsem_ctx* ctx);
/* transport of the halo exchange (replaces MP::isend/irecieve/waitall, src/mp/mp.h:117-131): "peer memory ...", "nccl send/recv" or "none" */
const char* nsem_halo_info(const nsem_ctx* ctx);
/* diagnostics: ms spent waiting for the neighbours' halo flags since the last call: after sweep A, after sweep B, state exchanges */
int nsem_halo_wait_ms(nsem_ctx* ctx, double out[3]);

/* ---- adaptive mesh refinement ----------------------------------------------------------------------- */
/* One regrid as MeshObject::refineMesh (src/mesh/mesh.cpp:2216-2748) reports it and MeshField::refineField
 * (src/field/field.h:1863-2015) consumes it; all arrays are the ones Prepare::refineMesh (field.cpp:884-906) passes on. */
typedef struct {
    uint32_t n_cells_new;         /* nCells: real cells of the regridded mesh */
    const uint32_t* refine_map;   /* families [nchildren, old parent, child_0 ..]: children index cell_map */
    uint32_t n_refine_map;
    const uint32_t* coarse_map;   /* families [nchildren, first, old child_0 ..]: new cell = cell_map[first] */
    uint32_t n_coarse_map;
    const uint32_t* cell_map;     /* intermediate cell index -> new cell, 1u << 31 (Constants::MAX_INT) = removed */
    uint32_t n_cell_map;
    const double* old_cV;         /* gCV of the old mesh [old real cells] */
    const double* old_cC;         /* gCC of the old mesh [*3] */
    const double* new_cV;         /* gCV of the regridded mesh (gCV[id1] in field.h:1983) */
    const double* new_cC;         /* gCC of the regridded mesh */
    const double* old_node_cC;    /* Mesh::cC of the old mesh [old real cells*NP*3]: corner nodes orient the children */
    const double* psi_ref[6];     /* DG::psiRef[d*2+half] (dg.cpp:576-590) */
    const double* psi_cor[6];     /* DG::psiCor[d*2+half] */
} nsem_regrid;
/* MeshField::refineField for rho, U, T, p at once, device to device: `old_ctx` holds the old mesh and the state,
 * `new_ctx` (same device, same order) the regridded mesh (nsem_upload_mesh); its real nodes receive the transferred
 * state -- copied, interpolated onto the children of a split cell with the mass-fix factor, or projected from the
 * children of a merged cell -- in the reference's floating-point operation order (bit-identical results).  The
 * reference writes the transferred fields to files and re-reads them; ghost cells and p are then rebuilt by the solver
 * set-up: nsem_restart_state. */
int nsem_refine_state(nsem_ctx* old_ctx, const nsem_regrid* regrid, nsem_ctx* new_ctx);
/* The restart branch of the solver set-up (apps/euler/euler.cpp:150-162) on the device: p = P0 (rho (T+T0) R / P0)^gamma
 * - p_ref on the real nodes, then the boundary conditions of rho, p, U, T into the ghost cells (applyExplicitBCs) and,
 * with nranks > 1, nsem_exchange_state_halos (collective).  Needs mesh, params, BCs, reference state and state. */
int nsem_restart_state(nsem_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
