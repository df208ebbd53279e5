// nsem_kernels.cuh -- sm_100a kernels of the explicit dGSEM Euler step (FP64, HBM-bound).
//
// One reference time step (apps/euler/euler.cpp:179-287) = two grid-wide element sweeps with a boundary
// (ghost) update after each:
//   sweep A : old (rho,U,theta)           -> rho_new, p' = P0 (rho_new theta R/P0)^gamma - p_ref, grad U, grad theta
//   bc    A : ghosts of rho_new, p', grad U, grad theta   (applyExplicitBCs field.h:2586-2727, fillBCs :2731-2769)
//   sweep B : + neighbours' rho_new, p', gradients        -> U_new, theta_new
//   bc    B : ghosts of U_new, T_new
// The split is forced by the sequential coupling rho -> p -> U -> theta of the reference: the
// U/theta fluxes need the NEIGHBOURS' new density and gradients.
//
// Element-centric evaluation: one thread per LGL node, each element recomputes the Rusanov flux on its own
// faces in the owner's frame (div_flux/grad_flux, field.h:3051-3117, are per-cell loops too), so there are no
// atomics and the summation order is fixed.  Node data is structure-of-arrays with element stride NPS
// (NP rounded up to 16 doubles = 128 B) so a warp reads 256 contiguous bytes per array.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nsem {

constexpr int MAXN = 8;   // points per direction (orders 1..7)

__host__ __device__ constexpr int pad_to(int n, int m) { return ((n + m - 1) / m) * m; }

template <int NX, int NY, int NZ>
struct Dims {
    static constexpr int NP = NX * NY * NZ;
    // DG::init_poly (dg.cpp:147-163): the face slot count is the product of the two larger extents
    static constexpr int NPF = (NX <= NY && NX <= NZ) ? NY * NZ : ((NY <= NX && NY <= NZ) ? NX * NZ : NX * NY);
    static constexpr int NPS = pad_to(NP, 16);
    static constexpr int GPS = pad_to(NPF, 4);
    static constexpr int TPE = pad_to(NP, 32);   // threads per element
};

// faceMeta bits
constexpr uint32_t FM_FID_MASK = 7u;      // 0..5 neighbour's local face id, 6 = ghost cell (compact), 7 = absent
constexpr uint32_t FM_GHOST = 6u;
constexpr uint32_t FM_ABSENT = 7u;
constexpr uint32_t FM_OWNER = 8u;         // this element is the face's owner (gFOC)
constexpr uint32_t FM_HALF = 16u;         // fI = 0.5 (interior / inter-rank); else fI = 0 (physical boundary)
constexpr uint32_t FM_MORTAR = 32u;       // non-conforming (2:1) face: the finished surface contributions of this local face are in
                                          // the mortar buffers (nsem_mortar.cuh), block id in faceOther
constexpr uint32_t FM_AFFINE = 64u;       // on face 0 of an ElemRec: the element is a parallelepiped; c[3..5] hold Jin / (w_i w_j w_k / 8)
constexpr int MORTAR_NA = 13, MORTAR_NB = 4;   // values per face node: sweep A {r_rho, gU[9], gT[3]}, sweep B {r_U[3], r_theta}
constexpr int MORTAR_MAXF = 64;                // face-node stride of a mortar block (MAXN * MAXN)

struct FaceRec;
struct ElemRec;
// Halo fused into the sweeps (nsem_halo.cuh): where the persistent sweeps store the face values of their partition-boundary nodes --
// straight into the neighbours' receive windows over NVLink -- and how the last CTA past its boundary elements publishes the epoch.
struct HaloFuse {
    double* win[32];                 // neighbour's region (kind, parity), at the slot where this rank's block starts
    uint64_t stride[32];             // doubles between two fields there
    unsigned long long* flag[32];    // neighbour's arrival flag for (kind, this rank)
    unsigned int* counter;           // local: CTAs that are past their boundary elements
    int npeers;
};
struct KParams {
    // sizes
    uint32_t nB;                 // real elements
    uint32_t nG;                 // ghost (boundary) cells
    uint64_t ghostBase;          // nB * NPS
    // physics
    double P0, T0, R, gamma, nu, iPr, dt, g[3];
    double mrdt, mdt;            // -1/dt, -dt
    int buoyancy, visc, has_gfield;
    int sms;                     // SM count (v4: rotates the heavy warp roles between the CTAs that share an SM)
    // operator-level views (nsem_op_*, unit parity against the reference's own operators): op_mode bit 0 = the sweeps store the RESIDUAL
    // of divf<weak> (field.h:3417-3478) where they would store the updated field (r_rho -> rho_new; r_U, r_theta -> U_new, T_new);
    // op_flux (plain-load sweep A only) = the rusanov mass flux . fN (field.h:2928-2943, 3093-3114) of every element face node,
    // [elem*6 + local face][NPF], in the owner's frame
    int op_mode;
    // face value of the MASS flux in sweep A (divf, field.h:3427-3437): 0 = RUSANOV (the euler app and the default), 1 = CDS, 2 = UDS (upwind by
    // the sign of flx(U) = cds(U).fN), 3 = BLENDED (blend * CDS + (1 - blend) * UDS).  Only the convection app sets it; plain-load sweeps only.
    int conv_scheme;
    double blend;
    double* op_flux;
    // halo fused into the v4 sweeps (null = the separate halo_push_kernel does it): table for this exchange, its epoch, per ghost cell
    // {neighbour index or 0xffffffff, first slot of the cell in the neighbour's window}, and the number of leading schedule positions
    // that hold the partition-boundary elements (the CTAs report once they are past them)
    const HaloFuse* halo;
    unsigned long long haloEpoch;
    const uint32_t* haloGhost;
    uint32_t nHalo;
    uint32_t run;                // v4 sweep A: consecutive schedule positions one CTA handles in a row (RunIter); 0/1 = grid-stride order
    int probe;                   // diagnostics only (NSEM_PROBE): 1 = stream the inputs and skip the arithmetic, 2 = also skip the gathers
    // basis
    double D[3][MAXN * MAXN];    // D[d][s*n+i] = l_i'(x_s)
    double W[3][MAXN];
    double X[3][MAXN];           // LGL nodes in [-1,1] (v4 kernels: metrics on the fly)
    // state in
    const double* rho_old;
    const double* U_old[3];
    const double* T_old;
    // sweep A out / sweep B in
    double* rho_new;
    double* p;                   // p' = p - p_ref
    double* GU[9];               // GU[a*3+b] = d_a U_b  (row-major of the reference Tensor G[ab])
    double* GT[3];
    // sweep B out
    double* U_new[3];
    double* T_new;
    // |U| + c = |U| + sqrt(gamma R theta) of every node (lambdaMax, euler.cpp:186), kept next to the state it belongs to: sweep B (and the
    // ghost update after it) writes S_new with U_new/T_new, the v4 sweeps of the next step read S_old -- no square roots in sweep A
    const double* S_old;
    double* S_new;
    // geometry
    const double* Jinv[9];       // row-major [a*3+d] = d xi_d / d x_a
    const double* cV;
    const double* rho_ref;
    const double* p_ref;
    const double* gfield[3];     // optional per-node gravity
    // element-face tables [nB*6]
    const uint32_t* faceOther;   // first device node of the other cell
    const uint32_t* faceMeta;
    const double* faceVec;       // [nB*6*3] un-weighted area vector gFN (outward from the OWNER)
    const double* faceUnit;      // [nB*6*3] unit(gFN)
    const uint32_t* sched;       // optional processing order
    const struct FaceRec* faceRec;   // [nB*6] the four tables above packed in one 64-byte record (v2 kernels)
    // face traces (v2): per face block 7 x FS doubles {normal fluxes of the 3 momentum eqs and theta, rho_new U.N,
    // rho_new theta, |U| + c} of the side that OWNS the block; block id = elem*6 + local face, ghost cells nB*6 + g
    double* traceA;
    const struct ElemRec* elemRec;   // [nB] v4 kernels: the six face records and the trilinear map of the element
    // non-conforming faces: [blocks][MORTAR_NA | MORTAR_NB][MORTAR_MAXF] surface contributions written by mortarA/B_kernel
    const double* mortarA;
    const double* mortarB;
    // v4 kernels: the arrays each sweep stages, in stage-slot order (so the issuing lanes index them instead of branching)
    const double* srcA[20];
    const double* srcB[32];
};
static_assert(sizeof(KParams) <= 4000, "KParams must fit the kernel parameter space");

struct alignas(16) FaceRec {
    uint32_t other, meta;
    double vec[3];
    double unit[3];
    uint64_t otherBlock;     // face-trace block of the other side (see KParams::traceA)
};
static_assert(sizeof(FaceRec) == 64, "FaceRec must be 64 bytes");
// x(xi) = c000 + c100 xi + c010 eta + c001 zeta + c110 xi eta + c101 xi zeta + c011 eta zeta + c111 xi eta zeta, xi in [-1,1]^3
struct alignas(16) ElemRec {
    FaceRec face[6];
    double c[7][3];          // c100, c010, c001, c110, c101, c011, c111; on an AFFINE element (face[0].meta & FM_AFFINE) the three
                             // mixed terms are zero and c[3..5] instead hold A[a*3+d] = cofactor(J)[a][d] * vol / det J, so that
                             // Jinv*cV at node (i,j,k) is A * (w_i w_j w_k / 8) without the per-node cofactors and division
    double vol;              // element volume: cV[node] = vol * w_i w_j w_k / 8 (dg.cpp:315-318)
};
static_assert(sizeof(ElemRec) == 560, "ElemRec must be 560 bytes");

// ---------------------------------------------------------------------------------------------------
// index helpers (dg.h:43-44 INDEX4; dg.cpp:372-404 face node maps)
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ>
__device__ __forceinline__ int face_slot(int s, int i, int j, int k, int& a, int& b) {
    if (s < 2) { a = i; b = j; return a * NY + b; }
    if (s < 4) { a = i; b = k; return a * NZ + b; }
    a = j; b = k; return a * NZ + b;
}
template <int NX, int NY, int NZ>
__device__ __forceinline__ int face_node(int fid, int a, int b) {
    if (fid < 2) return a * NY * NZ + b * NZ + (fid == 0 ? 0 : NZ - 1);
    if (fid < 4) return a * NY * NZ + (fid == 2 ? 0 : NY - 1) * NZ + b;
    return (fid == 4 ? 0 : NX - 1) * NY * NZ + a * NZ + b;
}
// slot n of a face with local id fid -> local node of the cell that owns that slot numbering
template <int NX, int NY, int NZ>
__device__ __forceinline__ int face_node_from_slot(int fid, int n, bool& valid) {
    int a, b;
    if (fid < 2) { a = n / NY; b = n % NY; valid = n < NX * NY; }
    else if (fid < 4) { a = n / NZ; b = n % NZ; valid = n < NX * NZ; }
    else { a = n / NZ; b = n % NZ; valid = n < NY * NZ; }
    return face_node<NX, NY, NZ>(fid, a, b);
}
template <int NX, int NY, int NZ>
__device__ __forceinline__ double face_weight(const KParams& P, int s, int a, int b) {
    // wgl[.][a] * wgl[.][b] / 4   (dg.cpp:374,387,400)
    if (s < 2) return P.W[0][a] * P.W[1][b] / 4;
    if (s < 4) return P.W[0][a] * P.W[2][b] / 4;
    return P.W[1][a] * P.W[2][b] / 4;
}
__device__ __forceinline__ bool on_face(int s, int i, int j, int k, int NX, int NY, int NZ) {
    switch (s) {
        case 0: return k == 0;
        case 1: return k == NZ - 1;
        case 2: return j == 0;
        case 3: return j == NY - 1;
        case 4: return i == 0;
        default: return i == NX - 1;
    }
}

// p = P0 * pow(rho*theta*R/P0, gamma)   (euler.cpp:211); explicit rounding so every call site agrees bitwise.
// x^gamma is evaluated as exp(gamma * log x): 2-4 ulp instead of pow()'s <= 2 ulp at about half the instructions and without
// the out-of-line call.  A few 1e-16 relative in p is 1e-11 Pa, far below what the parity metric resolves (the reference's
// own -O2 and -O3 builds differ by more, SURVEY finding 6); -DNSEM_EOS_POW restores pow().
__device__ __forceinline__ double eos_pressure(double P0, double R, double gamma, double rho, double theta) {
    double x = __ddiv_rn(__dmul_rn(__dmul_rn(rho, theta), R), P0);
#ifdef NSEM_EOS_POW
    return __dmul_rn(P0, pow(x, gamma));
#else
    return __dmul_rn(P0, exp(__dmul_rn(gamma, log(x))));
#endif
}

// |U| + c of one side (lambdaMax, euler.cpp:186)
__device__ __forceinline__ double side_speed(const double u[3], double th, double gammaR) {
    return sqrt(u[0] * u[0] + (u[1] * u[1] + u[2] * u[2])) + sqrt(gammaR * th);
}
// S = |U| + c of every device node (real and ghost) of the CURRENT state: run once after the state was set from outside a step (upload,
// regrid transfer, restart); inside the time loop sweep B and the ghost update keep it current.
__global__ void __launch_bounds__(256) speed_kernel(uint64_t n, double T0, double gammaR, const double* __restrict__ u0, const double* __restrict__ u1,
                                                    const double* __restrict__ u2, const double* __restrict__ T, double* __restrict__ S) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const double u[3] = {u0[q], u1[q], u2[q]};
    S[q] = side_speed(u, T[q] + T0, gammaR);
}

// LeVeque's deformational wind (apps/convection/convection.cpp:74-82, init_wind_field, re-evaluated at every step): on EVERY node, ghost
// nodes included (they carry their owner's coordinates), u = sin^2(pi x) sin(2 pi y) cos(pi t / period), v = -sin^2(pi y) sin(2 pi x) cos(..)
__global__ void __launch_bounds__(256) wind_leveque_kernel(uint64_t n, double time, double period, const double* __restrict__ x, const double* __restrict__ y,
                                                            double* __restrict__ u0, double* __restrict__ u1, double* __restrict__ u2) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    constexpr double PI = 3.14159265358979323846264;      // Constants::PI
    const double ct = cos(PI * time / period);
    const double sx = sin(PI * x[q]), sy = sin(PI * y[q]);
    u0[q] = pow(sx, 2.0) * sin(2 * PI * y[q]) * ct;
    u1[q] = -pow(sy, 2.0) * sin(2 * PI * x[q]) * ct;
    u2[q] = 0.0;
}

// Adams-Bashforth update of the transported scalar (ddt, field.h:3789-3806, with the history handling of addTemporal, :3885-3905): `r` holds
// the residual of this step (the sweep ran in residual mode).  It becomes PREV(0); prev[j] = PREV(j) after the shift, prev[0] is the buffer
// the oldest entry lived in.  `first`: the field's first step, every entry of the history is this residual (initStore).  `use` = the order
// actually applied, min(scheme order, entries stored).  T_new = (T ap + combination) / ap with ap = -cV / dt; written over r.
struct ABParams {
    uint32_t nB;
    int NP, NPS, order, use, first;
    double dt;
    const double* cV;
    const double* q_old;
    double* r;
    double* prev[5];
};
__global__ void __launch_bounds__(256) ab_update_kernel(const __grid_constant__ ABParams A) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)A.nB * A.NPS || (int)(idx % A.NPS) >= A.NP) return;
    const double r = A.r[idx];
    if (A.first) for (int j = 1; j < A.order; j++) A.prev[j][idx] = r;
    A.prev[0][idx] = r;
    double comb;
    switch (A.use) {
        case 5: comb = (1901 * r - 2774 * A.prev[1][idx] + 2616 * A.prev[2][idx] - 1274 * A.prev[3][idx] + 251 * A.prev[4][idx]) / 720.0; break;
        case 4: comb = (55 * r - 59 * A.prev[1][idx] + 37 * A.prev[2][idx] - 9 * A.prev[3][idx]) / 24.0; break;
        case 3: comb = (23 * r - 16 * A.prev[1][idx] + 5 * A.prev[2][idx]) / 12.0; break;
        case 2: comb = (3 * r - A.prev[1][idx]) / 2.0; break;
        default: comb = r;
    }
    const double ap = (-1.0 / A.dt) * A.cV[idx];
    A.r[idx] = __ddiv_rn(__dadd_rn(__dmul_rn(A.q_old[idx], ap), comb), ap);
}

// Lauritzen's deformational winds on the sphere (convection.cpp:55-72 with cart_to_sphere and wind_field, tensor.h:598-621): zonal and
// meridional components from latitude/longitude of the node, turned into a Cartesian vector.  kind 0 = LAURITZEN_0, 1 = LAURITZEN_1.
__global__ void __launch_bounds__(256) wind_lauritzen_kernel(uint64_t n, int kind, double time, double period, double radius, const double* __restrict__ x,
                                                              const double* __restrict__ y, const double* __restrict__ z, double* __restrict__ u0,
                                                              double* __restrict__ u1, double* __restrict__ u2) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    constexpr double PI = 3.14159265358979323846264;      // Constants::PI
    const double RoT = radius / period;
    const double ct = cos(PI * time / period);
    const double lat = atan2(z[q], sqrt(x[q] * x[q] + y[q] * y[q])), lon = atan2(y[q], x[q]);
    const double lambda = lon - 2.0 * PI * time / period;
    double u, v;
    if (kind == 0) {
        u = 10.0 * RoT * pow(sin(lambda), 2.0) * sin(2.0 * lat) * ct + 2.0 * PI * RoT * cos(lat);
        v = 10.0 * RoT * sin(2.0 * lambda) * cos(lat) * ct;
    } else {
        u = -5.0 * RoT * pow(sin(0.5 * lambda), 2.0) * sin(2.0 * lat) * pow(cos(lat), 2.0) * ct + 2.0 * PI * RoT * cos(lat);
        v = 2.5 * RoT * sin(lambda) * pow(cos(lat), 3.0) * ct;
    }
    u0[q] = -u * sin(lon) - v * sin(lat) * cos(lon);
    u1[q] = +u * cos(lon) - v * sin(lat) * sin(lon);
    u2[q] = +v * cos(lat);
}

// cds(cell field) (field.h:2881-2893): fF = fI * fFO + (1 - fI) * fFN on every element face node, [elem*6 + local face][NPF]; a face is
// evaluated from both of its elements and both write the same value (bitwise: fI is 0 or 1/2).  Operator-level view only (nsem_op_cds).
template <int NX, int NY, int NZ>
__global__ void __launch_bounds__(256) op_cds_kernel(const __grid_constant__ KParams P, const double* __restrict__ f, double* __restrict__ out) {
    using Dm = Dims<NX, NY, NZ>;
    constexpr int NPF = Dm::NPF, NPS = Dm::NPS;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)P.nB * 6 * NPF) return;
    const uint32_t ef = (uint32_t)(gid / NPF);
    const int n = (int)(gid % NPF), s = (int)(ef % 6);
    const uint32_t elem = ef / 6;
    const uint32_t meta = P.faceMeta[ef];
    const uint32_t fid = meta & FM_FID_MASK;
    bool valid;
    const int ln = face_node_from_slot<NX, NY, NZ>(s, n, valid);
    if (!valid || fid == FM_ABSENT || (meta & FM_MORTAR)) { out[gid] = 0.0; return; }
    int a, b;
    if (s < 2) { a = n / NY; b = n % NY; } else { a = n / NZ; b = n % NZ; }
    const size_t oidx = (size_t)P.faceOther[ef] + (fid == FM_GHOST ? n : face_node<NX, NY, NZ>(fid, a, b));
    const bool own = meta & FM_OWNER;
    const double al = (meta & FM_HALF) ? 0.5 : 0.0;
    const double mine = f[(size_t)elem * NPS + ln], other = f[oidx];
    out[gid] = own ? (mine * al + other * (1 - al)) : (other * al + mine * (1 - al));
}

// ---------------------------------------------------------------------------------------------------
// sweep A
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, int EPB, bool VISC>
__global__ void __launch_bounds__(EPB * Dims<NX, NY, NZ>::TPE)
sweepA_kernel(const __grid_constant__ KParams P) {
    using Dm = Dims<NX, NY, NZ>;
    constexpr int NP = Dm::NP, NPS = Dm::NPS, TPE = Dm::TPE;
    __shared__ double sD[3][MAXN * MAXN];
    extern __shared__ double dyn_smem[];
    double (*sR)[3][NP] = reinterpret_cast<double (*)[3][NP]>(dyn_smem);                      // [EPB] contravariant mass flux F.Jin[:,d]
    double (*sQ)[4][NP] = reinterpret_cast<double (*)[4][NP]>(dyn_smem + EPB * 3 * NP);       // [EPB] Ux,Uy,Uz,theta (VISC only)

    const int tid = threadIdx.x;
    for (int q = tid; q < 3 * MAXN * MAXN; q += EPB * TPE) sD[q / (MAXN * MAXN)][q % (MAXN * MAXN)] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];
    const int el = tid / TPE, t = tid % TPE;
    const uint32_t eseq = blockIdx.x * EPB + el;
    const bool active = (eseq < P.nB) && (t < NP);
    const uint32_t elem = (eseq < P.nB) ? (P.sched ? P.sched[eseq] : eseq) : 0u;
    const int i = t / (NY * NZ), j = (t / NZ) % NY, k = t % NZ;
    const size_t idx = (size_t)elem * NPS + t;

    double rho = 0, u[3] = {0, 0, 0}, th = 0, cV = 1, Jin[9];
#pragma unroll
    for (int c = 0; c < 9; c++) Jin[c] = 0;
    if (active) {
        rho = P.rho_old[idx];
        u[0] = P.U_old[0][idx]; u[1] = P.U_old[1][idx]; u[2] = P.U_old[2][idx];
        th = P.T_old[idx] + P.T0;                                   // euler.cpp:181
        cV = P.cV[idx];
#pragma unroll
        for (int c = 0; c < 9; c++) Jin[c] = P.Jinv[c][idx] * cV;   // Jin = Jinv*cV (field.h:3341,3448)
        const double F0 = u[0] * rho, F1 = u[1] * rho, F2 = u[2] * rho;    // fq = U*rho (euler.cpp:200)
#pragma unroll
        for (int d = 0; d < 3; d++) sR[el][d][t] = F0 * Jin[0 * 3 + d] + F1 * Jin[1 * 3 + d] + F2 * Jin[2 * 3 + d];
        if (VISC) { sQ[el][0][t] = u[0]; sQ[el][1][t] = u[1]; sQ[el][2][t] = u[2]; sQ[el][3][t] = th; }
    }
    __syncthreads();
    if (!active) return;

    // ---- volume terms --------------------------------------------------------------------------
    // weak divergence (field.h:3443-3464): r[m] -= sum_q (F_q . Jin_q . dpsi(q->m))
    double r_rho = 0;
    {
        double acc = 0;
#pragma unroll
        for (int ii = 0; ii < NX; ii++) acc += sR[el][0][ii * NY * NZ + j * NZ + k] * sD[0][ii * NX + i];
#pragma unroll
        for (int jj = 0; jj < NY; jj++) acc += sR[el][1][i * NY * NZ + jj * NZ + k] * sD[1][jj * NY + j];
#pragma unroll
        for (int kk = 0; kk < NZ; kk++) acc += sR[el][2][i * NY * NZ + j * NZ + kk] * sD[2][kk * NZ + k];
        r_rho = -acc;
    }
    // strong gradients (field.h:3336-3357): G[a][b] = sum_d Jin[a][d] * d(U_b)/d(xi_d)
    double gU[9], gT[3];
    if (VISC) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
            double d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
            for (int m = 0; m < NX; m++) d0 += sD[0][i * NX + m] * sQ[el][f][m * NY * NZ + j * NZ + k];
#pragma unroll
            for (int m = 0; m < NY; m++) d1 += sD[1][j * NY + m] * sQ[el][f][i * NY * NZ + m * NZ + k];
#pragma unroll
            for (int m = 0; m < NZ; m++) d2 += sD[2][k * NZ + m] * sQ[el][f][i * NY * NZ + j * NZ + m];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const double v = Jin[a * 3 + 0] * d0 + Jin[a * 3 + 1] * d1 + Jin[a * 3 + 2] * d2;
                if (f < 3) gU[a * 3 + f] = v; else gT[a] = v;
            }
        }
    }

    // ---- surface terms (rusanov field.h:2928-2943, div_flux/grad_flux :3051-3117), face-ID order ----
#pragma unroll
    for (int s = 0; s < 6; s++) {
        if (!on_face(s, i, j, k, NX, NY, NZ)) continue;
        const uint32_t meta = P.faceMeta[elem * 6 + s];
        const uint32_t fid = meta & FM_FID_MASK;
        int a, b;
        const int n = face_slot<NX, NY, NZ>(s, i, j, k, a, b);
        if (meta & FM_MORTAR) {
            // non-conforming face: mortarA_kernel left this node's surface terms (scatter/gather_non_conforming, field.h:2019-2248)
            const double* mc = P.mortarA + (size_t)P.faceOther[elem * 6 + s] * (MORTAR_NA * MORTAR_MAXF) + n;
            r_rho += mc[0];
            if (VISC) {
#pragma unroll
                for (int c = 0; c < 9; c++) gU[c] += mc[(1 + c) * MORTAR_MAXF];
#pragma unroll
                for (int c = 0; c < 3; c++) gT[c] += mc[(10 + c) * MORTAR_MAXF];
            }
            continue;
        }
        if (fid == FM_ABSENT) continue;
        const size_t oidx = (size_t)P.faceOther[elem * 6 + s] + (fid == FM_GHOST ? n : face_node<NX, NY, NZ>(fid, a, b));
        const double w = face_weight<NX, NY, NZ>(P, s, a, b);
        const double* fv = P.faceVec + (size_t)(elem * 6 + s) * 3;
        const double* fu = P.faceUnit + (size_t)(elem * 6 + s) * 3;
        const double N0 = fv[0] * w, N1 = fv[1] * w, N2 = fv[2] * w;          // fN[k] (dg.cpp:359)
        const double nN = fu[0] * N0 + fu[1] * N1 + fu[2] * N2;               // unit(fN).fN
        const double rho_x = P.rho_old[oidx];
        const double ux = P.U_old[0][oidx], uy = P.U_old[1][oidx], uz = P.U_old[2][oidx];
        const double th_x = P.T_old[oidx] + P.T0;
        const bool own = meta & FM_OWNER;
        const double al = (meta & FM_HALF) ? 0.5 : 0.0;                       // fI (field.cpp:257-270)
        // owner / neighbour roles
        const double rho_o = own ? rho : rho_x, rho_n = own ? rho_x : rho;
        const double uo0 = own ? u[0] : ux, uo1 = own ? u[1] : uy, uo2 = own ? u[2] : uz;
        const double un0 = own ? ux : u[0], un1 = own ? uy : u[1], un2 = own ? uz : u[2];
        const double th_o = own ? th : th_x, th_n = own ? th_x : th;
        // lambdaMax = (cds(mag(U)) + cds(sqrt(gamma R T)))/2   (euler.cpp:186)
        const double mo = sqrt(uo0 * uo0 + (uo1 * uo1 + uo2 * uo2)), mn = sqrt(un0 * un0 + (un1 * un1 + un2 * un2));
        const double co = sqrt(P.gamma * P.R * th_o), cn = sqrt(P.gamma * P.R * th_n);
        const double lam = ((mo * al + mn * (1 - al)) + (co * al + cn * (1 - al))) / 2;
        // mass flux
        const double fo = rho_o * (uo0 * N0 + uo1 * N1 + uo2 * N2), fn = rho_n * (un0 * N0 + un1 * N1 + un2 * N2);
        double flux = (fo * al + fn * (1 - al)) - lam * (rho_n - rho_o) * nN;
        if (P.conv_scheme != 0) {
            const double central = fo * al + fn * (1 - al);
            const double F = (uo0 * al + un0 * (1 - al)) * N0 + ((uo1 * al + un1 * (1 - al)) * N1 + (uo2 * al + un2 * (1 - al)) * N2);
            const double upwind = (F >= 0) ? fo : fn;
            flux = (P.conv_scheme == 1) ? central : (P.conv_scheme == 2 ? upwind : P.blend * central + (1.0 - P.blend) * upwind);
        }
        r_rho += own ? flux : -flux;
        if (P.op_flux) P.op_flux[((size_t)elem * 6 + s) * Dm::NPF + n] = flux;
        if (VISC) {
            // grad_flux<strong>: r[c1] += fN (x) (cds(P) - P_o) ; r[c2] -= fN (x) (cds(P) - P_n)
            const double sgn = own ? 1.0 : -1.0;
            const double dq[4] = {(uo0 * al + un0 * (1 - al)) - u[0], (uo1 * al + un1 * (1 - al)) - u[1],
                                  (uo2 * al + un2 * (1 - al)) - u[2], (th_o * al + th_n * (1 - al)) - th};
            const double Ns[3] = {sgn * N0, sgn * N1, sgn * N2};
#pragma unroll
            for (int aa = 0; aa < 3; aa++) {
                gU[aa * 3 + 0] += Ns[aa] * dq[0];
                gU[aa * 3 + 1] += Ns[aa] * dq[1];
                gU[aa * 3 + 2] += Ns[aa] * dq[2];
                gT[aa] += Ns[aa] * dq[3];
            }
        }
    }

    // ---- rho update (addTemporal<1> field.h:3875-3920, SolveTexplicit solve.cpp:563-570) ----
    const double ap0 = (-1.0 / P.dt) * cV;
    const double rho_new = (P.op_mode & 1) ? r_rho : (r_rho + rho * ap0) / ap0;
    P.rho_new[idx] = rho_new;
    // p = P0 (rho T R / P0)^gamma ; p -= p_ref   (euler.cpp:211-213)
    P.p[idx] = __dsub_rn(eos_pressure(P.P0, P.R, P.gamma, rho_new, th), P.p_ref[idx]);
    if (VISC) {
#pragma unroll
        for (int c = 0; c < 9; c++) P.GU[c][idx] = gU[c] / cV;      // per-unit-volume (field.h:3359)
#pragma unroll
        for (int c = 0; c < 3; c++) P.GT[c][idx] = gT[c] / cV;
    }
}

// ---------------------------------------------------------------------------------------------------
// sweep B
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, int EPB, bool VISC>
__global__ void __launch_bounds__(EPB * Dims<NX, NY, NZ>::TPE)
sweepB_kernel(const __grid_constant__ KParams P) {
    using Dm = Dims<NX, NY, NZ>;
    constexpr int NP = Dm::NP, NPS = Dm::NPS, TPE = Dm::TPE;
    __shared__ double sD[3][MAXN * MAXN];
    extern __shared__ double dyn_smem[];
    double (*sH)[12][NP] = reinterpret_cast<double (*)[12][NP]>(dyn_smem);   // [EPB] contravariant fluxes: [a*3+d] momentum a, [9+d] theta

    const int tid = threadIdx.x;
    for (int q = tid; q < 3 * MAXN * MAXN; q += EPB * TPE) sD[q / (MAXN * MAXN)][q % (MAXN * MAXN)] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];
    const int el = tid / TPE, t = tid % TPE;
    const uint32_t eseq = blockIdx.x * EPB + el;
    const bool active = (eseq < P.nB) && (t < NP);
    const uint32_t elem = (eseq < P.nB) ? (P.sched ? P.sched[eseq] : eseq) : 0u;
    const int i = t / (NY * NZ), j = (t / NZ) % NY, k = t % NZ;
    const size_t idx = (size_t)elem * NPS + t;

    double rho_o = 0, rho_nw = 1, u[3] = {0, 0, 0}, th = 0, pp = 0, cV = 1, mu = 0;
    double gU[9], gT[3];
    if (active) {
        rho_o = P.rho_old[idx];
        rho_nw = P.rho_new[idx];
        u[0] = P.U_old[0][idx]; u[1] = P.U_old[1][idx]; u[2] = P.U_old[2][idx];
        th = P.T_old[idx] + P.T0;
        pp = P.p[idx];
        cV = P.cV[idx];
        mu = VISC ? rho_o * P.nu : 0.0;                                   // mu = rho*viscosity, OLD rho (euler.cpp:189)
        double Jin[9];
#pragma unroll
        for (int c = 0; c < 9; c++) Jin[c] = P.Jinv[c][idx] * cV;
        if (VISC) {
#pragma unroll
            for (int c = 0; c < 9; c++) gU[c] = P.GU[c][idx];
#pragma unroll
            for (int c = 0; c < 3; c++) gT[c] = P.GT[c][idx];
        }
        // fq = mul(Fc,U) + I p - mu grad U (euler.cpp:232) ; Fc = rho_old U (:184)
        const double Fc[3] = {rho_o * u[0], rho_o * u[1], rho_o * u[2]};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double fq[3];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                fq[b] = Fc[a] * u[b] + (a == b ? pp : 0.0);
                if (VISC) fq[b] -= mu * gU[a * 3 + b];
            }
#pragma unroll
            for (int d = 0; d < 3; d++) sH[el][a * 3 + d][t] = fq[0] * Jin[0 * 3 + d] + fq[1] * Jin[1 * 3 + d] + fq[2] * Jin[2 * 3 + d];
        }
        {
            // fq = Fc*T - (mu/Pr) grad T (euler.cpp:249-250)
            double fq[3];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                fq[b] = Fc[b] * th;
                if (VISC) fq[b] -= (mu * P.iPr) * gT[b];
            }
#pragma unroll
            for (int d = 0; d < 3; d++) sH[el][9 + d][t] = fq[0] * Jin[0 * 3 + d] + fq[1] * Jin[1 * 3 + d] + fq[2] * Jin[2 * 3 + d];
        }
    }
    __syncthreads();
    if (!active) return;

    double r[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        double acc = 0;
#pragma unroll
        for (int ii = 0; ii < NX; ii++) acc += sH[el][a * 3 + 0][ii * NY * NZ + j * NZ + k] * sD[0][ii * NX + i];
#pragma unroll
        for (int jj = 0; jj < NY; jj++) acc += sH[el][a * 3 + 1][i * NY * NZ + jj * NZ + k] * sD[1][jj * NY + j];
#pragma unroll
        for (int kk = 0; kk < NZ; kk++) acc += sH[el][a * 3 + 2][i * NY * NZ + j * NZ + kk] * sD[2][kk * NZ + k];
        r[a] = -acc;
    }

#pragma unroll
    for (int s = 0; s < 6; s++) {
        if (!on_face(s, i, j, k, NX, NY, NZ)) continue;
        const uint32_t meta = P.faceMeta[elem * 6 + s];
        const uint32_t fid = meta & FM_FID_MASK;
        int a, b;
        const int n = face_slot<NX, NY, NZ>(s, i, j, k, a, b);
        if (meta & FM_MORTAR) {
            // non-conforming face: mortarB_kernel left this node's momentum and theta fluxes
            const double* mc = P.mortarB + (size_t)P.faceOther[elem * 6 + s] * (MORTAR_NB * MORTAR_MAXF) + n;
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] += mc[c * MORTAR_MAXF];
            continue;
        }
        if (fid == FM_ABSENT) continue;
        const size_t oidx = (size_t)P.faceOther[elem * 6 + s] + (fid == FM_GHOST ? n : face_node<NX, NY, NZ>(fid, a, b));
        const double w = face_weight<NX, NY, NZ>(P, s, a, b);
        const double* fv = P.faceVec + (size_t)(elem * 6 + s) * 3;
        const double* fu = P.faceUnit + (size_t)(elem * 6 + s) * 3;
        const double N[3] = {fv[0] * w, fv[1] * w, fv[2] * w};
        const double nN = fu[0] * N[0] + fu[1] * N[1] + fu[2] * N[2];
        const bool own = meta & FM_OWNER;
        const double al = (meta & FM_HALF) ? 0.5 : 0.0;
        // the other side
        const double xrho_o = P.rho_old[oidx], xrho_nw = P.rho_new[oidx];
        const double xu[3] = {P.U_old[0][oidx], P.U_old[1][oidx], P.U_old[2][oidx]};
        const double xth = P.T_old[oidx] + P.T0;
        const double xpp = P.p[oidx];
        // per-side normal fluxes  (fq.N)_a = Fc_a (U.N) + p N_a - mu (G.N)_a ;  (fT.N) = theta (Fc.N) - mu/Pr (gT.N)
        double me[4], xe[4];
        {
            const double un = u[0] * N[0] + u[1] * N[1] + u[2] * N[2];
            const double xun = xu[0] * N[0] + xu[1] * N[1] + xu[2] * N[2];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                me[c] = (rho_o * u[c]) * un + pp * N[c];
                xe[c] = (xrho_o * xu[c]) * xun + xpp * N[c];
            }
            me[3] = th * (rho_o * un);
            xe[3] = xth * (xrho_o * xun);
            if (VISC) {
                const double xmu = xrho_o * P.nu;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    me[c] -= mu * (gU[c * 3 + 0] * N[0] + gU[c * 3 + 1] * N[1] + gU[c * 3 + 2] * N[2]);
                    xe[c] -= xmu * (P.GU[c * 3 + 0][oidx] * N[0] + P.GU[c * 3 + 1][oidx] * N[1] + P.GU[c * 3 + 2][oidx] * N[2]);
                }
                me[3] -= (mu * P.iPr) * (gT[0] * N[0] + gT[1] * N[1] + gT[2] * N[2]);
                xe[3] -= (xmu * P.iPr) * (P.GT[0][oidx] * N[0] + P.GT[1][oidx] * N[1] + P.GT[2][oidx] * N[2]);
            }
        }
        // lambdaMax from the OLD state on both sides (euler.cpp:186); symmetric in owner/neighbour when fI = 0.5
        const double mm = sqrt(u[0] * u[0] + (u[1] * u[1] + u[2] * u[2])), xm = sqrt(xu[0] * xu[0] + (xu[1] * xu[1] + xu[2] * xu[2]));
        const double mc = sqrt(P.gamma * P.R * th), xc = sqrt(P.gamma * P.R * xth);
        const double wo = own ? al : 1 - al, wx = own ? 1 - al : al;       // weight of my side / the other side
        const double lam = ((mm * wo + xm * wx) + (mc * wo + xc * wx)) / 2;
        // dissipation: - unit(fN)_a lam ((q_n - q_o).fN) for q = rho_new U ; - lam (q_n - q_o) unit(fN).fN for q = rho_new theta
        const double sg = own ? 1.0 : -1.0;                                 // (q_n - q_o) = sg * (q_other - q_mine)
        double dqN = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) dqN += (xrho_nw * xu[c] - rho_nw * u[c]) * N[c];
        const double dqT = xrho_nw * xth - rho_nw * th;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double flux = (me[c] * wo + xe[c] * wx) - fu[c] * (lam * (sg * dqN));
            r[c] += sg * flux;
        }
        {
            const double flux = (me[3] * wo + xe[3] * wx) - lam * (sg * dqT) * nN;
            r[3] += sg * flux;
        }
    }

    // ---- updates (src field.h:3712-3723, addTemporal<1> :3875-3920 with rho/rho0, SolveTexplicit) ----
    const double ap0 = (-1.0 / P.dt) * cV;
    const double ap = ap0 * rho_nw;
    double g[3] = {P.g[0], P.g[1], P.g[2]};
    if (P.has_gfield) { g[0] = P.gfield[0][idx]; g[1] = P.gfield[1][idx]; g[2] = P.gfield[2][idx]; }
    const double drho = P.buoyancy ? (rho_nw - P.rho_ref[idx]) : 0.0;      // Sc = (rho - rho_ref) g (euler.cpp:224-225)
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double Su = (r[c] - (drho * g[c]) * cV) + (u[c] * rho_o) * ap0;
        P.U_new[c][idx] = (P.op_mode & 1) ? r[c] : Su / ap;
    }
    {
        const double Su = r[3] + (th * rho_o) * ap0;
        P.T_new[idx] = (P.op_mode & 1) ? r[3] : Su / ap - P.T0;             // euler.cpp:286
    }
}

// ---------------------------------------------------------------------------------------------------
// boundary (ghost) updates: applyExplicitBCs (field.h:2586-2727) and fillBCs(r, fIndex) (:2731-2769)
// ---------------------------------------------------------------------------------------------------
struct BCRec {            // BCondition<T> payload (field.h:146-150), in the reference's member order
    double value[3];
    double shape;
    double tvalue[3];
    double tshape;
    double zMin;
};

struct BCParams {
    uint32_t nB, nG;
    uint64_t ghostBase;
    double P0, T0, R, gamma;
    int visc;
    int phase;                       // 0 = after sweep A, 1 = after sweep B
    // ghost-cell tables [nG]
    const uint32_t* bOwner;          // owner element
    const uint8_t* bFid;             // owner's local face id
    const double* bUnit;             // [nG*3] unit(gFN)
    const uint8_t* kind[4];          // per field (enum nsem_field order): bc kind per ghost cell
    const uint32_t* rec[4];          // per field: index into recs
    const uint32_t* peer[4];         // per field: CYCLIC partner ghost cell
    const double* fixedv[4];         // per field: FIXED values [nG*NPF*comps] (may be null)
    const BCRec* recs;
    // fields
    double* rho_new;
    const double* rho_old;           // UNLISTED boundary cells continue from their old value
    double* p;
    const double* T_old;
    const double* p_ref;
    double* GU[9];
    double* GT[3];
    double* U_new[3];
    double* T_new;
    double* S_new;                   // |U| + c of the ghost nodes, written with U_new/T_new in phase 1 (may be null)
};

__device__ __forceinline__ double bc_scalar(int kind, double owner, const BCRec& rc, double peer, double fixedv) {
    switch (kind) {
        case 1: return owner;                                            // NEUMANN: owner + value*|dx|, |dx| == 0
        case 2: return rc.value[0];                                      // DIRICHLET
        case 3: return owner;                                            // SYMMETRY of a scalar (tensor.h:483-485)
        case 4: return peer;                                             // CYCLIC
        case 6: return fixedv;                                           // FIXED
        case 7: return rc.shape * rc.value[0] + (1 - rc.shape) * owner;  // ROBIN
        default: return owner;
    }
}

// sym(Vector p, Vector n) (tensor.h:486-494): tangential projection rescaled to |p|
__device__ __forceinline__ void sym_vector(const double p[3], const double en[3], double out[3]) {
    const double Axx = 1.0 - en[0] * en[0], Ayy = 1.0 - en[1] * en[1], Azz = 1.0 - en[2] * en[2];
    const double Axy = -(en[0] * en[1]), Ayz = -(en[1] * en[2]), Axz = -(en[0] * en[2]);
    const double r0 = Axx * p[0] + Axy * p[1] + Axz * p[2];
    const double r1 = Axy * p[0] + Ayy * p[1] + Ayz * p[2];
    const double r2 = Axz * p[0] + Ayz * p[1] + Azz * p[2];
    const double magR = sqrt(r0 * r0 + (r1 * r1 + r2 * r2));
    // equal(magR, 0) with EqualEpsilon = 1e-7 (tensor.h:462,469-474)
    if (magR <= 1e-7) { out[0] = r0; out[1] = r1; out[2] = r2; return; }
    const double f = sqrt(p[0] * p[0] + (p[1] * p[1] + p[2] * p[2])) / magR;
    out[0] = r0 * f; out[1] = r1 * f; out[2] = r2 * f;
}

template <int NX, int NY, int NZ>
__global__ void __launch_bounds__(256) bc_kernel(const __grid_constant__ BCParams B) {
    using Dm = Dims<NX, NY, NZ>;
    constexpr int NPF = Dm::NPF, NPS = Dm::NPS, GPS = Dm::GPS;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)B.nG * NPF) return;
    const uint32_t g = (uint32_t)(gid / NPF);
    const int n = (int)(gid % NPF);
    bool valid;
    const int ln = face_node_from_slot<NX, NY, NZ>(B.bFid[g], n, valid);
    if (!valid) return;
    const size_t oi = (size_t)B.bOwner[g] * NPS + ln;              // owner node
    const size_t gi = B.ghostBase + (size_t)g * GPS + n;           // ghost node

    auto peer_node = [&](int f) -> size_t {
        const uint32_t pg = B.peer[f][g];
        bool v2;
        const int pl = face_node_from_slot<NX, NY, NZ>(B.bFid[pg], n, v2);
        return (size_t)B.bOwner[pg] * NPS + pl;
    };

    if (B.phase == 0) {
        // rho (after Solve of the rho-equation, solve.cpp:574-581)
        {
            const int kd = B.kind[0][g];
            if (kd != 5) {
                const double peer = (kd == 4) ? B.rho_new[peer_node(0)] : 0.0;
                const double fx = (kd == 6 || kd == 8) ? B.fixedv[0][(size_t)g * NPF + n] : 0.0;
                // UNLISTED: no condition overwrites what Solve left there, the owner's residual over the boundary cell's volume
                B.rho_new[gi] = (kd == 8) ? B.rho_old[gi] + (B.rho_new[oi] - B.rho_old[oi]) * fx
                                          : bc_scalar(kd, B.rho_new[oi], B.recs[B.rec[0][g]], peer, fx);
            }
        }
        // p: the BC acts on the full pressure, then p -= p_ref (euler.cpp:211-213)
        {
            const int kd = B.kind[1][g];
            if (kd != 5) {
                const double po = eos_pressure(B.P0, B.R, B.gamma, B.rho_new[oi], B.T_old[oi] + B.T0);
                double peer = 0.0;
                if (kd == 4) {
                    const size_t pn = peer_node(1);
                    peer = eos_pressure(B.P0, B.R, B.gamma, B.rho_new[pn], B.T_old[pn] + B.T0);
                }
                const double fx = (kd == 6) ? B.fixedv[1][(size_t)g * NPF + n] : 0.0;
                const double pg = bc_scalar(kd, po, B.recs[B.rec[1][g]], peer, fx);
                B.p[gi] = __dsub_rn(pg, B.p_ref[gi]);
            }
        }
        // gradients: ghost = owner copy, then NEUMANN -> value (type-punned), SYMMETRY -> 0 (field.h:2735-2766)
        if (B.visc) {
            {
                const int kd = B.kind[2][g];
                if (kd != 5) {
                    const BCRec& rc = B.recs[B.rec[2][g]];
                    // BCondition<Vector> read as BCondition<Tensor>: 9 consecutive doubles from &value,
                    // in Tensor AoS order XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX (tensor.h:452-454)
                    const double raw[9] = {rc.value[0], rc.value[1], rc.value[2], rc.shape, rc.tvalue[0],
                                           rc.tvalue[1], rc.tvalue[2], rc.tshape, rc.zMin};
                    const int rm[9] = {0, 4, 8, 1, 5, 2, 3, 7, 6};       // AoS component -> row-major a*3+b
#pragma unroll
                    for (int c = 0; c < 9; c++) {
                        double v = B.GU[rm[c]][oi];
                        if (kd == 1) v = raw[c];
                        else if (kd == 3) v = 0.0;
                        B.GU[rm[c]][gi] = v;
                    }
                }
            }
            {
                const int kd = B.kind[3][g];
                if (kd != 5) {
                    const BCRec& rc = B.recs[B.rec[3][g]];
                    const double raw[3] = {rc.value[0], rc.shape, rc.tvalue[0]};   // BCondition<Scalar> read as <Vector>
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        double v = B.GT[c][oi];
                        if (kd == 1) v = raw[c];
                        else if (kd == 3) v = 0.0;
                        B.GT[c][gi] = v;
                    }
                }
            }
        }
    } else {
        // U (Solve of the U-equation)
        {
            const int kd = B.kind[2][g];
            if (kd != 5) {
                const BCRec& rc = B.recs[B.rec[2][g]];
                const double o[3] = {B.U_new[0][oi], B.U_new[1][oi], B.U_new[2][oi]};
                double out[3] = {o[0], o[1], o[2]};
                if (kd == 2) { out[0] = rc.value[0]; out[1] = rc.value[1]; out[2] = rc.value[2]; }
                else if (kd == 3) { const double en[3] = {B.bUnit[g * 3], B.bUnit[g * 3 + 1], B.bUnit[g * 3 + 2]}; sym_vector(o, en, out); }
                else if (kd == 4) { const size_t pn = peer_node(2); out[0] = B.U_new[0][pn]; out[1] = B.U_new[1][pn]; out[2] = B.U_new[2][pn]; }
                else if (kd == 6) { const double* fx = B.fixedv[2] + ((size_t)g * NPF + n) * 3; out[0] = fx[0]; out[1] = fx[1]; out[2] = fx[2]; }
                else if (kd == 7) {
#pragma unroll
                    for (int c = 0; c < 3; c++) out[c] = rc.shape * rc.value[c] + (1 - rc.shape) * o[c];
                }
                B.U_new[0][gi] = out[0]; B.U_new[1][gi] = out[1]; B.U_new[2][gi] = out[2];
            }
        }
        // T: the BC is applied while the field holds theta = T + T0; T -= T0 follows (euler.cpp:258,286)
        {
            const int kd = B.kind[3][g];
            if (kd != 5) {
                const BCRec& rc = B.recs[B.rec[3][g]];
                double v = B.T_new[oi];
                if (kd == 2) v = rc.value[0] - B.T0;
                else if (kd == 4) v = B.T_new[peer_node(3)];
                else if (kd == 6) v = B.fixedv[3][(size_t)g * NPF + n] - B.T0;
                else if (kd == 7) v = (rc.shape * rc.value[0] + (1 - rc.shape) * (v + B.T0)) - B.T0;
                B.T_new[gi] = v;
            }
        }
        // |U| + c of the ghost node, from the values just stored (the v4 sweep A of the next step reads it instead of taking square roots)
        if (B.S_new && B.kind[2][g] != 5 && B.kind[3][g] != 5) {
            const double gu[3] = {B.U_new[0][gi], B.U_new[1][gi], B.U_new[2][gi]};
            B.S_new[gi] = side_speed(gu, B.T_new[gi] + B.T0, B.gamma * B.R);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// face traces: what a neighbour needs from this side of a face node to evaluate the Rusanov fluxes of the U- and
// theta-equations (euler.cpp:232-233,249-251 with rusanov field.h:2928-2943), reduced with the face's weighted
// area vector N (identical on both sides): the side's central normal fluxes, its conserved q contracted with N,
// and its |U| + c for lambdaMax (euler.cpp:186).
// ---------------------------------------------------------------------------------------------------
// layout of a face-trace block: 7 components of NPF doubles each, dense; blocks padded to whole 128-byte lines
__host__ __device__ constexpr int trace_cs(int npf) { return npf; }
__host__ __device__ constexpr int trace_bs(int npf) { return pad_to(7 * npf, 16); }
struct SideState {
    double rho_o, rho_n, u[3], th, pp, gU[9], gT[3];
};
// The five N-dependent trace components are linear in N: out[c] = sum_b K[c][b] N_b.  The coefficients depend on the node
// only, so a node that lies on several faces evaluates them once.  Every producer of a trace (sweep A, the "my side" of
// sweep B, the ghost-cell kernel) goes through these two functions, with the operation order pinned by explicit
// fma/mul, so that the two sides of a face and the two sides of a partition boundary see bitwise identical values.
struct TraceCoef {
    double M[9];     // (rho_o u_c) u_b + p' delta_cb - mu gU[c][b]
    double V3[3];    // theta rho_o u_b - mu/Pr gT[b]
    double V4[3];    // rho_n u_b
    double q5, S;    // rho_n theta, |U| + c
};
__device__ __forceinline__ void trace_coef(const SideState& q, double S, double nu, double iPr, bool visc, TraceCoef& K) {
    const double mu = __dmul_rn(q.rho_o, nu);
    const double mup = __dmul_rn(mu, iPr);
    const double rth = __dmul_rn(q.th, q.rho_o);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double ruc = __dmul_rn(q.rho_o, q.u[c]);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            double m = (c == b) ? __fma_rn(ruc, q.u[b], q.pp) : __dmul_rn(ruc, q.u[b]);
            if (visc) m = __fma_rn(-mu, q.gU[c * 3 + b], m);
            K.M[c * 3 + b] = m;
        }
    }
#pragma unroll
    for (int b = 0; b < 3; b++) {
        double v = __dmul_rn(rth, q.u[b]);
        if (visc) v = __fma_rn(-mup, q.gT[b], v);
        K.V3[b] = v;
        K.V4[b] = __dmul_rn(q.rho_n, q.u[b]);
    }
    K.q5 = __dmul_rn(q.rho_n, q.th);
    K.S = S;
}
__device__ __forceinline__ void trace_apply(const TraceCoef& K, const double N[3], double out[7]) {
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] = __fma_rn(K.M[c * 3 + 2], N[2], __fma_rn(K.M[c * 3 + 1], N[1], __dmul_rn(K.M[c * 3], N[0])));
    out[3] = __fma_rn(K.V3[2], N[2], __fma_rn(K.V3[1], N[1], __dmul_rn(K.V3[0], N[0])));
    out[4] = __fma_rn(K.V4[2], N[2], __fma_rn(K.V4[1], N[1], __dmul_rn(K.V4[0], N[0])));
    out[5] = K.q5;
    out[6] = K.S;
}
__device__ __forceinline__ void side_trace(const SideState& q, const double N[3], double nu, double iPr, double gammaR, bool visc,
                                           double out[7]) {
    TraceCoef K;
    trace_coef(q, side_speed(q.u, q.th, gammaR), nu, iPr, visc, K);
    trace_apply(K, N, out);
}

struct GhostTraceParams {
    uint32_t nB, nG;
    uint64_t ghostBase;
    double T0, nu, iPr, gammaR;
    int visc;
    double W[3][MAXN];
    const uint8_t* bFid;
    const double* bVec;              // [nG*3] area vector gFN of the boundary face
    const double *rho_old, *rho_new, *U_old[3], *T_old, *p, *GU[9], *GT[3];
    double* traceA;
};
template <int NX, int NY, int NZ>
__global__ void __launch_bounds__(256) ghost_trace_kernel(const __grid_constant__ GhostTraceParams G) {
    using Dm = Dims<NX, NY, NZ>;
    constexpr int NPF = Dm::NPF, GPS = Dm::GPS, FS = trace_cs(NPF), TBS = trace_bs(NPF);
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)G.nG * NPF) return;
    const uint32_t g = (uint32_t)(gid / NPF);
    const int n = (int)(gid % NPF);
    const int fid = G.bFid[g];
    int a, b;
    bool valid;
    double w;
    if (fid < 2) { a = n / NY; b = n % NY; valid = n < NX * NY; w = G.W[0][a % MAXN] * G.W[1][b % MAXN] / 4; }
    else if (fid < 4) { a = n / NZ; b = n % NZ; valid = n < NX * NZ; w = G.W[0][a % MAXN] * G.W[2][b % MAXN] / 4; }
    else { a = n / NZ; b = n % NZ; valid = n < NY * NZ; w = G.W[1][a % MAXN] * G.W[2][b % MAXN] / 4; }
    if (!valid) return;
    const size_t gi = G.ghostBase + (size_t)g * GPS + n;
    SideState q;
    q.rho_o = G.rho_old[gi]; q.rho_n = G.rho_new[gi];
    q.u[0] = G.U_old[0][gi]; q.u[1] = G.U_old[1][gi]; q.u[2] = G.U_old[2][gi];
    q.th = G.T_old[gi] + G.T0;
    q.pp = G.p[gi];
    if (G.visc) {
#pragma unroll
        for (int c = 0; c < 9; c++) q.gU[c] = G.GU[c][gi];
#pragma unroll
        for (int c = 0; c < 3; c++) q.gT[c] = G.GT[c][gi];
    }
    const double N[3] = {G.bVec[g * 3] * w, G.bVec[g * 3 + 1] * w, G.bVec[g * 3 + 2] * w};
    double out[7];
    side_trace(q, N, G.nu, G.iPr, G.gammaR, G.visc != 0, out);
    double* dst = G.traceA + ((size_t)G.nB * 6 + g) * TBS + n;
#pragma unroll
    for (int c = 0; c < 7; c++) dst[c * FS] = out[c];
}

// ---------------------------------------------------------------------------------------------------
// diagnostics (euler.cpp:261-283, Mesh::calc_courant field.cpp:440-448, reduce_* field.h:953-1006):
// per block partials of {courant max, courant min, courant sum, mass, energy, volume}; a second launch with one
// block folds the partials in index order (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------------
struct DiagParams {
    uint32_t nB;
    int NP, NPS;
    double P0, T0, R, cp, cv, dt;
    const double *rho, *U[3], *T, *p, *p_ref, *cV, *gh;
    double* partial;     // [nblocks][6]
    int nparts;          // second pass: number of partials to fold
};
__device__ __forceinline__ void diag_block_fold(double v[6], double* sh) {
    // sh: [6][blockDim.x/32]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v[0] = fmax(v[0], __shfl_down_sync(0xffffffffu, v[0], o));
        v[1] = fmin(v[1], __shfl_down_sync(0xffffffffu, v[1], o));
#pragma unroll
        for (int c = 2; c < 6; c++) v[c] += __shfl_down_sync(0xffffffffu, v[c], o);
    }
    if (lane == 0)
        for (int c = 0; c < 6; c++) sh[c * nw + w] = v[c];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < nw; q++) {
            v[0] = fmax(v[0], sh[0 * nw + q]);
            v[1] = fmin(v[1], sh[1 * nw + q]);
            for (int c = 2; c < 6; c++) v[c] += sh[c * nw + q];
        }
    }
}
__global__ void __launch_bounds__(256) diag_kernel(const __grid_constant__ DiagParams D) {
    __shared__ double sh[6 * 8];
    double v[6] = {-1e300, 1e300, 0, 0, 0, 0};
    const uint64_t n = (uint64_t)D.nB * D.NP;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = (q / D.NP) * D.NPS + (q % D.NP);
        const double u0 = D.U[0][i], u1 = D.U[1][i], u2 = D.U[2][i], cv_ = D.cV[i], rho = D.rho[i];
        const double m2 = u0 * u0 + (u1 * u1 + u2 * u2);
        const double co = sqrt(m2) * D.dt / pow(cv_, 1.0 / 3);
        const double th = D.T[i] + D.T0;
        const double e = (D.gh ? D.gh[i] : 0.0) + 0.5 * m2 + pow((D.p[i] + D.p_ref[i]) / D.P0, D.R / D.cp) * th * D.cv;
        v[0] = fmax(v[0], co); v[1] = fmin(v[1], co); v[2] += co;
        v[3] += rho * cv_; v[4] += rho * cv_ * e; v[5] += cv_;
    }
    diag_block_fold(v, sh);
    if (threadIdx.x == 0)
        for (int c = 0; c < 6; c++) D.partial[(size_t)blockIdx.x * 6 + c] = v[c];
}
__global__ void __launch_bounds__(256) diag_fold_kernel(const __grid_constant__ DiagParams D) {
    __shared__ double sh[6 * 8];
    double v[6] = {-1e300, 1e300, 0, 0, 0, 0};
    for (int q = threadIdx.x; q < D.nparts; q += blockDim.x) {
        const double* p = D.partial + (size_t)q * 6;
        v[0] = fmax(v[0], p[0]); v[1] = fmin(v[1], p[1]);
        for (int c = 2; c < 6; c++) v[c] += p[c];
    }
    diag_block_fold(v, sh);
    if (threadIdx.x == 0)
        for (int c = 0; c < 6; c++) D.partial[(size_t)D.nparts * 6 + c] = v[c];
}

// ---------------------------------------------------------------------------------------------------
// halo pack (ASYNC_COMM::send, field.h:2283-2290): sendBuf[f][slot] = P[FO[k]] for every slot of every
// inter-partition face, laid out exactly like the receiver's ghost cells (stride GPS per face) so that the
// receive lands directly in the ghost region of each array (no unpack pass, cf. field.h:2314-2321).
// ---------------------------------------------------------------------------------------------------
struct PackParams {
    int nfields;
    uint64_t nslots;                 // total send slots (faces * GPS over all peers)
    const uint32_t* node;            // [nslots] owner node (device index) or 0xffffffff for padding slots
    const double* src[16];
    double* dst;                     // [nfields][nslots]
};
__global__ void __launch_bounds__(256) halo_pack_kernel(const __grid_constant__ PackParams H) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= H.nslots) return;
    const uint32_t nd = H.node[q];
    for (int f = 0; f < H.nfields; f++) H.dst[(uint64_t)f * H.nslots + q] = (nd != 0xffffffffu) ? H.src[f][nd] : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// layout conversion: reference AoS node arrays <-> device SoA (element stride NPS, compact ghost cells)
// ---------------------------------------------------------------------------------------------------
// src: [nRef*comps] AoS in reference node order; dst[c]: device arrays. ghostRef[g*NPF+n] = reference node of
// ghost slot (or 0xffffffff).
__global__ void scatter_to_device(const double* __restrict__ src, int comps, double* const* dst, const int* compMap,
                                  uint32_t nB, int NP, int NPS, uint32_t nG, int NPF, int GPS,
                                  const uint32_t* __restrict__ ghostRef, uint64_t ghostBase) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nReal = (uint64_t)nB * NP;
    if (gid < nReal) {
        const uint64_t c = gid / NP, t = gid % NP;
        for (int k = 0; k < comps; k++) dst[k][c * NPS + t] = src[gid * comps + compMap[k]];
    } else if (gid < nReal + (uint64_t)nG * NPF) {
        const uint64_t q = gid - nReal;
        const uint32_t ref = ghostRef[q];
        if (ref != 0xffffffffu) {
            const uint64_t g = q / NPF, n = q % NPF;
            for (int k = 0; k < comps; k++) dst[k][ghostBase + g * GPS + n] = src[(uint64_t)ref * comps + compMap[k]];
        }
    }
}
__global__ void gather_from_device(double* __restrict__ dstRef, int comps, const double* const* src, const int* compMap,
                                   uint32_t nB, int NP, int NPS, uint32_t nG, int NPF, int GPS,
                                   const uint32_t* __restrict__ ghostRef, uint64_t ghostBase) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nReal = (uint64_t)nB * NP;
    if (gid < nReal) {
        const uint64_t c = gid / NP, t = gid % NP;
        for (int k = 0; k < comps; k++) dstRef[gid * comps + compMap[k]] = src[k][c * NPS + t];
    } else if (gid < nReal + (uint64_t)nG * NPF) {
        const uint64_t q = gid - nReal;
        const uint32_t ref = ghostRef[q];
        if (ref != 0xffffffffu) {
            const uint64_t g = q / NPF, n = q % NPF;
            for (int k = 0; k < comps; k++) dstRef[(uint64_t)ref * comps + compMap[k]] = src[k][ghostBase + g * GPS + n];
        }
    }
}

}  // namespace nsem
