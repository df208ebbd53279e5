// nsem_mortar.cuh -- non-conforming (2:1 AMR) faces: the gFMC >= 1 branch of scatter_non_conforming / gather_non_conforming
// (field.h:2019-2248) for the explicit Euler step.
//
// A "mortar group" is one face of a coarse element together with the 2 (2-D) or 4 (3-D) sub-facets that refined
// neighbours share with it.  The reference evaluates every face operator on the nodes of the SUB-facet (the fine side):
//   scatter : the coarse side's trace there is the L2 projection of the coarse face values,
//             fine[ao,bo] = sum_{an,bn} coarse[an,bn] psiRef[d1][an*n1+ao] psiRef[d2][bn*n2+bo]          (field.h:2198-2208)
//   flux    : cds / rusanov on the sub-facet nodes as on any face                                          (field.h:2881-2943)
//   gather  : the fine side takes the flux as it is, the coarse side its projection back,
//             coarse[an,bn] = sum_{ao,bo} flux[ao,bo] psiCor[d1][ao*n1+an] psiCor[d2][bo*n2+bn]           (field.h:2082-2092)
//             and BOTH sides contract with the weighted area vector of the SUB-facet slot                  (div_flux/grad_flux)
// The element sweeps evaluate faces node by node from the two adjacent nodes, which a mortar face does not allow (every
// coarse node sees every fine node of every sub-facet).  So two small kernels, one CTA per group, run ahead of the sweeps
// and leave the finished surface contributions of both sides in a buffer indexed by (element, local face):
//   mortarA_kernel (before sweep A): mass flux + gradient jumps of U and theta            -> 13 values per face node
//   mortarB_kernel (before sweep B): momentum and theta fluxes (viscous part included)   ->  4 values per face node
// and the sweeps add them where their own face loop meets a face flagged FM_MORTAR.  Sub-facets are visited in the
// coarse cell's face order (allFaces), so the coarse sums have a fixed order; no atomics.
#pragma once
#include "nsem_kernels.cuh"

namespace nsem {

static_assert(MORTAR_MAXF == MAXN * MAXN, "mortar block stride");

struct alignas(16) MortarSub {
    uint32_t fine;        // first device node of the fine element
    uint32_t fid_f;       // the sub-facet's local face id in the fine element
    uint32_t flags;       // bit 0: half along the first face axis, bit 1: along the second (field.h:2196-2197); bit 2: the fine cell OWNS the facet
    uint32_t block;       // contribution block of the fine element's face
    double vec[3];        // un-weighted area vector gFN of the sub-facet (outward from the facet's owner)
    double unit[3];
};
static_assert(sizeof(MortarSub) == 64, "MortarSub must be 64 bytes");
struct MortarGroup {
    uint32_t coarse;      // first device node of the coarse element
    uint32_t fid_c;       // local face id in the coarse element
    uint32_t nsub, sub0;  // sub-facets [sub0, sub0 + nsub)
    uint32_t block;       // contribution block of the coarse element's face
    uint32_t pad[3];
};

struct MortarParams {
    uint32_t nGroups;
    int NX, NY, NZ, visc;
    int conv_scheme;          // face value of the mass flux: 0 RUSANOV, 1 CDS, 2 UDS, 3 BLENDED (KParams::conv_scheme; the convection app only)
    double blend;
    double T0, nu, iPr, gammaR;
    double W[3][MAXN];
    const MortarGroup* groups;
    const MortarSub* subs;
    const double* psiRef;     // [3 dirs][2 halves][MAXN*MAXN]: psiRef[d*2+h][in*n+io]
    const double* psiCor;     // same layout: psiCor[d*2+h][io*n+in]
    const double *rho_old, *rho_new, *U_old[3], *T_old, *p, *GU[9], *GT[3];
    double* outA;             // [blocks][13][MORTAR_MAXF]
    double* outB;             // [blocks][4][MORTAR_MAXF]
};

__device__ __forceinline__ int mortar_face_node(int NX, int NY, int NZ, int fid, int a, int b) {
    if (fid < 2) return a * NY * NZ + b * NZ + (fid == 0 ? 0 : NZ - 1);
    if (fid < 4) return a * NY * NZ + (fid == 2 ? 0 : NY - 1) * NZ + b;
    return (fid == 4 ? 0 : NX - 1) * NY * NZ + a * NZ + b;
}

// face axes (d1,d2) and extents (n1,n2) of a local face id; slot = a*n2 + b
__device__ __forceinline__ void mortar_face_axes(int NX, int NY, int NZ, int fid, int& d1, int& d2, int& n1, int& n2) {
    if (fid < 2) { d1 = 0; d2 = 1; n1 = NX; n2 = NY; }
    else if (fid < 4) { d1 = 0; d2 = 2; n1 = NX; n2 = NZ; }
    else { d1 = 1; d2 = 2; n1 = NY; n2 = NZ; }
}

// ---------------------------------------------------------------------------------------------------
// before sweep A: rho-equation flux (euler.cpp:200-203) and the strong-form gradient jumps of U and theta
// (gradf, field.h:3328-3362) on every sub-facet of a group; fine and coarse contributions.
// values per face node: [0] r_rho, [1 + a*3 + b] gU[a][b], [10 + a] gT[a]   (signs included)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MORTAR_MAXF) mortarA_kernel(const __grid_constant__ MortarParams M) {
    constexpr int NC = 9;      // projected coarse fields: rho u (3), rho, |U| + c, u (3), theta
    constexpr int NG = 7;      // gathered sub-facet fields: mass flux vector (3), cds of u (3) and theta
    __shared__ double sC[NC][MORTAR_MAXF];
    __shared__ double sF[NG][MORTAR_MAXF];
    const MortarGroup g = M.groups[blockIdx.x];
    int d1, d2, n1, n2;
    mortar_face_axes(M.NX, M.NY, M.NZ, (int)g.fid_c, d1, d2, n1, n2);
    const int t = threadIdx.x, nf = n1 * n2;
    const bool act = t < nf;
    const int a = act ? t / n2 : 0, b = act ? t % n2 : 0;
    const double w = M.W[d1][a] * M.W[d2][b] / 4;                         // wgl[d1][a] * wgl[d2][b] / 4 (dg.cpp:374,387,400)
    // the coarse cell's own face node (a,b)
    double crho = 0, cu[3] = {0, 0, 0}, cth = 0;
    if (act) {
        const size_t ci = (size_t)g.coarse + mortar_face_node(M.NX, M.NY, M.NZ, (int)g.fid_c, a, b);
        crho = M.rho_old[ci];
        cu[0] = M.U_old[0][ci]; cu[1] = M.U_old[1][ci]; cu[2] = M.U_old[2][ci];
        cth = M.T_old[ci] + M.T0;
        sC[0][t] = crho * cu[0]; sC[1][t] = crho * cu[1]; sC[2][t] = crho * cu[2];
        sC[3][t] = crho;
        sC[4][t] = side_speed(cu, cth, M.gammaR);
        sC[5][t] = cu[0]; sC[6][t] = cu[1]; sC[7][t] = cu[2];
        sC[8][t] = cth;
    }
    __syncthreads();
    double accC[MORTAR_NA];
#pragma unroll
    for (int q = 0; q < MORTAR_NA; q++) accC[q] = 0;

    for (uint32_t si = 0; si < g.nsub; si++) {
        const MortarSub sb = M.subs[g.sub0 + si];
        const int h1 = sb.flags & 1u, h2 = (sb.flags >> 1) & 1u;
        const bool fineOwns = (sb.flags & 4u) != 0;
        const double* R1 = M.psiRef + (size_t)(d1 * 2 + h1) * MAXN * MAXN;
        const double* R2 = M.psiRef + (size_t)(d2 * 2 + h2) * MAXN * MAXN;
        const double* C1 = M.psiCor + (size_t)(d1 * 2 + h1) * MAXN * MAXN;
        const double* C2 = M.psiCor + (size_t)(d2 * 2 + h2) * MAXN * MAXN;
        const double N[3] = {sb.vec[0] * w, sb.vec[1] * w, sb.vec[2] * w};
        const double sgF = fineOwns ? 1.0 : -1.0;           // owner adds the flux, neighbour subtracts it (field.h:3093-3114)
        if (act) {
            // scatter: coarse trace on the sub-facet node (a,b)
            double pc[NC];
#pragma unroll
            for (int q = 0; q < NC; q++) pc[q] = 0;
            for (int an = 0; an < n1; an++)
                for (int bn = 0; bn < n2; bn++) {
                    const double f = R1[an * n1 + a] * R2[bn * n2 + b];
#pragma unroll
                    for (int q = 0; q < NC; q++) pc[q] += sC[q][an * n2 + bn] * f;
                }
            // the fine side's node
            const size_t fi = (size_t)sb.fine + mortar_face_node(M.NX, M.NY, M.NZ, (int)sb.fid_f, a, b);
            const double frho = M.rho_old[fi];
            const double fu[3] = {M.U_old[0][fi], M.U_old[1][fi], M.U_old[2][fi]};
            const double fth = M.T_old[fi] + M.T0;
            const double fS = side_speed(fu, fth, M.gammaR);
            const double fF[3] = {frho * fu[0], frho * fu[1], frho * fu[2]};
            // cds with fI = 1/2 (interior face) and the Rusanov correction - unit(fN) lam (q_N - q_O)
            const double lam = (fS * 0.5 + pc[4] * 0.5) / 2;
            const double dq = fineOwns ? (pc[3] - frho) : (frho - pc[3]);
            double flux[3], cq[4];
#pragma unroll
            for (int d = 0; d < 3; d++) flux[d] = (fF[d] * 0.5 + pc[d] * 0.5) - sb.unit[d] * (lam * dq);
#pragma unroll
            for (int f = 0; f < 3; f++) cq[f] = fu[f] * 0.5 + pc[5 + f] * 0.5;
            cq[3] = fth * 0.5 + pc[8] * 0.5;
            if (M.conv_scheme != 0) {
                // CDS / UDS / BLENDED (divf, field.h:3427-3437): the upwind side by the sign of flx(U) = cds(U).fN on the sub-facet, its two
                // sides being the fine node and the projected coarse trace (uds over scatter_non_conforming, field.h:2904-2917)
                const double F = cq[0] * N[0] + (cq[1] * N[1] + cq[2] * N[2]);
                const bool ownerSide = (F >= 0);
                const bool takeFine = (ownerSide == fineOwns);
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double central = fF[d] * 0.5 + pc[d] * 0.5, upwind = takeFine ? fF[d] : pc[d];
                    flux[d] = (M.conv_scheme == 1) ? central : (M.conv_scheme == 2 ? upwind : M.blend * central + (1.0 - M.blend) * upwind);
                }
            }
            // fine side: the flux as it is
            double* o = M.outA + (size_t)sb.block * MORTAR_NA * MORTAR_MAXF + t;
            o[0] = sgF * (flux[0] * N[0] + flux[1] * N[1] + flux[2] * N[2]);
            if (M.visc) {
                const double dqf[4] = {cq[0] - fu[0], cq[1] - fu[1], cq[2] - fu[2], cq[3] - fth};
#pragma unroll
                for (int aa = 0; aa < 3; aa++) {
                    const double sN = sgF * N[aa];
                    o[(1 + aa * 3 + 0) * MORTAR_MAXF] = sN * dqf[0];
                    o[(1 + aa * 3 + 1) * MORTAR_MAXF] = sN * dqf[1];
                    o[(1 + aa * 3 + 2) * MORTAR_MAXF] = sN * dqf[2];
                    o[(10 + aa) * MORTAR_MAXF] = sN * dqf[3];
                }
            }
#pragma unroll
            for (int d = 0; d < 3; d++) sF[d][t] = flux[d];
#pragma unroll
            for (int f = 0; f < 4; f++) sF[3 + f][t] = cq[f];
        }
        __syncthreads();
        if (act) {
            // gather: projection of the sub-facet fields onto the coarse face node (a,b)
            double G[NG];
#pragma unroll
            for (int q = 0; q < NG; q++) G[q] = 0;
            for (int ao = 0; ao < n1; ao++)
                for (int bo = 0; bo < n2; bo++) {
                    const double f = C1[ao * n1 + a] * C2[bo * n2 + b];
#pragma unroll
                    for (int q = 0; q < NG; q++) G[q] += sF[q][ao * n2 + bo] * f;
                }
            const double sgC = -sgF;
            accC[0] += sgC * (G[0] * N[0] + G[1] * N[1] + G[2] * N[2]);
            if (M.visc) {
                const double dqc[4] = {G[3] - cu[0], G[4] - cu[1], G[5] - cu[2], G[6] - cth};
#pragma unroll
                for (int aa = 0; aa < 3; aa++) {
                    const double sN = sgC * N[aa];
                    accC[1 + aa * 3 + 0] += sN * dqc[0];
                    accC[1 + aa * 3 + 1] += sN * dqc[1];
                    accC[1 + aa * 3 + 2] += sN * dqc[2];
                    accC[10 + aa] += sN * dqc[3];
                }
            }
        }
        __syncthreads();
    }
    if (act) {
        double* o = M.outA + (size_t)g.block * MORTAR_NA * MORTAR_MAXF + t;
#pragma unroll
        for (int q = 0; q < MORTAR_NA; q++) o[q * MORTAR_MAXF] = accC[q];
    }
}

// ---------------------------------------------------------------------------------------------------
// before sweep B: Rusanov fluxes of the U- and theta-equations (euler.cpp:232-233, 249-251).  The projected cell fields
// are exactly the TraceCoef of a node: flux tensor M (9), theta flux V3 (3), q = rho_new U (3), rho_new theta, |U| + c.
// values per face node: [0..2] momentum, [3] theta   (signs included)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mortar_node_coef(const MortarParams& M, size_t i, double K[17]) {
    SideState q;
    q.rho_o = M.rho_old[i]; q.rho_n = M.rho_new[i];
    q.u[0] = M.U_old[0][i]; q.u[1] = M.U_old[1][i]; q.u[2] = M.U_old[2][i];
    q.th = M.T_old[i] + M.T0;
    q.pp = M.p[i];
    if (M.visc) {
#pragma unroll
        for (int c = 0; c < 9; c++) q.gU[c] = M.GU[c][i];
#pragma unroll
        for (int c = 0; c < 3; c++) q.gT[c] = M.GT[c][i];
    }
    TraceCoef T;
    trace_coef(q, side_speed(q.u, q.th, M.gammaR), M.nu, M.iPr, M.visc != 0, T);
#pragma unroll
    for (int c = 0; c < 9; c++) K[c] = T.M[c];
#pragma unroll
    for (int c = 0; c < 3; c++) { K[9 + c] = T.V3[c]; K[12 + c] = T.V4[c]; }
    K[15] = T.q5;
    K[16] = T.S;
}

__global__ void __launch_bounds__(MORTAR_MAXF) mortarB_kernel(const __grid_constant__ MortarParams M) {
    constexpr int NC = 17, NG = 12;
    __shared__ double sC[NC][MORTAR_MAXF];
    __shared__ double sF[NG][MORTAR_MAXF];
    const MortarGroup g = M.groups[blockIdx.x];
    int d1, d2, n1, n2;
    mortar_face_axes(M.NX, M.NY, M.NZ, (int)g.fid_c, d1, d2, n1, n2);
    const int t = threadIdx.x, nf = n1 * n2;
    const bool act = t < nf;
    const int a = act ? t / n2 : 0, b = act ? t % n2 : 0;
    const double w = M.W[d1][a] * M.W[d2][b] / 4;
    if (act) {
        double K[NC];
        mortar_node_coef(M, (size_t)g.coarse + mortar_face_node(M.NX, M.NY, M.NZ, (int)g.fid_c, a, b), K);
#pragma unroll
        for (int q = 0; q < NC; q++) sC[q][t] = K[q];
    }
    __syncthreads();
    double accC[MORTAR_NB] = {0, 0, 0, 0};

    for (uint32_t si = 0; si < g.nsub; si++) {
        const MortarSub sb = M.subs[g.sub0 + si];
        const int h1 = sb.flags & 1u, h2 = (sb.flags >> 1) & 1u;
        const bool fineOwns = (sb.flags & 4u) != 0;
        const double* R1 = M.psiRef + (size_t)(d1 * 2 + h1) * MAXN * MAXN;
        const double* R2 = M.psiRef + (size_t)(d2 * 2 + h2) * MAXN * MAXN;
        const double* C1 = M.psiCor + (size_t)(d1 * 2 + h1) * MAXN * MAXN;
        const double* C2 = M.psiCor + (size_t)(d2 * 2 + h2) * MAXN * MAXN;
        const double N[3] = {sb.vec[0] * w, sb.vec[1] * w, sb.vec[2] * w};
        const double sgF = fineOwns ? 1.0 : -1.0;
        if (act) {
            double pc[NC];
#pragma unroll
            for (int q = 0; q < NC; q++) pc[q] = 0;
            for (int an = 0; an < n1; an++)
                for (int bn = 0; bn < n2; bn++) {
                    const double f = R1[an * n1 + a] * R2[bn * n2 + b];
#pragma unroll
                    for (int q = 0; q < NC; q++) pc[q] += sC[q][an * n2 + bn] * f;
                }
            double K[NC];
            mortar_node_coef(M, (size_t)sb.fine + mortar_face_node(M.NX, M.NY, M.NZ, (int)sb.fid_f, a, b), K);
            const double lam = (K[16] * 0.5 + pc[16] * 0.5) / 2;
            // (q_N - q_O) with the facet's owner/neighbour roles
            double dq[4];
#pragma unroll
            for (int c = 0; c < 3; c++) dq[c] = fineOwns ? (pc[12 + c] - K[12 + c]) : (K[12 + c] - pc[12 + c]);
            dq[3] = fineOwns ? (pc[15] - K[15]) : (K[15] - pc[15]);
            // fF[c][b] = cds(M)[c][b] - unit[c] lam dq[b]   (mul(Vector,Vector) outer product, field.h:2928-2943)
            double fl[NG];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int bb = 0; bb < 3; bb++) fl[c * 3 + bb] = (K[c * 3 + bb] * 0.5 + pc[c * 3 + bb] * 0.5) - sb.unit[c] * (lam * dq[bb]);
#pragma unroll
            for (int bb = 0; bb < 3; bb++) fl[9 + bb] = (K[9 + bb] * 0.5 + pc[9 + bb] * 0.5) - sb.unit[bb] * (lam * dq[3]);
            double* o = M.outB + (size_t)sb.block * MORTAR_NB * MORTAR_MAXF + t;
#pragma unroll
            for (int c = 0; c < 4; c++) o[c * MORTAR_MAXF] = sgF * (fl[c * 3] * N[0] + fl[c * 3 + 1] * N[1] + fl[c * 3 + 2] * N[2]);
#pragma unroll
            for (int q = 0; q < NG; q++) sF[q][t] = fl[q];
        }
        __syncthreads();
        if (act) {
            double G[NG];
#pragma unroll
            for (int q = 0; q < NG; q++) G[q] = 0;
            for (int ao = 0; ao < n1; ao++)
                for (int bo = 0; bo < n2; bo++) {
                    const double f = C1[ao * n1 + a] * C2[bo * n2 + b];
#pragma unroll
                    for (int q = 0; q < NG; q++) G[q] += sF[q][ao * n2 + bo] * f;
                }
            const double sgC = -sgF;
#pragma unroll
            for (int c = 0; c < 4; c++) accC[c] += sgC * (G[c * 3] * N[0] + G[c * 3 + 1] * N[1] + G[c * 3 + 2] * N[2]);
        }
        __syncthreads();
    }
    if (act) {
        double* o = M.outB + (size_t)g.block * MORTAR_NB * MORTAR_MAXF + t;
#pragma unroll
        for (int c = 0; c < 4; c++) o[c * MORTAR_MAXF] = accC[c];
    }
}

}  // namespace nsem
