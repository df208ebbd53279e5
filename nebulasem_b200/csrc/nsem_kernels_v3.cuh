// nsem_kernels_v3.cuh -- sm_100a sweeps, one WARP per element, one THREAD per k-pencil (isotropic 3-D orders, N*N <= 32).
//
// v2 (nsem_kernels_v2.cuh) is latency-bound: 15-20 warps per SM, four CTA-wide barriers per element, one node per
// thread.  Here a warp owns an element and lane (i,j) owns the N nodes (i,j,0..N-1) of its pencil:
//   * no CTA barrier at all (only __syncwarp), every warp runs its element independently of its neighbours;
//   * N independent nodes per thread give instruction- and memory-level parallelism inside the thread;
//   * the zeta-direction part of every tensor-product contraction stays in registers with the derivative matrix
//     as constant-bank operands (k is a compile-time index after unrolling); only xi/eta go through shared memory;
//   * face work is six dense rounds (one per local face id), 25 lanes <-> 25 face nodes, with the owner's side
//     pre-reduced to 9 numbers per face node (normal fluxes, conserved q, |U|+c) so a round moves little through
//     shared memory; results return to the owning lanes and are added in local-face-id order.
// Same mathematics as v1/v2 (same owner-frame Rusanov flux, same update); summation order inside a node differs at
// rounding level.  Deterministic, atomics-free.
#pragma once
#include "nsem_kernels.cuh"

namespace nsem {
namespace v3 {

template <int N>
struct Cfg3 {
    static constexpr int NP = N * N * N, NF = N * N, NPS = pad_to(NP, 16), GPS = pad_to(NF, 4);
    // shared doubles per warp: sweep A: xi/eta mass flux (2 NP) + U,theta (4 NP) + face buffers 6*8*NF + face vectors 6*3(+pad)
    static constexpr int smemA_w = 6 * NP + 6 * 8 * NF + 24;
    // sweep B: xi/eta contravariant fluxes of 4 equations (8 NP) + face buffers 6*9*NF + face vectors
    static constexpr int smemB_w = 8 * NP + 6 * 9 * NF + 24;
};

__device__ __forceinline__ uint32_t rec_meta(double2 q0) { return (uint32_t)((unsigned long long)__double_as_longlong(q0.x) >> 32); }
__device__ __forceinline__ uint32_t rec_other(double2 q0) { return (uint32_t)((unsigned long long)__double_as_longlong(q0.x) & 0xffffffffull); }

template <int N>
__device__ __forceinline__ int fnode(int fid, int a, int b) { return face_node<N, N, N>(fid, a, b); }

// ---------------------------------------------------------------------------------------------------
// sweep B (v3)
// ---------------------------------------------------------------------------------------------------
template <int N, int WPB, bool VISC, int MINB>
__global__ void __launch_bounds__(WPB * 32, MINB) sweepB_v3(const __grid_constant__ KParams P) {
    using C = Cfg3<N>;
    constexpr int NP = C::NP, NF = C::NF, NPS = C::NPS;
    extern __shared__ __align__(16) double smem3[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sw = smem3 + (size_t)warp * C::smemB_w;
    double* sXY = sw;                       // [2][4][NP]
    double* sFace = sw + 8 * NP;            // [6][9][NF]  own-side data in, signed fluxes (first 4 fields) out
    double* sVec = sFace + 6 * 9 * NF;      // [6][3] area vectors gFN (+ pad)
    const uint32_t eseq = blockIdx.x * WPB + warp;
    if (eseq >= P.nB) return;               // whole warp leaves together
    const uint32_t elem = P.sched ? P.sched[eseq] : eseq;
    const bool on = lane < NF;
    const int i = on ? lane / N : 0, j = on ? lane % N : 0;
    const size_t base = (size_t)elem * NPS;

    // face records: lanes 0..5 fetch one each; the area vectors go to shared memory for the owner-side reductions
    double2 q0 = {0, 0}, q1 = {0, 0}, q2 = {0, 0}, q3 = {0, 0};
    if (lane < 6) {
        const double2* rp = reinterpret_cast<const double2*>(P.faceRec + ((size_t)elem * 6 + lane));
        q0 = rp[0]; q1 = rp[1]; q2 = rp[2]; q3 = rp[3];
        sVec[lane * 3 + 0] = q0.y; sVec[lane * 3 + 1] = q1.x; sVec[lane * 3 + 2] = q1.y;
    }
    // per-lane columns of the derivative matrices for the weak-form xi/eta sums: dx[ii] = D0[ii][i], dy[jj] = D1[jj][j]
    double dx[N], dy[N];
#pragma unroll
    for (int m = 0; m < N; m++) { dx[m] = P.D[0][m * N + i]; dy[m] = P.D[1][m * N + j]; }
    const double wi0 = P.W[0][i], wi1 = P.W[1][i], wj1 = P.W[1][j], wj2 = P.W[2][j];
    __syncwarp();

    double r[4][N], rap[N];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int k = 0; k < N; k++) r[a][k] = 0.0;

    // ---- pass 1 over the pencil: contravariant fluxes, zeta sums in registers, owner-side face reductions ----
#pragma unroll
    for (int k = 0; k < N; k++) {
        if (!on) break;
        const int t = i * N * N + j * N + k;
        const size_t idx = base + t;
        const double rho_o = P.rho_old[idx], rho_n = P.rho_new[idx];
        const double u[3] = {P.U_old[0][idx], P.U_old[1][idx], P.U_old[2][idx]};
        const double th = P.T_old[idx] + P.T0;
        const double pp = P.p[idx];
        const double cV = P.cV[idx];
        double Jin[9], gU[9], gT[3];
#pragma unroll
        for (int c = 0; c < 9; c++) Jin[c] = P.Jinv[c][idx] * cV;
        if (VISC) {
#pragma unroll
            for (int c = 0; c < 9; c++) gU[c] = P.GU[c][idx];
#pragma unroll
            for (int c = 0; c < 3; c++) gT[c] = P.GT[c][idx];
        }
        const double mu = VISC ? rho_o * P.nu : 0.0;
        const double Fc[3] = {rho_o * u[0], rho_o * u[1], rho_o * u[2]};
        double fq[4][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                fq[a][b] = Fc[a] * u[b] + (a == b ? pp : 0.0);
                if (VISC) fq[a][b] -= mu * gU[a * 3 + b];
            }
#pragma unroll
        for (int b = 0; b < 3; b++) {
            fq[3][b] = Fc[b] * th;
            if (VISC) fq[3][b] -= (mu * P.iPr) * gT[b];
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const double h0 = fq[a][0] * Jin[0] + fq[a][1] * Jin[3] + fq[a][2] * Jin[6];
            const double h1 = fq[a][0] * Jin[1] + fq[a][1] * Jin[4] + fq[a][2] * Jin[7];
            const double h2 = fq[a][0] * Jin[2] + fq[a][1] * Jin[5] + fq[a][2] * Jin[8];
            sXY[(0 * 4 + a) * NP + t] = h0;
            sXY[(1 * 4 + a) * NP + t] = h1;
#pragma unroll
            for (int m = 0; m < N; m++) r[a][m] -= h2 * P.D[2][k * N + m];     // weak form: r[m] -= H_zeta(q) * l'_m(x_q)
        }
        // temporal + source part of the update: Su = r - Sc cV + (x rho_old) ap0 ; x = Su / (ap0 rho_new)
        const double ap0 = (-1.0 / P.dt) * cV;
        rap[k] = 1.0 / (ap0 * rho_n);
        {
            double g[3] = {P.g[0], P.g[1], P.g[2]};
            if (P.has_gfield) { g[0] = P.gfield[0][idx]; g[1] = P.gfield[1][idx]; g[2] = P.gfield[2][idx]; }
            const double drho = P.buoyancy ? (rho_n - P.rho_ref[idx]) : 0.0;
#pragma unroll
            for (int c = 0; c < 3; c++) r[c][k] += (u[c] * rho_o) * ap0 - (drho * g[c]) * cV;
            r[3][k] += (th * rho_o) * ap0;
        }
        // owner-side reduction for every face this node lies on: {normal fluxes (4), q = rho_new (U, theta) (4), |U| + c}
        const double alam = sqrt(u[0] * u[0] + (u[1] * u[1] + u[2] * u[2])) + sqrt(P.gamma * P.R * th);
        auto own_side = [&](int f, int slot, double w) {
            const double Nx = sVec[f * 3 + 0] * w, Ny = sVec[f * 3 + 1] * w, Nz = sVec[f * 3 + 2] * w;
            const double un = u[0] * Nx + u[1] * Ny + u[2] * Nz;
            double me[4];
#pragma unroll
            for (int c = 0; c < 3; c++) me[c] = Fc[c] * un + pp * (c == 0 ? Nx : (c == 1 ? Ny : Nz));
            me[3] = th * (rho_o * un);
            if (VISC) {
#pragma unroll
                for (int c = 0; c < 3; c++) me[c] -= mu * (gU[c * 3 + 0] * Nx + gU[c * 3 + 1] * Ny + gU[c * 3 + 2] * Nz);
                me[3] -= (mu * P.iPr) * (gT[0] * Nx + gT[1] * Ny + gT[2] * Nz);
            }
            double* o = sFace + (size_t)f * 9 * NF + slot;
#pragma unroll
            for (int c = 0; c < 4; c++) o[c * NF] = me[c];
#pragma unroll
            for (int c = 0; c < 3; c++) o[(4 + c) * NF] = rho_n * u[c];
            o[7 * NF] = rho_n * th;
            o[8 * NF] = alam;
        };
        if (k == 0) own_side(0, i * N + j, wi0 * wj1 / 4);
        if (k == N - 1) own_side(1, i * N + j, wi0 * wj1 / 4);
        if (j == 0) own_side(2, i * N + k, wi0 * P.W[2][k] / 4);
        if (j == N - 1) own_side(3, i * N + k, wi0 * P.W[2][k] / 4);
        if (i == 0) own_side(4, j * N + k, wj1 * P.W[2][k] / 4);
        if (i == N - 1) own_side(5, j * N + k, wj1 * P.W[2][k] / 4);
    }
    __syncwarp();

    // ---- six dense face rounds: lane n <-> face node n = a*N + b of local face f ----
#pragma unroll 1
    for (int f = 0; f < 6; f++) {
        // the record of face f lives in lane f
        const double r0x = __shfl_sync(0xffffffffu, q0.x, f);
        const double vx = __shfl_sync(0xffffffffu, q0.y, f), vy = __shfl_sync(0xffffffffu, q1.x, f), vz = __shfl_sync(0xffffffffu, q1.y, f);
        const double ux = __shfl_sync(0xffffffffu, q2.x, f), uy = __shfl_sync(0xffffffffu, q2.y, f), uz = __shfl_sync(0xffffffffu, q3.x, f);
        if (on) {
            const unsigned long long om = (unsigned long long)__double_as_longlong(r0x);
            const uint32_t meta = (uint32_t)(om >> 32), fid = meta & FM_FID_MASK;
            const int a = i, b = j;                                  // slot n = lane = a*N + b
            const size_t oidx = (size_t)(uint32_t)om + (fid == FM_GHOST ? lane : fnode<N>(fid, a, b));
            const double w = (f < 2) ? wi0 * wj1 / 4 : (f < 4 ? wi0 * wj2 / 4 : wi1 * wj2 / 4);
            const double Nv[3] = {vx * w, vy * w, vz * w};
            const double nN = ux * Nv[0] + uy * Nv[1] + uz * Nv[2];
            const bool own = meta & FM_OWNER;
            const double al = (meta & FM_HALF) ? 0.5 : 0.0;
            const double xro = P.rho_old[oidx], xrn = P.rho_new[oidx];
            const double xu[3] = {P.U_old[0][oidx], P.U_old[1][oidx], P.U_old[2][oidx]};
            const double xth = P.T_old[oidx] + P.T0;
            const double xpp = P.p[oidx];
            const double xun = xu[0] * Nv[0] + xu[1] * Nv[1] + xu[2] * Nv[2];
            double xe[4];
#pragma unroll
            for (int c = 0; c < 3; c++) xe[c] = (xro * xu[c]) * xun + xpp * Nv[c];
            xe[3] = xth * (xro * xun);
            if (VISC) {
                const double xmu = xro * P.nu;
#pragma unroll
                for (int c = 0; c < 3; c++)
                    xe[c] -= xmu * (P.GU[c * 3 + 0][oidx] * Nv[0] + P.GU[c * 3 + 1][oidx] * Nv[1] + P.GU[c * 3 + 2][oidx] * Nv[2]);
                xe[3] -= (xmu * P.iPr) * (P.GT[0][oidx] * Nv[0] + P.GT[1][oidx] * Nv[1] + P.GT[2][oidx] * Nv[2]);
            }
            const double xal = sqrt(xu[0] * xu[0] + (xu[1] * xu[1] + xu[2] * xu[2])) + sqrt(P.gamma * P.R * xth);
            double* o = sFace + (size_t)f * 9 * NF + lane;
            const double wo = own ? al : 1 - al, wx = own ? 1 - al : al;
            const double lam = (o[8 * NF] * wo + xal * wx) / 2;
            const double sg = own ? 1.0 : -1.0;
            double dqN = 0;
#pragma unroll
            for (int c = 0; c < 3; c++) dqN += (xrn * xu[c] - o[(4 + c) * NF]) * Nv[c];
            const double dqT = xrn * xth - o[7 * NF];
            const double fuv[3] = {ux, uy, uz};
            double out[4];
#pragma unroll
            for (int c = 0; c < 3; c++) out[c] = sg * ((o[c * NF] * wo + xe[c] * wx) - fuv[c] * (lam * (sg * dqN)));
            out[3] = sg * ((o[3 * NF] * wo + xe[3] * wx) - lam * (sg * dqT) * nN);
#pragma unroll
            for (int c = 0; c < 4; c++) o[c * NF] = out[c];
        }
    }
    __syncwarp();
    if (!on) return;

    // ---- pass 2: xi/eta sums, face contributions in local-face-id order, update ----
#pragma unroll
    for (int k = 0; k < N; k++) {
        const int t = i * N * N + j * N + k;
        const size_t idx = base + t;
        double res[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            double acc = 0;
#pragma unroll
            for (int m = 0; m < N; m++) acc += sXY[(0 * 4 + a) * NP + m * N * N + j * N + k] * dx[m];
#pragma unroll
            for (int m = 0; m < N; m++) acc += sXY[(1 * 4 + a) * NP + i * N * N + m * N + k] * dy[m];
            res[a] = r[a][k] - acc;
        }
        auto add_face = [&](int f, int slot) {
            const double* o = sFace + (size_t)f * 9 * NF + slot;
#pragma unroll
            for (int a = 0; a < 4; a++) res[a] += o[a * NF];
        };
        if (k == 0) add_face(0, i * N + j);
        if (k == N - 1) add_face(1, i * N + j);
        if (j == 0) add_face(2, i * N + k);
        if (j == N - 1) add_face(3, i * N + k);
        if (i == 0) add_face(4, j * N + k);
        if (i == N - 1) add_face(5, j * N + k);
#pragma unroll
        for (int c = 0; c < 3; c++) P.U_new[c][idx] = res[c] * rap[k];
        P.T_new[idx] = res[3] * rap[k] - P.T0;
    }
}

// ---------------------------------------------------------------------------------------------------
// sweep A (v3)
// ---------------------------------------------------------------------------------------------------
template <int N, int WPB, bool VISC, int MINB>
__global__ void __launch_bounds__(WPB * 32, MINB) sweepA_v3(const __grid_constant__ KParams P) {
    using C = Cfg3<N>;
    constexpr int NP = C::NP, NF = C::NF, NPS = C::NPS;
    extern __shared__ __align__(16) double smem3[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sw = smem3 + (size_t)warp * C::smemA_w;
    double* sXY = sw;                       // [2][NP] xi/eta contravariant mass flux
    double* sQ = sw + 2 * NP;               // [4][NP] Ux,Uy,Uz,theta
    double* sFace = sw + 6 * NP;            // [6][8][NF]: in {rho,Ux,Uy,Uz,theta}; out {signed mass flux, dq[4], signed N[3]}
    const uint32_t eseq = blockIdx.x * WPB + warp;
    if (eseq >= P.nB) return;
    const uint32_t elem = P.sched ? P.sched[eseq] : eseq;
    const bool on = lane < NF;
    const int i = on ? lane / N : 0, j = on ? lane % N : 0;
    const size_t base = (size_t)elem * NPS;

    double2 q0 = {0, 0}, q1 = {0, 0}, q2 = {0, 0}, q3 = {0, 0};
    if (lane < 6) {
        const double2* rp = reinterpret_cast<const double2*>(P.faceRec + ((size_t)elem * 6 + lane));
        q0 = rp[0]; q1 = rp[1]; q2 = rp[2]; q3 = rp[3];
    }
    double dxc[N], dyc[N];                  // weak form columns  D0[m][i], D1[m][j]
#pragma unroll
    for (int m = 0; m < N; m++) { dxc[m] = P.D[0][m * N + i]; dyc[m] = P.D[1][m * N + j]; }
    const double wi0 = P.W[0][i], wi1 = P.W[1][i], wj1 = P.W[1][j], wj2 = P.W[2][j];

    double rr[N], rho[N], cVk[N], pu[4][N];
#pragma unroll
    for (int k = 0; k < N; k++) rr[k] = 0.0;

    // ---- pass 1: mass flux (zeta part in registers), stage U/theta, owner-side face values ----
#pragma unroll
    for (int k = 0; k < N; k++) {
        if (!on) break;
        const int t = i * N * N + j * N + k;
        const size_t idx = base + t;
        rho[k] = P.rho_old[idx];
        pu[0][k] = P.U_old[0][idx]; pu[1][k] = P.U_old[1][idx]; pu[2][k] = P.U_old[2][idx];
        pu[3][k] = P.T_old[idx] + P.T0;
        cVk[k] = P.cV[idx];
        double Jin[9];
#pragma unroll
        for (int c = 0; c < 9; c++) Jin[c] = P.Jinv[c][idx] * cVk[k];
        const double F0 = pu[0][k] * rho[k], F1 = pu[1][k] * rho[k], F2 = pu[2][k] * rho[k];
        sXY[0 * NP + t] = F0 * Jin[0] + F1 * Jin[3] + F2 * Jin[6];
        sXY[1 * NP + t] = F0 * Jin[1] + F1 * Jin[4] + F2 * Jin[7];
        const double h2 = F0 * Jin[2] + F1 * Jin[5] + F2 * Jin[8];
#pragma unroll
        for (int m = 0; m < N; m++) rr[m] -= h2 * P.D[2][k * N + m];
        if (VISC) {
#pragma unroll
            for (int f = 0; f < 4; f++) sQ[f * NP + t] = pu[f][k];
        }
        auto own_side = [&](int f, int slot) {
            double* o = sFace + (size_t)f * 8 * NF + slot;
            o[0] = rho[k];
            o[1 * NF] = pu[0][k]; o[2 * NF] = pu[1][k]; o[3 * NF] = pu[2][k]; o[4 * NF] = pu[3][k];
        };
        if (k == 0) own_side(0, i * N + j);
        if (k == N - 1) own_side(1, i * N + j);
        if (j == 0) own_side(2, i * N + k);
        if (j == N - 1) own_side(3, i * N + k);
        if (i == 0) own_side(4, j * N + k);
        if (i == N - 1) own_side(5, j * N + k);
    }
    __syncwarp();

    // ---- six dense face rounds ----
#pragma unroll 1
    for (int f = 0; f < 6; f++) {
        const double r0x = __shfl_sync(0xffffffffu, q0.x, f);
        const double vx = __shfl_sync(0xffffffffu, q0.y, f), vy = __shfl_sync(0xffffffffu, q1.x, f), vz = __shfl_sync(0xffffffffu, q1.y, f);
        const double ux = __shfl_sync(0xffffffffu, q2.x, f), uy = __shfl_sync(0xffffffffu, q2.y, f), uz = __shfl_sync(0xffffffffu, q3.x, f);
        if (on) {
            const unsigned long long om = (unsigned long long)__double_as_longlong(r0x);
            const uint32_t meta = (uint32_t)(om >> 32), fid = meta & FM_FID_MASK;
            const size_t oidx = (size_t)(uint32_t)om + (fid == FM_GHOST ? lane : fnode<N>(fid, i, j));
            const double w = (f < 2) ? wi0 * wj1 / 4 : (f < 4 ? wi0 * wj2 / 4 : wi1 * wj2 / 4);
            const double N0 = vx * w, N1 = vy * w, N2 = vz * w;
            const double nN = ux * N0 + uy * N1 + uz * N2;
            const bool own = meta & FM_OWNER;
            const double al = (meta & FM_HALF) ? 0.5 : 0.0;
            const double xr = P.rho_old[oidx];
            const double xu0 = P.U_old[0][oidx], xu1 = P.U_old[1][oidx], xu2 = P.U_old[2][oidx];
            const double xth = P.T_old[oidx] + P.T0;
            double* o = sFace + (size_t)f * 8 * NF + lane;
            const double mr = o[0], m0 = o[1 * NF], m1 = o[2 * NF], m2 = o[3 * NF], mth = o[4 * NF];
            const double rho_o = own ? mr : xr, rho_n = own ? xr : mr;
            const double uo0 = own ? m0 : xu0, uo1 = own ? m1 : xu1, uo2 = own ? m2 : xu2;
            const double un0 = own ? xu0 : m0, un1 = own ? xu1 : m1, un2 = own ? xu2 : m2;
            const double th_o = own ? mth : xth, th_n = own ? xth : mth;
            const double mo = sqrt(uo0 * uo0 + (uo1 * uo1 + uo2 * uo2)), mn = sqrt(un0 * un0 + (un1 * un1 + un2 * un2));
            const double co = sqrt(P.gamma * P.R * th_o), cn = sqrt(P.gamma * P.R * th_n);
            const double lam = ((mo * al + mn * (1 - al)) + (co * al + cn * (1 - al))) / 2;
            const double fo = rho_o * (uo0 * N0 + uo1 * N1 + uo2 * N2), fn = rho_n * (un0 * N0 + un1 * N1 + un2 * N2);
            const double flux = (fo * al + fn * (1 - al)) - lam * (rho_n - rho_o) * nN;
            o[0] = own ? flux : -flux;
            if (VISC) {
                const double sgn = own ? 1.0 : -1.0;
                o[1 * NF] = (uo0 * al + un0 * (1 - al)) - m0;
                o[2 * NF] = (uo1 * al + un1 * (1 - al)) - m1;
                o[3 * NF] = (uo2 * al + un2 * (1 - al)) - m2;
                o[4 * NF] = (th_o * al + th_n * (1 - al)) - mth;
                o[5 * NF] = sgn * N0; o[6 * NF] = sgn * N1; o[7 * NF] = sgn * N2;
            }
        }
    }
    __syncwarp();
    if (!on) return;

    // ---- pass 2: xi/eta sums, gradients, faces, rho update, EOS ----
    double dxr[N], dyr[N];                  // strong form rows  D0[i][m], D1[j][m]
#pragma unroll
    for (int m = 0; m < N; m++) { dxr[m] = P.D[0][i * N + m]; dyr[m] = P.D[1][j * N + m]; }
#pragma unroll
    for (int k = 0; k < N; k++) {
        const int t = i * N * N + j * N + k;
        const size_t idx = base + t;
        double acc = 0;
#pragma unroll
        for (int m = 0; m < N; m++) acc += sXY[0 * NP + m * N * N + j * N + k] * dxc[m];
#pragma unroll
        for (int m = 0; m < N; m++) acc += sXY[1 * NP + i * N * N + m * N + k] * dyc[m];
        double r_rho = rr[k] - acc;
        double gU[9], gT[3];
        if (VISC) {
            double Jin[9];
#pragma unroll
            for (int c = 0; c < 9; c++) Jin[c] = P.Jinv[c][idx] * cVk[k];
#pragma unroll
            for (int f = 0; f < 4; f++) {
                double d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
                for (int m = 0; m < N; m++) d0 += dxr[m] * sQ[f * NP + m * N * N + j * N + k];
#pragma unroll
                for (int m = 0; m < N; m++) d1 += dyr[m] * sQ[f * NP + i * N * N + m * N + k];
#pragma unroll
                for (int m = 0; m < N; m++) d2 += P.D[2][k * N + m] * pu[f][m];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const double v = Jin[a * 3 + 0] * d0 + Jin[a * 3 + 1] * d1 + Jin[a * 3 + 2] * d2;
                    if (f < 3) gU[a * 3 + f] = v; else gT[a] = v;
                }
            }
        }
        auto add_face = [&](int f, int slot) {
            const double* o = sFace + (size_t)f * 8 * NF + slot;
            r_rho += o[0];
            if (VISC) {
                const double e0 = o[1 * NF], e1 = o[2 * NF], e2 = o[3 * NF], e3 = o[4 * NF];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const double sn = o[(5 + a) * NF];
                    gU[a * 3 + 0] += sn * e0; gU[a * 3 + 1] += sn * e1; gU[a * 3 + 2] += sn * e2;
                    gT[a] += sn * e3;
                }
            }
        };
        if (k == 0) add_face(0, i * N + j);
        if (k == N - 1) add_face(1, i * N + j);
        if (j == 0) add_face(2, i * N + k);
        if (j == N - 1) add_face(3, i * N + k);
        if (i == 0) add_face(4, j * N + k);
        if (i == N - 1) add_face(5, j * N + k);
        const double ap0 = (-1.0 / P.dt) * cVk[k];
        const double rho_new = (r_rho + rho[k] * ap0) / ap0;
        P.rho_new[idx] = rho_new;
        P.p[idx] = __dsub_rn(eos_pressure(P.P0, P.R, P.gamma, rho_new, pu[3][k]), P.p_ref[idx]);
        if (VISC) {
            const double rcV = 1.0 / cVk[k];
#pragma unroll
            for (int c = 0; c < 9; c++) P.GU[c][idx] = gU[c] * rcV;
#pragma unroll
            for (int c = 0; c < 3; c++) P.GT[c][idx] = gT[c] * rcV;
        }
    }
}

}  // namespace v3
}  // namespace nsem
