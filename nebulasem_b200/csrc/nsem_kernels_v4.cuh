// nsem_kernels_v4.cuh -- persistent, software-pipelined sm_100a sweeps (3-D orders).
//
// Same mathematics as nsem_kernels.cuh (v1) / nsem_kernels_v2.cuh; what changes is how the SM is kept busy:
//   * PERSISTENT CTAs (grid = SMs x resident CTAs) loop over elements; while element e is being computed the
//     bulk-async copies (cp.async.bulk, SASS UBLKCP, completion on an mbarrier) of element e+1 are already in
//     flight into the other half of a two-stage shared-memory ring, so no warp ever waits for HBM latency;
//   * the 560-byte element record (six face records + the element's trilinear map) runs two elements ahead in a
//     three-slot ring, so the neighbour gathers of sweep A (cp.async 8-byte, LDGSTS) and the six neighbour face
//     traces of sweep B can be requested a full element early;
//   * METRICS ON THE FLY (TRI = true): for straight-edged hexahedra (every non-curved mesh: dg.cpp:257-263
//     interpolates the nodes trilinearly) Jinv*cV at a node is a closed form of the element's 7 trilinear
//     coefficient vectors and its volume, so the 10 per-node metric arrays (Jinv, cV) are neither stored nor
//     streamed; nsem_upload_mesh verifies the closed form against the uploaded Jinv/cV at every node and keeps
//     the stored-metric instantiation (TRI = false) for curved meshes;
//   * results leave through shared memory and bulk-async stores (full 128-byte lines, element padding included),
//     the face traces as one dense [6][7][NPF] block per element;
//   * the issuing thread lives in the warp that has no node work.
#pragma once
#include "nsem_kernels_v2.cuh"

namespace nsem {
namespace v4 {

using v2::bulk_g2s;
using v2::mbar_expect_tx;
using v2::mbar_fence_init;
using v2::mbar_init;
using v2::mbar_wait;
using v2::smem_u32;
using v2::Tasks;

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int RECD = (int)(sizeof(ElemRec) / sizeof(double));      // 70 doubles
static_assert(sizeof(ElemRec) % 16 == 0, "ElemRec must be a multiple of 16 bytes (bulk copy)");

template <int NX, int NY, int NZ, bool VISC, bool TRI>
struct Cfg {
    using Dm = Dims<NX, NY, NZ>;
    using Tk = Tasks<NX, NY, NZ>;
    static constexpr int NP = Dm::NP, NPS = Dm::NPS, NPF = Dm::NPF, NFT = Tk::NFT;
    static constexpr int WORK = NP > NFT ? NP : NFT;
    static constexpr int NT = pad_to(WORK, 32);
    static constexpr int ISSUER = (NT - 1 >= WORK) ? NT - 1 : 0;       // an idle lane of the last warp when there is one
    static constexpr int TBS = trace_bs(NPF);                          // doubles per face-trace block
    static constexpr int NFTP = pad_to(NFT, 2);
    // ---- sweep A ----
    // staged arrays: rho, U(3), T, p_ref, [Jinv(9), cV]
    static constexpr int NIN_A = TRI ? 6 : 16;
    static constexpr int A_PREF = 5, A_J = 6, A_CV = 15;
    static constexpr int NOUT_A = VISC ? 14 : 2;                       // rho_new, p, [GU(9), GT(3)]
    static constexpr int FSA = VISC ? 8 : 1;                           // doubles per face task of sweep A
    static constexpr int SF_A = (FSA * NFT > 6 * TBS) ? FSA * NFT : 6 * TBS;
    static constexpr int oInA = 0;
    static constexpr int oRecA = oInA + 2 * NIN_A * NPS;
    static constexpr int oGA = oRecA + 3 * RECD;
    static constexpr int oOutA = oGA + 5 * NFTP;
    static constexpr int oRA = oOutA + NOUT_A * NPS;
    static constexpr int oFA = oRA + pad_to(3 * NP, 2);
    static constexpr int oDA = oFA + pad_to(SF_A, 2);
    static constexpr int oKA = oDA + 3 * MAXN * MAXN;
    static constexpr int oBarA = oKA + 5 * NT;
    static constexpr size_t smemA = sizeof(double) * (size_t)(oBarA + 6);
    // ---- sweep B ----
    // staged arrays: rho_old, rho_new, U(3), T, p, [GU(9), GT(3)], rho_ref, [Jinv(9), cV]
    static constexpr int B_RO = 0, B_RN = 1, B_U = 2, B_T = 5, B_P = 6, B_GU = 7, B_GT = 16;
    static constexpr int B_RR = VISC ? 19 : 7, B_J = B_RR + 1, B_CV = B_J + 9;
    static constexpr int NIN_B = B_RR + 1 + (TRI ? 0 : 10);
    static constexpr int STG_B = NIN_B * NPS + 6 * TBS;               // doubles per stage: arrays, then the six neighbour traces
    static constexpr int oInB = 0;
    static constexpr int oRecB = oInB + 2 * STG_B;
    static constexpr int oHB = oRecB + 3 * RECD;                       // contravariant fluxes [12][NPS] (inviscid runs only; else in place over GU/GT)
    static constexpr int oFB = oHB + (VISC ? 0 : 12 * NPS);
    static constexpr int oDB = oFB + pad_to(4 * NFT, 2);
    static constexpr int oBarB = oDB + 3 * MAXN * MAXN;
    static constexpr size_t smemB = sizeof(double) * (size_t)(oBarB + 6);
    static_assert(4 * NPS <= 6 * TBS, "sweep B stages its four outputs in the consumed trace block");
    static constexpr int minb(size_t smem, int regs) {
        int bs = (int)((227 * 1024) / (smem + 1024));
        int br = 65536 / (NT * regs);
        int b = bs < br ? bs : br;
        return b < 1 ? 1 : (b > 6 ? 6 : b);
    }
};

// Jin = Jinv * cV at reference coordinates (x0,x1,x2) of a straight-edged hexahedron:  x(xi) = sum c_abc xi^a eta^b zeta^c,
// J[a][d] = d x_a / d xi_d,  Jinv[a][d] = d xi_d / d x_a = cofactor(J)[a][d] / det J  (dg.cpp:413-476 evaluates the
// same quantity from the interpolated node coordinates).  c = {c100,c010,c001,c110,c101,c011,c111}[3], then the volume.
__device__ __forceinline__ void tri_metrics(const double* __restrict__ c, double x0, double x1, double x2, double wcv, double Jin[9], double& cV) {
    const double x12 = x1 * x2, x02 = x0 * x2, x01 = x0 * x1;
    double J[9];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const double c100 = c[0 + a], c010 = c[3 + a], c001 = c[6 + a], c110 = c[9 + a], c101 = c[12 + a], c011 = c[15 + a], c111 = c[18 + a];
        J[a * 3 + 0] = c100 + c110 * x1 + c101 * x2 + c111 * x12;
        J[a * 3 + 1] = c010 + c110 * x0 + c011 * x2 + c111 * x02;
        J[a * 3 + 2] = c001 + c101 * x0 + c011 * x1 + c111 * x01;
    }
    double C[9];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, d1 = (d + 1) % 3, d2 = (d + 2) % 3;
            C[a * 3 + d] = J[a1 * 3 + d1] * J[a2 * 3 + d2] - J[a1 * 3 + d2] * J[a2 * 3 + d1];
        }
    const double det = J[0] * C[0] + J[1] * C[1] + J[2] * C[2];
    cV = c[21] * wcv;                                  // element volume * w_i w_j w_k / 8  (dg.cpp:315-318)
    const double sc = cV / det;
#pragma unroll
    for (int q = 0; q < 9; q++) Jin[q] = C[q] * sc;
}

// ---------------------------------------------------------------------------------------------------
// sweep A (v4): one thread per node (NT = NP rounded up to whole warps, every warp has node work); the
// 2(NX NY + NX NZ + NY NZ) face tasks run in one or two passes over the same threads, the second pass on the last warp(s).
// Results leave from registers (coalesced stores); the element's face traces are computed by the node threads from
// registers, collected in shared memory and stored as one dense block by a bulk-async copy.
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, bool VISC, bool TRI>
struct CfgA {
    using Dm = Dims<NX, NY, NZ>;
    using Tk = Tasks<NX, NY, NZ>;
    static constexpr int NP = Dm::NP, NPS = Dm::NPS, NPF = Dm::NPF, NFT = Tk::NFT;
    static constexpr int NT = pad_to(NP, 32);
    static constexpr int EXTRA = NFT > NT ? NFT - NT : 0;               // face tasks of the second pass
    static constexpr int EXW = pad_to(EXTRA, 32);                       // threads (whole warps) of the second pass
    static constexpr bool ok = EXTRA <= NT;                             // at most two face passes
    static constexpr int NGRP = EXW > 0 ? NT / EXW : 1;                 // positions the second pass can rotate through
    static constexpr int TBS = trace_bs(NPF);
    static constexpr int NFTP = pad_to(NFT, 2);
    static constexpr int NIN = TRI ? 6 : 16;                            // rho, U(3), T, p_ref, [Jinv(9), cV]
    static constexpr int A_PREF = 5, A_J = 6, A_CV = 15;
    static constexpr int FS = VISC ? 8 : 1;                             // doubles per face task
    static constexpr int oIn = 0;
    static constexpr int oRec = oIn + 2 * NIN * NPS;
    static constexpr int oG = oRec + 3 * RECD;
    static constexpr int oR = oG + 5 * NFTP;                            // [5][NP]: contravariant mass flux (3), theta, |U| + c
    static constexpr int oF = oR + pad_to(5 * NP, 2);
    static constexpr int oTr = oF + pad_to(FS * NFT, 2);                // [6][TBS]
    static constexpr int oD = oTr + 6 * TBS;
    static constexpr int oBar = oD + 3 * MAXN * MAXN;
    static constexpr size_t smem = sizeof(double) * (size_t)(oBar + 6);
    static constexpr int minb(int regs) {
        int bs = (int)((227 * 1024) / (smem + 1024));
        int br = 65536 / (NT * regs);
        int b = bs < br ? bs : br;
        return b < 1 ? 1 : (b > 8 ? 8 : b);
    }
};

__device__ __forceinline__ int fresh_tid() {
    // re-read %tid so that index arithmetic derived from it is recomputed where it is used instead of living in
    // registers across the whole element loop
    int t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t;
}

template <int NX, int NY, int NZ, bool VISC, bool TRI, int MINB>
__global__ void __launch_bounds__((CfgA<NX, NY, NZ, VISC, TRI>::NT), MINB) sweepA_v4(const __grid_constant__ KParams P) {
    using C = CfgA<NX, NY, NZ, VISC, TRI>;
    using Tk = Tasks<NX, NY, NZ>;
    constexpr int NP = C::NP, NPS = C::NPS, NFT = C::NFT, NT = C::NT, NIN = C::NIN, TBS = C::TBS, NPF = C::NPF;
    constexpr int FS = C::FS, NFTP = C::NFTP, TCS = trace_cs(NPF);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const sm = reinterpret_cast<double*>(smem_raw);
    double* const sIn = sm + C::oIn;         // [2][NIN][NPS]
    double* const sRec = sm + C::oRec;       // [3][RECD]
    double* const sG = sm + C::oG;           // [5][NFTP]  neighbour values of the face tasks (cp.async)
    double* const sR = sm + C::oR;           // [5][NP]
    double* const sF = sm + C::oF;           // [FS][NFT]
    double* const sTr = sm + C::oTr;         // [6][TBS]
    double* const sD = sm + C::oD;           // [3][MAXN*MAXN]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + C::oBar);   // full[2], rec[3]

    const int tid = threadIdx.x;
    const uint32_t stride = gridDim.x;
    uint32_t seq = blockIdx.x;
    if (seq >= P.nB) return;

    auto elem_of = [&](uint32_t q) -> uint32_t { return P.sched ? P.sched[q] : q; };
    auto issue_rec = [&](int slot, uint32_t elem) {
        mbar_expect_tx(&bars[2 + slot], (uint32_t)sizeof(ElemRec));
        bulk_g2s(sRec + slot * RECD, P.elemRec + elem, (uint32_t)sizeof(ElemRec), &bars[2 + slot]);
    };
    auto issue_arrays = [&](int st, uint32_t elem) {
        uint64_t* bar = &bars[st];
        double* dst = sIn + (size_t)st * NIN * NPS;
        const size_t off = (size_t)elem * NPS;
        constexpr uint32_t B = NPS * sizeof(double);
        mbar_expect_tx(bar, (uint32_t)(NIN * B));
        bulk_g2s(dst + 0 * NPS, P.rho_old + off, B, bar);
        bulk_g2s(dst + 1 * NPS, P.U_old[0] + off, B, bar);
        bulk_g2s(dst + 2 * NPS, P.U_old[1] + off, B, bar);
        bulk_g2s(dst + 3 * NPS, P.U_old[2] + off, B, bar);
        bulk_g2s(dst + 4 * NPS, P.T_old + off, B, bar);
        bulk_g2s(dst + C::A_PREF * NPS, P.p_ref + off, B, bar);
        if (!TRI) {
#pragma unroll
            for (int q = 0; q < 9; q++) bulk_g2s(dst + (C::A_J + q) * NPS, P.Jinv[q] + off, B, bar);
            bulk_g2s(dst + C::A_CV * NPS, P.cV + off, B, bar);
        }
    };
    // The warps that run the second face pass and the warp that hosts the issuing thread rotate with the CTA's
    // residency round (CTAs sharing an SM differ by multiples of the SM count), so the heavier warps of the resident CTAs
    // do not all sit on the same SM sub-partition.
    const int round = (int)(blockIdx.x / (unsigned)P.sms);
    const int ex0 = (C::EXW > 0) ? (round % C::NGRP) * C::EXW : 0;
    const int issuer = (((round + 2) * 32) % NT) + 31;
    // second face task of thread t (or -1)
    auto task2_of = [&](int t) -> int { return (C::EXTRA > 0 && t >= ex0 && t - ex0 < C::EXTRA) ? NT + (t - ex0) : -1; };
    // neighbour values of face task `task` for the element whose record sits in `slot`
    auto gather_task = [&](int task, int slot) {
        int fs, fa, fb;
        Tk::decode(task, fs, fa, fb);
        const FaceRec* fr = reinterpret_cast<const FaceRec*>(sRec + slot * RECD) + fs;
        const uint32_t other = fr->other, fid = fr->meta & FM_FID_MASK;
        const int fslot = (fs < 2) ? fa * NY + fb : fa * NZ + fb;
        const size_t oidx = (size_t)other + (fid == FM_GHOST ? fslot : face_node<NX, NY, NZ>(fid, fa, fb));
        cp_async8(sG + 0 * NFTP + task, P.rho_old + oidx);
        cp_async8(sG + 1 * NFTP + task, P.U_old[0] + oidx);
        cp_async8(sG + 2 * NFTP + task, P.U_old[1] + oidx);
        cp_async8(sG + 3 * NFTP + task, P.U_old[2] + oidx);
        cp_async8(sG + 4 * NFTP + task, P.T_old + oidx);
    };
    auto issue_gathers = [&](int slot) {
        const int t = fresh_tid();
        if (t < NFT) gather_task(t, slot);
        const int t2 = task2_of(t);
        if (t2 >= 0) gather_task(t2, slot);
        cp_async_commit();
    };

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 5; q++) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == issuer) {
        issue_rec(0, elem_of(seq));
        if (seq + stride < P.nB) issue_rec(1, elem_of(seq + stride));
        issue_arrays(0, elem_of(seq));
    }
    for (int q = tid; q < 3 * MAXN * MAXN; q += NT) sD[q] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];
    mbar_wait(&bars[2], 0);
    issue_gathers(0);

    int st = 0, rs = 0;                    // stage / record slot of the current element
    uint32_t phase = 1u << 2;              // bit q = parity the next wait on bars[q] uses; rec[0] has completed phase 0
    for (;;) {
        const uint32_t elem = elem_of(seq);
        const uint32_t nxt = seq + stride, nxt2 = nxt + stride;
        const bool hasNext = nxt < P.nB;
        const int rs1 = (rs == 2) ? 0 : rs + 1, rs2 = (rs1 == 2) ? 0 : rs1 + 1;
        if (tid == issuer) {
            if (hasNext) issue_arrays(st ^ 1, elem_of(nxt));
            if (nxt2 < P.nB) issue_rec(rs2, elem_of(nxt2));
        }
        mbar_wait(&bars[st], (phase >> st) & 1u);
        phase ^= 1u << st;
        const double* const in = sIn + (size_t)st * NIN * NPS;
        const double* const rec = sRec + rs * RECD;

        // ---- node: contravariant mass flux, theta ----
        const int nt = fresh_tid();
        const bool nodeT = nt < NP;
        const int i = nt / (NY * NZ), j = (nt / NZ) % NY, k = nt % NZ;
        double rho = 0, th = 0, cV = 1, Jin[9];
        if (nodeT) {
            rho = in[0 * NPS + nt];
            const double u0 = in[1 * NPS + nt], u1 = in[2 * NPS + nt], u2 = in[3 * NPS + nt];
            th = in[4 * NPS + nt] + P.T0;
            if (TRI) {
                const double wcv = ((P.W[0][i] * P.W[1][j]) * P.W[2][k]) / 8;
                tri_metrics(rec + 48, P.X[0][i], P.X[1][j], P.X[2][k], wcv, Jin, cV);
            } else {
                cV = in[C::A_CV * NPS + nt];
#pragma unroll
                for (int c = 0; c < 9; c++) Jin[c] = in[(C::A_J + c) * NPS + nt] * cV;
            }
            const double F0 = u0 * rho, F1 = u1 * rho, F2 = u2 * rho;
#pragma unroll
            for (int d = 0; d < 3; d++) sR[d * NP + nt] = F0 * Jin[d] + F1 * Jin[3 + d] + F2 * Jin[6 + d];
            sR[3 * NP + nt] = th;
            const double uu[3] = {u0, u1, u2};
            sR[4 * NP + nt] = side_speed(uu, th, P.gamma * P.R);       // used by this node's face tasks and by its traces
        }
        __syncthreads();                                                                   // (1)

        double r_rho = 0, gU[9], gT[3];
        if (nodeT) {
            double acc = 0;
#pragma unroll
            for (int ii = 0; ii < NX; ii++) acc += sR[0 * NP + ii * NY * NZ + j * NZ + k] * sD[0 * MAXN * MAXN + ii * NX + i];
#pragma unroll
            for (int jj = 0; jj < NY; jj++) acc += sR[1 * NP + i * NY * NZ + jj * NZ + k] * sD[1 * MAXN * MAXN + jj * NY + j];
#pragma unroll
            for (int kk = 0; kk < NZ; kk++) acc += sR[2 * NP + i * NY * NZ + j * NZ + kk] * sD[2 * MAXN * MAXN + kk * NZ + k];
            r_rho = -acc;
            if (VISC) {
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    const double* q = (f < 3) ? in + (1 + f) * NPS : sR + 3 * NP;
                    double d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
                    for (int m = 0; m < NX; m++) d0 += sD[0 * MAXN * MAXN + i * NX + m] * q[m * NY * NZ + j * NZ + k];
#pragma unroll
                    for (int m = 0; m < NY; m++) d1 += sD[1 * MAXN * MAXN + j * NY + m] * q[i * NY * NZ + m * NZ + k];
#pragma unroll
                    for (int m = 0; m < NZ; m++) d2 += sD[2 * MAXN * MAXN + k * NZ + m] * q[i * NY * NZ + j * NZ + m];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        const double v = Jin[a * 3 + 0] * d0 + Jin[a * 3 + 1] * d1 + Jin[a * 3 + 2] * d2;
                        if (f < 3) gU[a * 3 + f] = v; else gT[a] = v;
                    }
                }
            }
        }

        // ---- face tasks (the neighbour values were requested one element ago) ----
        cp_async_wait_all();
        auto face_task = [&](int task) {
            int fs, fa, fb;
            Tk::decode(task, fs, fa, fb);
            const int fln = face_node<NX, NY, NZ>(fs, fa, fb);
            const double fw = face_weight<NX, NY, NZ>(P, fs, fa, fb);
            const FaceRec* fr = reinterpret_cast<const FaceRec*>(rec) + fs;
            const uint32_t meta = fr->meta;
            const double xr = sG[0 * NFTP + task], xu0 = sG[1 * NFTP + task], xu1 = sG[2 * NFTP + task], xu2 = sG[3 * NFTP + task];
            const double xth = sG[4 * NFTP + task] + P.T0;
            // written for "my side" / "other side": with fI in {0, 1/2} this is bitwise cds() = fI*owner + (1-fI)*neighbour
            const bool own = meta & FM_OWNER;
            const double al = (meta & FM_HALF) ? 0.5 : 0.0;
            const double wo = own ? al : 1 - al, wx = own ? 1 - al : al;       // weight of my side / the other side
            const double sg = own ? 1.0 : -1.0;                                // (q_n - q_o) = sg * (q_other - q_mine)
            const double N0 = fr->vec[0] * fw, N1 = fr->vec[1] * fw, N2 = fr->vec[2] * fw;      // fN[k] = gFN * w_a w_b / 4
            const double nN = fr->unit[0] * N0 + fr->unit[1] * N1 + fr->unit[2] * N2;           // unit(fN).fN
            const double mr = in[0 * NPS + fln], m0 = in[1 * NPS + fln], m1 = in[2 * NPS + fln], m2 = in[3 * NPS + fln];
            const double mth = in[4 * NPS + fln] + P.T0;
            // lambdaMax = cds(|U| + c) / 2: my side's |U| + c comes from the node pass, the other side's is evaluated here
            const double xu[3] = {xu0, xu1, xu2};
            const double lam = (sR[4 * NP + fln] * wo + side_speed(xu, xth, P.gamma * P.R) * wx) / 2;
            const double fm = mr * (m0 * N0 + m1 * N1 + m2 * N2), fx = xr * (xu0 * N0 + xu1 * N1 + xu2 * N2);
            const double flux = (fm * wo + fx * wx) - lam * (sg * (xr - mr)) * nN;
            double* out = &sF[task];
            out[0] = sg * flux;
            if (VISC) {
                out[1 * NFT] = (m0 * wo + xu0 * wx) - m0;
                out[2 * NFT] = (m1 * wo + xu1 * wx) - m1;
                out[3 * NFT] = (m2 * wo + xu2 * wx) - m2;
                out[4 * NFT] = (mth * wo + xth * wx) - mth;
                out[5 * NFT] = sg * N0; out[6 * NFT] = sg * N1; out[7 * NFT] = sg * N2;
            }
        };
        {
            const int t = fresh_tid();
            if (t < NFT) face_task(t);
            const int t2 = task2_of(t);
            if (t2 >= 0) face_task(t2);
        }
        if (hasNext) {
            // the next element's record was requested a whole element ago; its gathers land while this element finishes
            mbar_wait(&bars[2 + rs1], (phase >> (2 + rs1)) & 1u);
            phase ^= 1u << (2 + rs1);
            issue_gathers(rs1);
        }
        if (tid == issuer) bulk_wait_read0();     // the previous element's trace block has left sTr
        __syncthreads();                                                                   // (2)

        if (nodeT) {
            // faces this node lies on, in local-face-id order (field.h:3093-3114): k-faces 0/1, j-faces 2/3, i-faces 4/5
#pragma unroll
            for (int ax = 0; ax < 3; ax++) {
                const int cx = (ax == 0) ? k : (ax == 1 ? j : i);
                const int nx = (ax == 0) ? NZ : (ax == 1 ? NY : NX);
                if (cx != 0 && cx != nx - 1) continue;
                const int s = 2 * ax + (cx != 0 ? 1 : 0);
                const int ti = (ax == 0) ? Tk::index(s, i, j) : (ax == 1 ? Tk::index(s, i, k) : Tk::index(s, j, k));
                const double* fi = &sF[ti];
                r_rho += fi[0];
                if (VISC) {
                    const double q0_ = fi[1 * NFT], q1_ = fi[2 * NFT], q2_ = fi[3 * NFT], q3_ = fi[4 * NFT];
#pragma unroll
                    for (int aa = 0; aa < 3; aa++) {
                        const double sn = fi[(5 + aa) * NFT];
                        gU[aa * 3 + 0] += sn * q0_;
                        gU[aa * 3 + 1] += sn * q1_;
                        gU[aa * 3 + 2] += sn * q2_;
                        gT[aa] += sn * q3_;
                    }
                }
            }
            const size_t idx = (size_t)elem * NPS + nt;
            // (Su + rho*ap0) / ap0 with ap0 = (-1/dt) cV (addTemporal<1>, SolveTexplicit) as ONE reciprocal of cV
            const double rcV = 1.0 / cV;
            const double ap0 = P.mrdt * cV;
            const double rho_new = (r_rho + rho * ap0) * (rcV * P.mdt);
            const double ppn = __dsub_rn(eos_pressure(P.P0, P.R, P.gamma, rho_new, th), in[C::A_PREF * NPS + nt]);
            P.rho_new[idx] = rho_new;
            P.p[idx] = ppn;
            if (VISC) {
                // r / cV (field.h:3359) as 12 products with the reciprocal
#pragma unroll
                for (int c = 0; c < 9; c++) { gU[c] *= rcV; P.GU[c][idx] = gU[c]; }
#pragma unroll
                for (int c = 0; c < 3; c++) { gT[c] *= rcV; P.GT[c][idx] = gT[c]; }
            }
            // ---- this side's face traces for sweep B of the neighbours (and of the peers behind a partition boundary),
            //      straight from the registers of the node that lies on the face ----
            const bool onK = (k == 0 || k == NZ - 1), onJ = (j == 0 || j == NY - 1), onI = (i == 0 || i == NX - 1);
            if (onK || onJ || onI) {
                SideState q;
                q.rho_o = rho; q.rho_n = rho_new; q.th = th; q.pp = ppn;
                q.u[0] = in[1 * NPS + nt]; q.u[1] = in[2 * NPS + nt]; q.u[2] = in[3 * NPS + nt];
                if (VISC) {
#pragma unroll
                    for (int c = 0; c < 9; c++) q.gU[c] = gU[c];
#pragma unroll
                    for (int c = 0; c < 3; c++) q.gT[c] = gT[c];
                }
                TraceCoef K;
                trace_coef(q, sR[4 * NP + nt], P.nu, P.iPr, VISC, K);
                const double wi = P.W[0][i], wj = P.W[1][j], wk = P.W[2][k];
#pragma unroll
                for (int ax = 0; ax < 3; ax++) {
                    const bool on = (ax == 0) ? onK : (ax == 1 ? onJ : onI);
                    if (!on) continue;
                    const int cx = (ax == 0) ? k : (ax == 1 ? j : i);
                    const int s = 2 * ax + (cx != 0 ? 1 : 0);
                    const int slot = (ax == 0) ? i * NY + j : (ax == 1 ? i * NZ + k : j * NZ + k);
                    const double fw = (ax == 0) ? wi * wj / 4 : (ax == 1 ? wi * wk / 4 : wj * wk / 4);      // face_weight()
                    const FaceRec* fr = reinterpret_cast<const FaceRec*>(rec) + s;
                    const double Nv[3] = {fr->vec[0] * fw, fr->vec[1] * fw, fr->vec[2] * fw};
                    double out[7];
                    trace_apply(K, Nv, out);
                    double* dst = sTr + s * TBS + slot;
#pragma unroll
                    for (int c = 0; c < 7; c++) dst[c * TCS] = out[c];
                }
            }
        }
        fence_async_smem();
        __syncthreads();                                                                   // (3)
        if (tid == issuer) {
            bulk_s2g(P.traceA + (size_t)elem * 6 * TBS, sTr, (uint32_t)(6 * TBS * sizeof(double)));
            bulk_commit();
        }
        if (!hasNext) break;
        seq = nxt;
        st ^= 1;
        rs = rs1;
    }
    if (tid == issuer) bulk_wait0();
}

// ---------------------------------------------------------------------------------------------------
// sweep B (v4)
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, bool VISC, bool TRI, int MINB>
__global__ void __launch_bounds__((Cfg<NX, NY, NZ, VISC, TRI>::NT), MINB) sweepB_v4(const __grid_constant__ KParams P) {
    using C = Cfg<NX, NY, NZ, VISC, TRI>;
    using Tk = Tasks<NX, NY, NZ>;
    constexpr int NP = C::NP, NPS = C::NPS, NFT = C::NFT, NT = C::NT, NIN = C::NIN_B, TBS = C::TBS, NPF = C::NPF, STG = C::STG_B;
    constexpr int A_RO = C::B_RO, A_RN = C::B_RN, A_U = C::B_U, A_T = C::B_T, A_P = C::B_P, A_GU = C::B_GU, A_GT = C::B_GT, A_RR = C::B_RR,
                  A_J = C::B_J, A_CV = C::B_CV;
    constexpr int TCS = trace_cs(NPF);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const sm = reinterpret_cast<double*>(smem_raw);
    double* const sIn = sm + C::oInB;        // [2]{[NIN][NPS], [6][TBS]}
    double* const sRec = sm + C::oRecB;      // [3][RECD]
    double* const sF = sm + C::oFB;          // [4][NFT]
    double* const sD = sm + C::oDB;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + C::oBarB);   // full[2], rec[3]

    const int tid = threadIdx.x;
    const uint32_t stride = gridDim.x;
    uint32_t seq = blockIdx.x;
    if (seq >= P.nB) return;

    const bool nodeT = tid < NP;
    const int nt = nodeT ? tid : 0;
    const int i = nt / (NY * NZ), j = (nt / NZ) % NY, k = nt % NZ;
    const double wcv = ((P.W[0][i] * P.W[1][j]) * P.W[2][k]) / 8;
    const double xi0 = P.X[0][i], xi1 = P.X[1][j], xi2 = P.X[2][k];
    const bool faceT = tid < NFT;
    int fs = 0, fa = 0, fb = 0;
    Tk::decode(faceT ? tid : 0, fs, fa, fb);
    const int fslot = (fs < 2) ? fa * NY + fb : fa * NZ + fb;
    const int fln = face_node<NX, NY, NZ>(fs, fa, fb);
    const double fw = face_weight<NX, NY, NZ>(P, fs, fa, fb);

    auto elem_of = [&](uint32_t q) -> uint32_t { return P.sched ? P.sched[q] : q; };
    auto issue_rec = [&](int slot, uint32_t elem) {
        mbar_expect_tx(&bars[2 + slot], (uint32_t)sizeof(ElemRec));
        bulk_g2s(sRec + slot * RECD, P.elemRec + elem, (uint32_t)sizeof(ElemRec), &bars[2 + slot]);
    };
    // arrays of `elem` and the six neighbour traces named by the record in `slot` (which has landed)
    auto issue_stage = [&](int st, uint32_t elem, int slot) {
        uint64_t* bar = &bars[st];
        double* dst = sIn + (size_t)st * STG;
        const size_t off = (size_t)elem * NPS;
        constexpr uint32_t B = NPS * sizeof(double);
        mbar_expect_tx(bar, (uint32_t)(NIN * B + 6 * TBS * sizeof(double)));
        const FaceRec* fr = reinterpret_cast<const FaceRec*>(sRec + slot * RECD);
#pragma unroll
        for (int f = 0; f < 6; f++)
            bulk_g2s(dst + NIN * NPS + f * TBS, P.traceA + (size_t)fr[f].otherBlock * TBS, (uint32_t)(TBS * sizeof(double)), bar);
        bulk_g2s(dst + A_RO * NPS, P.rho_old + off, B, bar);
        bulk_g2s(dst + A_RN * NPS, P.rho_new + off, B, bar);
        bulk_g2s(dst + (A_U + 0) * NPS, P.U_old[0] + off, B, bar);
        bulk_g2s(dst + (A_U + 1) * NPS, P.U_old[1] + off, B, bar);
        bulk_g2s(dst + (A_U + 2) * NPS, P.U_old[2] + off, B, bar);
        bulk_g2s(dst + A_T * NPS, P.T_old + off, B, bar);
        bulk_g2s(dst + A_P * NPS, P.p + off, B, bar);
        if (VISC) {
#pragma unroll
            for (int q = 0; q < 9; q++) bulk_g2s(dst + (A_GU + q) * NPS, P.GU[q] + off, B, bar);
#pragma unroll
            for (int q = 0; q < 3; q++) bulk_g2s(dst + (A_GT + q) * NPS, P.GT[q] + off, B, bar);
        }
        bulk_g2s(dst + A_RR * NPS, P.rho_ref + off, B, bar);
        if (!TRI) {
#pragma unroll
            for (int q = 0; q < 9; q++) bulk_g2s(dst + (A_J + q) * NPS, P.Jinv[q] + off, B, bar);
            bulk_g2s(dst + A_CV * NPS, P.cV + off, B, bar);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 5; q++) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == C::ISSUER) {
        issue_rec(0, elem_of(seq));
        if (seq + stride < P.nB) issue_rec(1, elem_of(seq + stride));
    }
    for (int q = tid; q < 3 * MAXN * MAXN; q += NT) sD[q] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];
    mbar_wait(&bars[2], 0);
    if (tid == C::ISSUER) issue_stage(0, elem_of(seq), 0);

    constexpr int NODE_THREADS = pad_to(NP, 32);      // warps beyond these only run face tasks (and the issuer)
    int st = 0, rs = 0;
    uint32_t it = 0;
    for (;;) {
        const uint32_t elem = elem_of(seq);
        const uint32_t nxt = seq + stride, nxt2 = nxt + stride;
        const bool hasNext = nxt < P.nB;
        const int rs1 = (rs == 2) ? 0 : rs + 1, rs2 = (rs1 == 2) ? 0 : rs1 + 1;
        if (it > 0) mbar_wait(&bars[2 + rs], (it / 3) & 1u);       // this element's record (requested two elements ago)
        mbar_wait(&bars[st], (it >> 1) & 1u);
        double* const in = sIn + (size_t)st * STG;
        double* const sT = in + NIN * NPS;                         // [6][TBS] neighbour-side traces, later the four outputs [4][NPS]
        double* const sH = VISC ? in + A_GU * NPS : sm + C::oHB;   // [12][NPS] contravariant fluxes (in place over the gradients)
        const double* const rec = sRec + rs * RECD;

        // ---- face tasks first: they need the gradients at the face nodes, which the node pass overwrites ----
        if (faceT) {
            const FaceRec* fr = reinterpret_cast<const FaceRec*>(rec) + fs;
            const uint32_t meta = fr->meta;
            const uint32_t fid = meta & FM_FID_MASK;
            const bool own = meta & FM_OWNER;
            const double al = (meta & FM_HALF) ? 0.5 : 0.0;
            const double N[3] = {fr->vec[0] * fw, fr->vec[1] * fw, fr->vec[2] * fw};
            const double fu[3] = {fr->unit[0], fr->unit[1], fr->unit[2]};
            const double nN = fu[0] * N[0] + fu[1] * N[1] + fu[2] * N[2];
            SideState me;
            me.rho_o = in[A_RO * NPS + fln]; me.rho_n = in[A_RN * NPS + fln];
            me.u[0] = in[A_U * NPS + fln]; me.u[1] = in[(A_U + 1) * NPS + fln]; me.u[2] = in[(A_U + 2) * NPS + fln];
            me.th = in[A_T * NPS + fln] + P.T0;
            me.pp = in[A_P * NPS + fln];
            if (VISC) {
#pragma unroll
                for (int c = 0; c < 9; c++) me.gU[c] = in[(A_GU + c) * NPS + fln];
#pragma unroll
                for (int c = 0; c < 3; c++) me.gT[c] = in[(A_GT + c) * NPS + fln];
            }
            double mt[7];
            side_trace(me, N, P.nu, P.iPr, P.gamma * P.R, VISC, mt);
            // the other side: its face trace, slot = the same (a,b) in ITS face numbering (ghost blocks use mine)
            const int oslot = (fid == FM_GHOST) ? fslot : ((fid < 2) ? fa * NY + fb : fa * NZ + fb);
            const double* xt = sT + fs * TBS + oslot;
            const double wo = own ? al : 1 - al, wx = own ? 1 - al : al;
            const double lam = (mt[6] * wo + xt[6 * TCS] * wx) / 2;
            const double sg = own ? 1.0 : -1.0;
            const double dqN = xt[4 * TCS] - mt[4];
            const double dqT = xt[5 * TCS] - mt[5];
            double* out = &sF[tid];
#pragma unroll
            for (int c = 0; c < 3; c++) out[c * NFT] = sg * ((mt[c] * wo + xt[c * TCS] * wx) - fu[c] * (lam * (sg * dqN)));
            out[3 * NFT] = sg * ((mt[3] * wo + xt[3 * TCS] * wx) - lam * (sg * dqT) * nN);
        }
        // (1) face results ready, gradients at the face nodes consumed: face-only warps do not wait
        if (NODE_THREADS == NT || tid < NODE_THREADS) bar_sync(1, NT); else bar_arrive(1, NT);
        if (tid == C::ISSUER) {
            bulk_wait_read0();             // the previous element's stores have left the other stage's trace block
            if (hasNext) {
                // the next element's record (requested a whole element ago) names the trace blocks to fetch
                mbar_wait(&bars[2 + rs1], ((it + 1) / 3) & 1u);
                issue_stage(st ^ 1, elem_of(nxt), rs1);
            }
            if (nxt2 < P.nB) issue_rec(rs2, elem_of(nxt2));
        }

        double rho_o = 0, rho_nw = 1, u[3] = {0, 0, 0}, th = 0, cV = 1, rref = 0;
        if (nodeT) {
            rho_o = in[A_RO * NPS + nt]; rho_nw = in[A_RN * NPS + nt];
            u[0] = in[A_U * NPS + nt]; u[1] = in[(A_U + 1) * NPS + nt]; u[2] = in[(A_U + 2) * NPS + nt];
            th = in[A_T * NPS + nt] + P.T0;
            const double pp = in[A_P * NPS + nt];
            rref = in[A_RR * NPS + nt];
            const double mu = VISC ? rho_o * P.nu : 0.0;
            double Jin[9];
            if (TRI) {
                tri_metrics(rec + 48, xi0, xi1, xi2, wcv, Jin, cV);
            } else {
                cV = in[A_CV * NPS + nt];
#pragma unroll
                for (int c = 0; c < 9; c++) Jin[c] = in[(A_J + c) * NPS + nt] * cV;
            }
            const double Fc[3] = {rho_o * u[0], rho_o * u[1], rho_o * u[2]};
            double H[12];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                double fq[3];
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    fq[b] = Fc[a] * u[b] + (a == b ? pp : 0.0);
                    if (VISC) fq[b] -= mu * in[(A_GU + a * 3 + b) * NPS + nt];
                }
#pragma unroll
                for (int d = 0; d < 3; d++) H[a * 3 + d] = fq[0] * Jin[d] + fq[1] * Jin[3 + d] + fq[2] * Jin[6 + d];
            }
            {
                double fq[3];
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    fq[b] = Fc[b] * th;
                    if (VISC) fq[b] -= (mu * P.iPr) * in[(A_GT + b) * NPS + nt];
                }
#pragma unroll
                for (int d = 0; d < 3; d++) H[9 + d] = fq[0] * Jin[d] + fq[1] * Jin[3 + d] + fq[2] * Jin[6 + d];
            }
#pragma unroll
            for (int c = 0; c < 12; c++) sH[c * NPS + nt] = H[c];
        }
        if (tid < NODE_THREADS) bar_sync(2, NODE_THREADS);                                 // (2) node warps only

        if (nodeT) {
            double r[4];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                double acc = 0;
#pragma unroll
                for (int ii = 0; ii < NX; ii++) acc += sH[(a * 3 + 0) * NPS + ii * NY * NZ + j * NZ + k] * sD[0 * MAXN * MAXN + ii * NX + i];
#pragma unroll
                for (int jj = 0; jj < NY; jj++) acc += sH[(a * 3 + 1) * NPS + i * NY * NZ + jj * NZ + k] * sD[1 * MAXN * MAXN + jj * NY + j];
#pragma unroll
                for (int kk = 0; kk < NZ; kk++) acc += sH[(a * 3 + 2) * NPS + i * NY * NZ + j * NZ + kk] * sD[2 * MAXN * MAXN + kk * NZ + k];
                r[a] = -acc;
            }
#pragma unroll
            for (int ax = 0; ax < 3; ax++) {
                const int cx = (ax == 0) ? k : (ax == 1 ? j : i);
                const int nx = (ax == 0) ? NZ : (ax == 1 ? NY : NX);
                if (cx != 0 && cx != nx - 1) continue;
                const int s = 2 * ax + (cx != 0 ? 1 : 0);
                const int ti = (ax == 0) ? Tk::index(s, i, j) : (ax == 1 ? Tk::index(s, i, k) : Tk::index(s, j, k));
#pragma unroll
                for (int c = 0; c < 4; c++) r[c] += sF[c * NFT + ti];
            }
            const double ap0 = (-1.0 / P.dt) * cV;
            const double ap = ap0 * rho_nw;
            double g[3] = {P.g[0], P.g[1], P.g[2]};
            if (P.has_gfield) {
                const size_t idx = (size_t)elem * NPS + nt;
                g[0] = P.gfield[0][idx]; g[1] = P.gfield[1][idx]; g[2] = P.gfield[2][idx];
            }
            const double drho = P.buoyancy ? (rho_nw - rref) : 0.0;
            const double rap = 1.0 / ap;                 // x = Su / ap (solve.cpp:563-570) as one reciprocal and 4 products
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double Su = (r[c] - (drho * g[c]) * cV) + (u[c] * rho_o) * ap0;
                sT[c * NPS + nt] = Su * rap;
            }
            {
                const double Su = r[3] + (th * rho_o) * ap0;
                sT[3 * NPS + nt] = Su * rap - P.T0;
            }
        } else if (tid < NPS) {
#pragma unroll
            for (int c = 0; c < 4; c++) sT[c * NPS + tid] = 0.0;
        }
        fence_async_smem();
        __syncthreads();                                                                   // (3)
        if (tid == C::ISSUER) {
            const size_t off = (size_t)elem * NPS;
            constexpr uint32_t B = NPS * sizeof(double);
            bulk_s2g(P.U_new[0] + off, sT + 0 * NPS, B);
            bulk_s2g(P.U_new[1] + off, sT + 1 * NPS, B);
            bulk_s2g(P.U_new[2] + off, sT + 2 * NPS, B);
            bulk_s2g(P.T_new + off, sT + 3 * NPS, B);
            bulk_commit();
        }
        if (!hasNext) break;
        seq = nxt;
        st ^= 1;
        rs = rs1;
        it++;
    }
    if (tid == C::ISSUER) bulk_wait0();
}

}  // namespace v4
}  // namespace nsem
