// nsem_kernels_v4.cuh -- persistent, software-pipelined sm_100a sweeps (cubic 3-D orders).
//
// Same mathematics as nsem_kernels.cuh (v1) / nsem_kernels_v2.cuh; what changes is how the SM is kept busy:
//   * PERSISTENT CTAs (grid = SMs x resident CTAs) loop over elements; while element e is being computed the
//     bulk-async copies (cp.async.bulk, SASS UBLKCP, completion on an mbarrier) of element e+1 are already in
//     flight into the other half of a two-stage shared-memory ring, so no warp waits for HBM latency;
//   * the 560-byte element record (six face records + the element's trilinear map) runs two elements ahead in a
//     three-slot ring, so the neighbour values of sweep A (8-byte cp.async, LDGSTS) and the six neighbour face
//     traces of sweep B (bulk copies) are requested a full element early;
//   * ONE THREAD PER NODE and nothing else: NT = NP rounded up to whole warps, every warp does the same work.  A node
//     that lies on a face evaluates the Rusanov flux there itself (its side from registers, the other side from the
//     gathered values / the neighbour's trace), so there are no face tasks, no face-result round trip through
//     shared memory and only two CTA barriers per element;
//   * METRICS ON THE FLY (TRI = true): for straight-edged hexahedra (every non-curved mesh: dg.cpp:257-263
//     interpolates the nodes trilinearly) Jinv*cV at a node is a closed form of the element's 7 trilinear
//     coefficient vectors and its volume, so the 10 per-node metric arrays (Jinv, cV) are neither stored nor
//     streamed; nsem_upload_mesh verifies the closed form against the uploaded Jinv/cV at every node and keeps
//     the stored-metric instantiation (TRI = false) for curved meshes;
//   * node results leave from registers with coalesced stores; the element's six face-trace blocks are collected in
//     shared memory and leave as one dense bulk-async store (full 128-byte lines);
//   * the bulk copies of a stage are issued by lane 31 of up to four warps, each owning every fourth copy.
#pragma once
#include "nsem_kernels_v2.cuh"

// differentiation-matrix entry: straight from the kernel-parameter constant bank (30 fewer shared-memory loads per node in sweep A: -3 %,
// profiles/r1_variants.md), or (-DNSEM_D_SMEM) from the shared-memory copy
#ifndef NSEM_D_SMEM
#define NSEM_DM(d, idx) P.D[d][idx]
#else
#define NSEM_DM(d, idx) sD[(d) * MAXN * MAXN + (idx)]
#endif

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

// metrics of one node from the element record (rec + 48 = ElemRec::c): the closed trilinear form, or -- the branch is uniform over the
// CTA -- the per-element constants of a parallelepiped (box meshes, the bulk of a terrain-following mesh away from the hill)
__device__ __forceinline__ void elem_metrics(const double* __restrict__ rec, double x0, double x1, double x2, double wcv, double Jin[9], double& cV) {
    const double* c = rec + 48;
    if (reinterpret_cast<const FaceRec*>(rec)->meta & FM_AFFINE) {
        cV = c[21] * wcv;
#pragma unroll
        for (int q = 0; q < 9; q++) Jin[q] = c[9 + q] * wcv;
    } else {
        tri_metrics(c, x0, x1, x2, wcv, Jin, cV);
    }
}

// ---------------------------------------------------------------------------------------------------
// sweep A (v4): one thread per node (NT = NP rounded up to whole warps), no separate face tasks.  The neighbour values of
// the 2(NX NY + NX NZ + NY NZ) face nodes are requested one element ahead with 8-byte cp.async (LDGSTS) into a
// double-buffered table; the node that lies on a face evaluates the Rusanov mass flux and the gradient jumps there
// itself, adds them in local-face-id order, and publishes its side's face trace for sweep B of the neighbour.
// Results leave from registers (coalesced stores); the element's six trace blocks are collected in shared memory and
// stored by one bulk-async copy.
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, bool VISC, bool TRI>
struct CfgA {
    using Dm = Dims<NX, NY, NZ>;
    using Tk = Tasks<NX, NY, NZ>;
    static constexpr int NP = Dm::NP, NPS = Dm::NPS, NPF = Dm::NPF, NFT = Tk::NFT;
    static constexpr int NT = pad_to(NP, 32);
    static constexpr int NISS = (NT / 32) < 4 ? (NT / 32) : 4;          // issuing threads: lane 31 of the first warps
    static constexpr int NPASS = (NFT + NT - 1) / NT;                   // gather requests per thread
    static constexpr bool ok = true;
    static constexpr int TBS = trace_bs(NPF);
    static constexpr int NFTP = pad_to(NFT, 2);
    static constexpr int NIN = TRI ? 7 : 17;                            // rho, U(3), T, p_ref, S, [Jinv(9), cV]
    static constexpr int A_PREF = 5, A_S = 6, A_J = 7, A_CV = 16;
    static constexpr int NGV = 6;                                       // gathered per face node: rho, U(3), T, S
    static constexpr int oIn = 0;
    static constexpr int oRec = oIn + 2 * NIN * NPS;
    static constexpr int oG = oRec + 3 * RECD;                          // [2][NGV][NFTP] neighbour values of the face nodes
    static constexpr int oR = oG + 2 * NGV * NFTP;                      // [4][NP]: contravariant mass flux (3), theta
    static constexpr int oTr = oR + pad_to(4 * NP, 2);                  // [6][TBS]
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

// Order in which one persistent CTA walks the elements: RUNS of `run` consecutive schedule positions, the runs dealt round-robin to the
// CTAs (run c of CTA b starts at position (c * gridDim + b) * run).  With run = 1 this is the plain grid-stride loop.  On a structured
// mesh consecutive elements are k-neighbours, so with run > 1 the element across a k-face is the one this CTA handled a moment ago (or
// handles next): its values are in L2 (or in this CTA's own shared-memory ring) instead of being fetched a second time from DRAM while
// another CTA streams them.
struct RunIter {
    uint32_t base, r;          // start of the current run, offset inside it
    bool valid;
    __device__ __forceinline__ uint32_t pos() const { return base + r; }
    __device__ __forceinline__ static RunIter first(uint32_t nB, uint32_t run) {
        RunIter it;
        it.base = blockIdx.x * run; it.r = 0; it.valid = it.base < nB;
        return it;
    }
    __device__ __forceinline__ RunIter next(uint32_t nB, uint32_t run) const {
        RunIter n = *this;
        if (!valid) return n;
        if (r + 1 < run && base + r + 1 < nB) { n.r = r + 1; return n; }
        n.base = base + gridDim.x * run; n.r = 0; n.valid = n.base < nB;
        return n;
    }
};

// Halo fused into the sweep: every CTA reports ONCE, when it is past the partition-boundary elements (they come first in the schedule) or
// has run out of work; the last one to report publishes the exchange's epoch in the neighbours' flag words.  All threads call it.
__device__ __forceinline__ void halo_report(const KParams& P) {
    __syncthreads();                                   // every thread's stores into the neighbours' windows have been issued
    if (threadIdx.x == 0) {
        __threadfence_system();
        const HaloFuse* H = P.halo;
        if (atomicAdd(H->counter, 1u) == gridDim.x - 1) {
            *H->counter = 0;
            __threadfence_system();
            for (int p = 0; p < H->npeers; p++)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(H->flag[p]), "l"(P.haloEpoch) : "memory");
        }
    }
}

// MORTAR: the mesh has non-conforming faces (FM_MORTAR): such a face takes its finished surface terms from the mortar buffers
// (nsem_mortar.cuh) instead of a two-point flux.  A separate instantiation, so conforming meshes run exactly the code they ran before.
template <int NX, int NY, int NZ, bool VISC, bool TRI, int MINB, bool MORTAR = false>
__global__ void __launch_bounds__((CfgA<NX, NY, NZ, VISC, TRI>::NT), MINB) sweepA_v4(const __grid_constant__ KParams P) {
    using C = CfgA<NX, NY, NZ, VISC, TRI>;
    using Tk = Tasks<NX, NY, NZ>;
    constexpr int NP = C::NP, NPS = C::NPS, NFT = C::NFT, NT = C::NT, NIN = C::NIN, TBS = C::TBS, NPF = C::NPF, NISS = C::NISS;
    constexpr int NFTP = C::NFTP, TCS = trace_cs(NPF), NGV = C::NGV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const sm = reinterpret_cast<double*>(smem_raw);
    double* const sIn = sm + C::oIn;         // [2][NIN][NPS]
    double* const sRec = sm + C::oRec;       // [3][RECD]
    double* const sG = sm + C::oG;           // [2][NGV][NFTP]
    double* const sR = sm + C::oR;           // [4][NP]
    double* const sTr = sm + C::oTr;         // [6][TBS]
    double* const sD = sm + C::oD;           // [3][MAXN*MAXN]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + C::oBar);   // full[2], rec[3]

    const int tid = threadIdx.x;
    const uint32_t run = P.run ? P.run : 1u;
    RunIter cur = RunIter::first(P.nB, run);
    if (!cur.valid) {
        if (P.halo) halo_report(P);        // a CTA without work still counts
        return;
    }
    const int iss = ((tid & 31) == 31 && (tid >> 5) < NISS) ? (tid >> 5) : -1;

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
        const int mine = (NIN - iss + NISS - 1) / NISS;          // copies q = iss, iss + NISS, ... < NIN
        mbar_expect_tx(bar, (uint32_t)(mine * B));
        for (int q = iss; q < NIN; q += NISS) bulk_g2s(dst + q * NPS, P.srcA[q] + off, B, bar);
    };
    // request the neighbour values of every face node of the element whose record sits in `slot` into table `buf`
    auto issue_gathers = [&](int slot, int buf) {
#ifdef NSEM_EXP_NOGATHER       // timing experiment only (wrong results): what sweep A costs without its neighbour gathers
        if (slot >= 0) { cp_async_commit(); return; }
#endif
        const int t = fresh_tid();
#pragma unroll
        for (int ps = 0; ps < C::NPASS; ps++) {
            const int task = t + ps * NT;
            if (task < NFT) {
                int fs, fa, fb;
                Tk::decode(task, fs, fa, fb);
                const FaceRec* fr = reinterpret_cast<const FaceRec*>(sRec + slot * RECD) + fs;
                const uint32_t other = fr->other, fid = fr->meta & FM_FID_MASK;
                if (MORTAR && (fr->meta & FM_MORTAR)) continue;          // `other` is a mortar block id there, nothing to gather
                const int fslot = (fs < 2) ? fa * NY + fb : fa * NZ + fb;
                const size_t oidx = (size_t)other + (fid == FM_GHOST ? fslot : face_node<NX, NY, NZ>(fid, fa, fb));
                double* g = sG + buf * NGV * NFTP + task;
                cp_async8(g + 0 * NFTP, P.rho_old + oidx);
                cp_async8(g + 1 * NFTP, P.U_old[0] + oidx);
                cp_async8(g + 2 * NFTP, P.U_old[1] + oidx);
                cp_async8(g + 3 * NFTP, P.U_old[2] + oidx);
                cp_async8(g + 4 * NFTP, P.T_old + oidx);
                cp_async8(g + 5 * NFTP, P.S_old + oidx);
            }
        }
        cp_async_commit();
    };

    if (tid == 0) {
        mbar_init(&bars[0], NISS); mbar_init(&bars[1], NISS);
#pragma unroll
        for (int q = 2; q < 5; q++) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    RunIter nx1 = cur.next(P.nB, run);
    if (iss == 0) {
        issue_rec(0, elem_of(cur.pos()));
        if (nx1.valid) issue_rec(1, elem_of(nx1.pos()));
    }
    if (iss >= 0) issue_arrays(0, elem_of(cur.pos()));
    for (int q = tid; q < 3 * MAXN * MAXN; q += NT) sD[q] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];
    mbar_wait(&bars[2], 0);
    issue_gathers(0, 0);

    int st = 0, rs = 0;                    // stage (= gather table) / record slot of the current element
    uint32_t it = 0;
    bool reported = (P.halo == nullptr);
    for (;;) {
        if (!reported && cur.pos() >= P.nHalo) { halo_report(P); reported = true; }
        const uint32_t elem = elem_of(cur.pos());
        const RunIter nx2 = nx1.next(P.nB, run);
        const bool hasNext = nx1.valid;
        const int rs1 = (rs == 2) ? 0 : rs + 1, rs2 = (rs1 == 2) ? 0 : rs1 + 1;
        if (iss >= 0) {
            if (hasNext) issue_arrays(st ^ 1, elem_of(nx1.pos()));
            if (iss == 0 && hasNext && nx2.valid) issue_rec(rs2, elem_of(nx2.pos()));
            if (iss == NISS - 1) bulk_wait_read0();       // the previous element's trace blocks have left sTr
        }
        if (it > 0) mbar_wait(&bars[2 + rs], (it / 3) & 1u);           // this element's record
        if (hasNext) {
            // the next element's record was requested a whole element ago; its neighbour values land while this element runs
            mbar_wait(&bars[2 + rs1], ((it + 1) / 3) & 1u);
            issue_gathers(rs1, st ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");       // this element's neighbour values (requested an element ago)
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        mbar_wait(&bars[st], (it >> 1) & 1u);
        const double* const in = sIn + (size_t)st * NIN * NPS;
        const double* const gx = sG + st * NGV * NFTP;
        const double* const rec = sRec + rs * RECD;

        // ---- node: contravariant mass flux, theta ----
        const int nt = fresh_tid();
        const bool nodeT = nt < NP;
        const int i = nt / (NY * NZ), j = (nt / NZ) % NY, k = nt % NZ;
        double rho = 0, u0 = 0, u1 = 0, u2 = 0, th = 0, cV = 1, Jin[9];
        if (nodeT) {
            rho = in[0 * NPS + nt];
            u0 = in[1 * NPS + nt]; u1 = in[2 * NPS + nt]; u2 = in[3 * NPS + nt];
            th = in[4 * NPS + nt] + P.T0;
            if (TRI) {
                const double wcv = ((P.W[0][i] * P.W[1][j]) * P.W[2][k]) / 8;
                elem_metrics(rec, P.X[0][i], P.X[1][j], P.X[2][k], wcv, Jin, cV);
            } else {
                cV = in[C::A_CV * NPS + nt];
#pragma unroll
                for (int c = 0; c < 9; c++) Jin[c] = in[(C::A_J + c) * NPS + nt] * cV;
            }
            const double F0 = u0 * rho, F1 = u1 * rho, F2 = u2 * rho;
#pragma unroll
            for (int d = 0; d < 3; d++) sR[d * NP + nt] = F0 * Jin[d] + F1 * Jin[3 + d] + F2 * Jin[6 + d];
            sR[3 * NP + nt] = th;
        }
        __syncthreads();                                   // (1) mass flux and theta of all nodes; everybody's gathers have landed

        if (nodeT) {
            double r_rho, gU[9], gT[3];
            {
                double acc = 0;         // (one chain: three chains, one per direction, measured 2 % slower, profiles/r2_variants.md)
#pragma unroll
                for (int ii = 0; ii < NX; ii++) acc += sR[0 * NP + ii * NY * NZ + j * NZ + k] * NSEM_DM(0, ii * NX + i);
#pragma unroll
                for (int jj = 0; jj < NY; jj++) acc += sR[1 * NP + i * NY * NZ + jj * NZ + k] * NSEM_DM(1, jj * NY + j);
#pragma unroll
                for (int kk = 0; kk < NZ; kk++) acc += sR[2 * NP + i * NY * NZ + j * NZ + kk] * NSEM_DM(2, kk * NZ + k);
                r_rho = -acc;
            }
            if (VISC) {
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    const double* q = (f < 3) ? in + (1 + f) * NPS : sR + 3 * NP;
                    double d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
                    for (int m = 0; m < NX; m++) d0 += NSEM_DM(0, i * NX + m) * q[m * NY * NZ + j * NZ + k];
#pragma unroll
                    for (int m = 0; m < NY; m++) d1 += NSEM_DM(1, j * NY + m) * q[i * NY * NZ + m * NZ + k];
#pragma unroll
                    for (int m = 0; m < NZ; m++) d2 += NSEM_DM(2, k * NZ + m) * q[i * NY * NZ + j * NZ + m];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        const double v = Jin[a * 3 + 0] * d0 + Jin[a * 3 + 1] * d1 + Jin[a * 3 + 2] * d2;
                        if (f < 3) gU[a * 3 + f] = v; else gT[a] = v;
                    }
                }
            }

            // ---- faces this node lies on, in local-face-id order (field.h:3093-3114): k-faces 0/1, j-faces 2/3, i-faces 4/5 ----
            const bool onK = (k == 0 || k == NZ - 1), onJ = (j == 0 || j == NY - 1), onI = (i == 0 || i == NX - 1);
            const bool onAny = onK || onJ || onI;
            const double wi = P.W[0][i], wj = P.W[1][j], wk = P.W[2][k];
            double S = 0;
            if (onAny) {
                S = in[C::A_S * NPS + nt];                              // |U| + c of this node (written with the state): lambdaMax here, trace below
#pragma unroll
                for (int ax = 0; ax < 3; ax++) {
                    const bool on = (ax == 0) ? onK : (ax == 1 ? onJ : onI);
                    if (!on) continue;
                    const int cx = (ax == 0) ? k : (ax == 1 ? j : i);
                    const int s = 2 * ax + (cx != 0 ? 1 : 0);
                    const int ti = (ax == 0) ? Tk::index(s, i, j) : (ax == 1 ? Tk::index(s, i, k) : Tk::index(s, j, k));
                    const double fw = (ax == 0) ? wi * wj / 4 : (ax == 1 ? wi * wk / 4 : wj * wk / 4);      // face_weight()
                    const FaceRec* fr = reinterpret_cast<const FaceRec*>(rec) + s;
                    const uint32_t meta = fr->meta;
                    if (MORTAR && (meta & FM_MORTAR)) {
                        // non-conforming face: mortarA_kernel left this node's surface terms (scatter/gather_non_conforming, field.h:2019-2248)
                        const int mslot = (ax == 0) ? i * NY + j : (ax == 1 ? i * NZ + k : j * NZ + k);
                        const double* mc = P.mortarA + (size_t)fr->other * (MORTAR_NA * MORTAR_MAXF) + mslot;
                        r_rho += mc[0];
                        if (VISC) {
#pragma unroll
                            for (int c = 0; c < 9; c++) gU[c] += mc[(1 + c) * MORTAR_MAXF];
#pragma unroll
                            for (int c = 0; c < 3; c++) gT[c] += mc[(10 + c) * MORTAR_MAXF];
                        }
                        continue;
                    }
                    const double xr = gx[0 * NFTP + ti], xu0 = gx[1 * NFTP + ti], xu1 = gx[2 * NFTP + ti], xu2 = gx[3 * NFTP + ti];
                    const double xth = gx[4 * NFTP + ti] + P.T0;
                    // written for "my side" / "other side": with fI in {0, 1/2} this is bitwise cds() = fI*owner + (1-fI)*neighbour
                    const bool own = meta & FM_OWNER;
                    const double al = (meta & FM_HALF) ? 0.5 : 0.0;
                    const double wo = own ? al : 1 - al, wx = own ? 1 - al : al;       // weight of my side / the other side
                    const double sg = own ? 1.0 : -1.0;                                // (q_n - q_o) = sg * (q_other - q_mine)
                    const double N0 = fr->vec[0] * fw, N1 = fr->vec[1] * fw, N2 = fr->vec[2] * fw;      // fN[k] = gFN * w_a w_b / 4
                    const double nN = fr->unit[0] * N0 + fr->unit[1] * N1 + fr->unit[2] * N2;           // unit(fN).fN
                    const double lam = (S * wo + gx[5 * NFTP + ti] * wx) / 2;                             // cds(|U| + c) / 2
                    const double fm = rho * (u0 * N0 + u1 * N1 + u2 * N2), fx = xr * (xu0 * N0 + xu1 * N1 + xu2 * N2);
                    const double flux = (fm * wo + fx * wx) - lam * (sg * (xr - rho)) * nN;
                    r_rho += sg * flux;
                    if (P.op_flux) P.op_flux[((size_t)elem * 6 + s) * NPF + ((ax == 0) ? i * NY + j : (ax == 1 ? i * NZ + k : j * NZ + k))] = flux;   // nsem_op_rusanov
                    if (VISC) {
                        // grad_flux<strong>: r += (+-fN) (x) (cds(q) - q_mine)
                        const double q0_ = (u0 * wo + xu0 * wx) - u0, q1_ = (u1 * wo + xu1 * wx) - u1, q2_ = (u2 * wo + xu2 * wx) - u2;
                        const double q3_ = (th * wo + xth * wx) - th;
                        const double sN[3] = {sg * N0, sg * N1, sg * N2};
#pragma unroll
                        for (int aa = 0; aa < 3; aa++) {
                            gU[aa * 3 + 0] += sN[aa] * q0_;
                            gU[aa * 3 + 1] += sN[aa] * q1_;
                            gU[aa * 3 + 2] += sN[aa] * q2_;
                            gT[aa] += sN[aa] * q3_;
                        }
                    }
                }
            }

            // ---- updates: (Su + rho*ap0) / ap0 with ap0 = (-1/dt) cV (addTemporal<1>, SolveTexplicit) as ONE reciprocal of cV ----
            const size_t idx = (size_t)elem * NPS + nt;
            const double rcV = 1.0 / cV;
            const double ap0 = P.mrdt * cV;
            const double rho_new = (P.op_mode & 1) ? r_rho : (r_rho + rho * ap0) * (rcV * P.mdt);
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
            // ---- partition boundary: this node's values go straight into the neighbour's receive window (NVLink stores) ----
            if (P.halo && onAny) {
#pragma unroll
                for (int ax = 0; ax < 3; ax++) {
                    const bool on = (ax == 0) ? onK : (ax == 1 ? onJ : onI);
                    if (!on) continue;
                    const int cx = (ax == 0) ? k : (ax == 1 ? j : i);
                    const FaceRec* fr = reinterpret_cast<const FaceRec*>(rec) + (2 * ax + (cx != 0 ? 1 : 0));
                    if ((fr->meta & FM_FID_MASK) != FM_GHOST || (MORTAR && (fr->meta & FM_MORTAR))) continue;
                    const uint32_t g = (uint32_t)((fr->other - P.ghostBase) / Dims<NX, NY, NZ>::GPS);
                    const uint32_t peer = P.haloGhost[2 * g];
                    if (peer == 0xffffffffu) continue;
                    const int slot = (ax == 0) ? i * NY + j : (ax == 1 ? i * NZ + k : j * NZ + k);
                    double* w = P.halo->win[peer] + P.haloGhost[2 * g + 1] + slot;
                    const uint64_t hs = P.halo->stride[peer];
                    w[0] = rho_new; w[hs] = ppn;
                    if (VISC) {
#pragma unroll
                        for (int c = 0; c < 9; c++) w[(2 + c) * hs] = gU[c];
#pragma unroll
                        for (int c = 0; c < 3; c++) w[(11 + c) * hs] = gT[c];
                    }
                }
            }
            // ---- this side's face traces for sweep B of the neighbours (and of the peers behind a partition boundary) ----
            if (onAny) {
                SideState q;
                q.rho_o = rho; q.rho_n = rho_new; q.th = th; q.pp = ppn;
                q.u[0] = u0; q.u[1] = u1; q.u[2] = u2;
                if (VISC) {
#pragma unroll
                    for (int c = 0; c < 9; c++) q.gU[c] = gU[c];
#pragma unroll
                    for (int c = 0; c < 3; c++) q.gT[c] = gT[c];
                }
                TraceCoef K;
                trace_coef(q, S, P.nu, P.iPr, VISC, K);
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
        __syncthreads();                                   // (2) trace blocks complete; stage and gather table free
        if (iss == NISS - 1) {
            bulk_s2g(P.traceA + (size_t)elem * 6 * TBS, sTr, (uint32_t)(6 * TBS * sizeof(double)));
            bulk_commit();
        }
        if (!hasNext) break;
        cur = nx1;
        nx1 = nx2;
        st ^= 1;
        rs = rs1;
        it++;
    }
    if (iss == NISS - 1) bulk_wait0();
    if (!reported) halo_report(P);
}

// ---------------------------------------------------------------------------------------------------
// sweep B (v4): one thread per node, no separate face tasks.  A node that lies on a face evaluates the Rusanov fluxes of
// the U- and theta-equations there itself: its own side from its registers (trace_coef/trace_apply, the very functions
// sweep A used for the trace it published), the other side from the neighbour's trace block that a bulk-async copy
// brought into shared memory.  The contravariant fluxes replace the gradients in place (every thread reads only its own
// node's gradients), so one CTA barrier separates them from the tensor-product divergence; results leave from registers.
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, bool VISC, bool TRI>
struct CfgB {
    using Dm = Dims<NX, NY, NZ>;
    static constexpr int NP = Dm::NP, NPS = Dm::NPS, NPF = Dm::NPF;
    static constexpr int NT = pad_to(NP, 32);
    static constexpr int NISS = (NT / 32) < 4 ? (NT / 32) : 4;         // issuing threads: lane 31 of the first warps
    static constexpr int TBS = trace_bs(NPF);
    // staged arrays: rho_old, rho_new, U(3), T, p, S, rho_ref, [GU(9), GT(3)], [Jinv(9), cV]
    static constexpr int B_RO = 0, B_RN = 1, B_U = 2, B_T = 5, B_P = 6, B_S = 7, B_RR = 8, B_GU = 9, B_GT = 18;
    static constexpr int B_J = VISC ? 21 : 9, B_CV = B_J + 9;
    static constexpr int NIN = B_J + (TRI ? 0 : 10);
    static constexpr int oIn = 0;
    static constexpr int oT = oIn + 2 * NIN * NPS;                      // [6][TBS] neighbour traces (single buffer)
    static constexpr int oRec = oT + 6 * TBS;
    static constexpr int oH = oRec + 3 * RECD;                          // [12][NPS] (inviscid runs only; else in place over GU/GT)
    static constexpr int oD = oH + (VISC ? 0 : 12 * NPS);
    static constexpr int oBar = oD + 3 * MAXN * MAXN;
    static constexpr size_t smem = sizeof(double) * (size_t)(oBar + 8);
    static constexpr int minb(int regs) {
        int bs = (int)((227 * 1024) / (smem + 1024));
        int br = 65536 / (NT * regs);
        int b = bs < br ? bs : br;
        return b < 1 ? 1 : (b > 8 ? 8 : b);
    }
};

template <int NX, int NY, int NZ, bool VISC, bool TRI, int MINB, bool MORTAR = false>
__global__ void __launch_bounds__((CfgB<NX, NY, NZ, VISC, TRI>::NT), MINB) sweepB_v4(const __grid_constant__ KParams P) {
    using C = CfgB<NX, NY, NZ, VISC, TRI>;
    constexpr int NP = C::NP, NPS = C::NPS, NT = C::NT, NIN = C::NIN, TBS = C::TBS, NPF = C::NPF, NISS = C::NISS;
    constexpr int A_RO = C::B_RO, A_RN = C::B_RN, A_U = C::B_U, A_T = C::B_T, A_P = C::B_P, A_GU = C::B_GU, A_GT = C::B_GT, A_J = C::B_J,
                  A_CV = C::B_CV;
    constexpr int TCS = trace_cs(NPF);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const sm = reinterpret_cast<double*>(smem_raw);
    double* const sIn = sm + C::oIn;         // [2][NIN][NPS]
    double* const sT = sm + C::oT;           // [6][TBS]
    double* const sRec = sm + C::oRec;       // [3][RECD]
    double* const sD = sm + C::oD;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sm + C::oBar);   // full[2], rec[3], traces

    const int tid = threadIdx.x;
    const uint32_t stride = gridDim.x;
    uint32_t seq = blockIdx.x;
    if (seq >= P.nB) {
        if (P.halo) halo_report(P);
        return;
    }
    // issuing threads: lane 31 of the first NISS warps, each owning every NISS-th bulk copy
    const int iss = ((tid & 31) == 31 && (tid >> 5) < NISS) ? (tid >> 5) : -1;

    auto elem_of = [&](uint32_t q) -> uint32_t { return P.sched ? P.sched[q] : q; };
    auto issue_rec = [&](int slot, uint32_t elem) {
        mbar_expect_tx(&bars[2 + slot], (uint32_t)sizeof(ElemRec));
        bulk_g2s(sRec + slot * RECD, P.elemRec + elem, (uint32_t)sizeof(ElemRec), &bars[2 + slot]);
    };
    // this issuer's share of the arrays of `elem`
    auto issue_arrays = [&](int st, uint32_t elem) {
        uint64_t* bar = &bars[st];
        double* dst = sIn + (size_t)st * NIN * NPS;
        const size_t off = (size_t)elem * NPS;
        constexpr uint32_t B = NPS * sizeof(double);
        const int mine = (NIN - iss + NISS - 1) / NISS;          // copies q = iss, iss + NISS, ... < NIN
        mbar_expect_tx(bar, (uint32_t)(mine * B));
        for (int q = iss; q < NIN; q += NISS) bulk_g2s(dst + q * NPS, P.srcB[q] + off, B, bar);
    };
    // the six neighbour traces named by the record in `slot` (which has landed)
    auto issue_traces = [&](int slot) {
        mbar_expect_tx(&bars[5], (uint32_t)(6 * TBS * sizeof(double)));
        const FaceRec* fr = reinterpret_cast<const FaceRec*>(sRec + slot * RECD);
#pragma unroll
        for (int f = 0; f < 6; f++)
            bulk_g2s(sT + f * TBS, P.traceA + (size_t)fr[f].otherBlock * TBS, (uint32_t)(TBS * sizeof(double)), &bars[5]);
    };

    if (tid == 0) {
        mbar_init(&bars[0], NISS); mbar_init(&bars[1], NISS);
#pragma unroll
        for (int q = 2; q < 6; q++) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (iss == 0) {
        issue_rec(0, elem_of(seq));
        if (seq + stride < P.nB) issue_rec(1, elem_of(seq + stride));
    }
    if (iss >= 0) issue_arrays(0, elem_of(seq));
    for (int q = tid; q < 3 * MAXN * MAXN; q += NT) sD[q] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];
    mbar_wait(&bars[2], 0);
    if (iss == NISS - 1) issue_traces(0);
    __syncthreads();                           // sD

    int st = 0, rs = 0;
    uint32_t it = 0;
    bool reported = (P.halo == nullptr);
    for (;;) {
        if (!reported && seq >= P.nHalo) { halo_report(P); reported = true; }
        const uint32_t elem = elem_of(seq);
        const uint32_t nxt = seq + stride, nxt2 = nxt + stride;
        const bool hasNext = nxt < P.nB;
        const int rs1 = (rs == 2) ? 0 : rs + 1, rs2 = (rs1 == 2) ? 0 : rs1 + 1;
        if (iss >= 0) {
            if (hasNext) issue_arrays(st ^ 1, elem_of(nxt));
            if (iss == 0 && nxt2 < P.nB) issue_rec(rs2, elem_of(nxt2));
        }
        if (it > 0) mbar_wait(&bars[2 + rs], (it / 3) & 1u);       // this element's record (requested two elements ago)
        mbar_wait(&bars[st], (it >> 1) & 1u);
        mbar_wait(&bars[5], it & 1u);
        double* const in = sIn + (size_t)st * NIN * NPS;
        double* const sH = VISC ? in + A_GU * NPS : sm + C::oH;    // [12][NPS] contravariant fluxes (in place over the gradients)
        const double* const rec = sRec + rs * RECD;

        const int nt = fresh_tid();
        const bool nodeT = nt < NP;
        const int i = nt / (NY * NZ), j = (nt / NZ) % NY, k = nt % NZ;
        const size_t idx = (size_t)elem * NPS + nt;
        double rho_o = 0, rho_nw = 1, u[3] = {0, 0, 0}, th = 0, cV = 1, rf[4] = {0, 0, 0, 0};
        if (nodeT) {
            SideState me;
            me.rho_o = rho_o = in[A_RO * NPS + nt]; me.rho_n = rho_nw = in[A_RN * NPS + nt];
            me.u[0] = u[0] = in[A_U * NPS + nt]; me.u[1] = u[1] = in[(A_U + 1) * NPS + nt]; me.u[2] = u[2] = in[(A_U + 2) * NPS + nt];
            me.th = th = in[A_T * NPS + nt] + P.T0;
            me.pp = in[A_P * NPS + nt];
            if (VISC) {
#pragma unroll
                for (int c = 0; c < 9; c++) me.gU[c] = in[(A_GU + c) * NPS + nt];
#pragma unroll
                for (int c = 0; c < 3; c++) me.gT[c] = in[(A_GT + c) * NPS + nt];
            }
            // ---- faces this node lies on, local-face-id order: k-faces 0/1, j-faces 2/3, i-faces 4/5 ----
            const bool onK = (k == 0 || k == NZ - 1), onJ = (j == 0 || j == NY - 1), onI = (i == 0 || i == NX - 1);
            if (onK || onJ || onI) {
                TraceCoef K;
                trace_coef(me, in[C::B_S * NPS + nt], P.nu, P.iPr, VISC, K);
                const double wi = P.W[0][i], wj = P.W[1][j], wk = P.W[2][k];
#pragma unroll
                for (int ax = 0; ax < 3; ax++) {
                    const bool on = (ax == 0) ? onK : (ax == 1 ? onJ : onI);
                    if (!on) continue;
                    const int cx = (ax == 0) ? k : (ax == 1 ? j : i);
                    const int s = 2 * ax + (cx != 0 ? 1 : 0);
                    const int fa = (ax == 2) ? j : i, fb = (ax == 0) ? j : k;                    // face coordinates (a,b)
                    const double fw = (ax == 0) ? wi * wj / 4 : (ax == 1 ? wi * wk / 4 : wj * wk / 4);      // face_weight()
                    const FaceRec* fr = reinterpret_cast<const FaceRec*>(rec) + s;
                    const uint32_t meta = fr->meta;
                    const uint32_t fid = meta & FM_FID_MASK;
                    if (MORTAR && (meta & FM_MORTAR)) {
                        // non-conforming face: mortarB_kernel left this node's momentum and theta fluxes
                        const int mslot = (ax == 0) ? fa * NY + fb : fa * NZ + fb;
                        const double* mc = P.mortarB + (size_t)fr->other * (MORTAR_NB * MORTAR_MAXF) + mslot;
#pragma unroll
                        for (int c = 0; c < 4; c++) rf[c] += mc[c * MORTAR_MAXF];
                        continue;
                    }
                    const bool own = meta & FM_OWNER;
                    const double al = (meta & FM_HALF) ? 0.5 : 0.0;
                    const double N[3] = {fr->vec[0] * fw, fr->vec[1] * fw, fr->vec[2] * fw};
                    const double fu[3] = {fr->unit[0], fr->unit[1], fr->unit[2]};
                    const double nN = fu[0] * N[0] + fu[1] * N[1] + fu[2] * N[2];
                    double mt[7];
                    trace_apply(K, N, mt);
                    // the other side: its face trace, slot = the same (a,b) in ITS face numbering (ghost blocks use mine)
                    const int myslot = (ax == 0) ? fa * NY + fb : fa * NZ + fb;
                    const int oslot = (fid == FM_GHOST) ? myslot : ((fid < 2) ? fa * NY + fb : fa * NZ + fb);
                    const double* xt = sT + s * TBS + oslot;
                    const double wo = own ? al : 1 - al, wx = own ? 1 - al : al;
                    const double lam = (mt[6] * wo + xt[6 * TCS] * wx) / 2;
                    const double sg = own ? 1.0 : -1.0;
                    const double dqN = xt[4 * TCS] - mt[4];
                    const double dqT = xt[5 * TCS] - mt[5];
#pragma unroll
                    for (int c = 0; c < 3; c++) rf[c] += sg * ((mt[c] * wo + xt[c * TCS] * wx) - fu[c] * (lam * (sg * dqN)));
                    rf[3] += sg * ((mt[3] * wo + xt[3 * TCS] * wx) - lam * (sg * dqT) * nN);
                }
            }
            // ---- contravariant fluxes of the four equations ----
            const double mu = VISC ? rho_o * P.nu : 0.0;
            double Jin[9];
            if (TRI) {
                const double wcv = ((P.W[0][i] * P.W[1][j]) * P.W[2][k]) / 8;
                elem_metrics(rec, P.X[0][i], P.X[1][j], P.X[2][k], wcv, Jin, cV);
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
                    fq[b] = Fc[a] * u[b] + (a == b ? me.pp : 0.0);
                    if (VISC) fq[b] -= mu * me.gU[a * 3 + b];
                }
#pragma unroll
                for (int d = 0; d < 3; d++) H[a * 3 + d] = fq[0] * Jin[d] + fq[1] * Jin[3 + d] + fq[2] * Jin[6 + d];
            }
            {
                double fq[3];
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    fq[b] = Fc[b] * th;
                    if (VISC) fq[b] -= (mu * P.iPr) * me.gT[b];
                }
#pragma unroll
                for (int d = 0; d < 3; d++) H[9 + d] = fq[0] * Jin[d] + fq[1] * Jin[3 + d] + fq[2] * Jin[6 + d];
            }
#pragma unroll
            for (int c = 0; c < 12; c++) sH[c * NPS + nt] = H[c];
        }
        __syncthreads();                                                                   // (1) fluxes stored, traces consumed
        if (hasNext && iss == NISS - 1) {
            // the next element's record (requested a whole element ago) names the trace blocks to fetch
            mbar_wait(&bars[2 + rs1], ((it + 1) / 3) & 1u);
            issue_traces(rs1);
        }
        if (nodeT) {
            const double rref = P.buoyancy ? in[C::B_RR * NPS + nt] : 0.0;
            double r[4];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                double acc = 0;
#pragma unroll
                for (int ii = 0; ii < NX; ii++) acc += sH[(a * 3 + 0) * NPS + ii * NY * NZ + j * NZ + k] * NSEM_DM(0, ii * NX + i);
#pragma unroll
                for (int jj = 0; jj < NY; jj++) acc += sH[(a * 3 + 1) * NPS + i * NY * NZ + jj * NZ + k] * NSEM_DM(1, jj * NY + j);
#pragma unroll
                for (int kk = 0; kk < NZ; kk++) acc += sH[(a * 3 + 2) * NPS + i * NY * NZ + j * NZ + kk] * NSEM_DM(2, kk * NZ + k);
                r[a] = rf[a] - acc;
            }
            const double ap0 = P.mrdt * cV;
            const double ap = ap0 * rho_nw;
            double g[3] = {P.g[0], P.g[1], P.g[2]};
            if (P.has_gfield) { g[0] = P.gfield[0][idx]; g[1] = P.gfield[1][idx]; g[2] = P.gfield[2][idx]; }
            const double drho = P.buoyancy ? (rho_nw - rref) : 0.0;
            const double rap = 1.0 / ap;                 // x = Su / ap (solve.cpp:563-570) as one reciprocal and 4 products
            double un[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double Su = (r[c] - (drho * g[c]) * cV) + (u[c] * rho_o) * ap0;
                un[c] = (P.op_mode & 1) ? r[c] : Su * rap;
                P.U_new[c][idx] = un[c];
            }
            {
                const double Su = r[3] + (th * rho_o) * ap0;
                const double Tn = (P.op_mode & 1) ? r[3] : Su * rap - P.T0;
                P.T_new[idx] = Tn;
                // |U| + c of the new state, from the values as stored (every other producer of S reads them back from memory)
                const double Sn = side_speed(un, Tn + P.T0, P.gamma * P.R);
                P.S_new[idx] = Sn;
                // ---- partition boundary: U, T, S of this node into the neighbour's receive window ----
                if (P.halo) {
                    const bool onK = (k == 0 || k == NZ - 1), onJ = (j == 0 || j == NY - 1), onI = (i == 0 || i == NX - 1);
#pragma unroll
                    for (int ax = 0; ax < 3; ax++) {
                        const bool on = (ax == 0) ? onK : (ax == 1 ? onJ : onI);
                        if (!on) continue;
                        const int cx = (ax == 0) ? k : (ax == 1 ? j : i);
                        const FaceRec* fr = reinterpret_cast<const FaceRec*>(rec) + (2 * ax + (cx != 0 ? 1 : 0));
                        if ((fr->meta & FM_FID_MASK) != FM_GHOST || (MORTAR && (fr->meta & FM_MORTAR))) continue;
                        const uint32_t g = (uint32_t)((fr->other - P.ghostBase) / Dims<NX, NY, NZ>::GPS);
                        const uint32_t peer = P.haloGhost[2 * g];
                        if (peer == 0xffffffffu) continue;
                        const int slot = (ax == 0) ? i * NY + j : (ax == 1 ? i * NZ + k : j * NZ + k);
                        double* w = P.halo->win[peer] + P.haloGhost[2 * g + 1] + slot;
                        const uint64_t hs = P.halo->stride[peer];
                        w[0] = un[0]; w[hs] = un[1]; w[2 * hs] = un[2]; w[3 * hs] = Tn; w[4 * hs] = Sn;
                    }
                }
            }
        }
        fence_async_smem();                    // the in-place fluxes (generic writes) precede the next bulk copies into this stage
        __syncthreads();                                                                   // (2) stage free
        if (!hasNext) break;
        seq = nxt;
        st ^= 1;
        rs = rs1;
        it++;
    }
    if (!reported) halo_report(P);
}

}  // namespace v4
}  // namespace nsem
