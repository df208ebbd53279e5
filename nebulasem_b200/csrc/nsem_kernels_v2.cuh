// nsem_kernels_v2.cuh -- sm_100a sweeps with bulk-async (TMA 1-D, cp.async.bulk) staging and dense face tasks.
//
// Same mathematics and data flow as nsem_kernels.cuh (v1); what changes is how the SM is fed:
//   * every per-node input array of the element (state, gradients, metrics, reference state) is brought into
//     shared memory with one cp.async.bulk (UBLKCP) per array, completion tracked by an mbarrier, so the HBM
//     stream is decoupled from the arithmetic and several resident CTAs overlap load and compute;
//   * while the copies are in flight the threads decode their face task, read the face tables and issue the
//     neighbour-trace gathers (which mostly hit L2);
//   * face work is remapped from "each node walks the faces it lies on" (4.5 divergent face bodies per warp
//     in v1) to one dense task per face node: 2(NX NY + NX NZ + NY NZ) tasks per element, results go through
//     shared memory and are summed by the node threads in local-face-id order (same order as the reference's
//     per-cell loop, field.h:3093-3114), so results stay deterministic and atomics-free.
#pragma once
#include "nsem_kernels.cuh"

namespace nsem {
namespace v2 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

template <int NX, int NY, int NZ>
struct Tasks {
    static constexpr int n0 = NX * NY, n2 = NX * NZ, n4 = NY * NZ;
    static constexpr int o1 = n0, o2 = 2 * n0, o3 = 2 * n0 + n2, o4 = 2 * n0 + 2 * n2, o5 = o4 + n4;
    static constexpr int NFT = 2 * (n0 + n2 + n4);
    __device__ static __forceinline__ void decode(int f, int& s, int& a, int& b) {
        if (f < o2) { s = (f >= o1); const int l = f - s * n0; a = l / NY; b = l % NY; }
        else if (f < o4) { s = 2 + (f >= o3); const int l = f - o2 - (s - 2) * n2; a = l / NZ; b = l % NZ; }
        else { s = 4 + (f >= o5); const int l = f - o4 - (s - 4) * n4; a = l / NZ; b = l % NZ; }
    }
    __device__ static __forceinline__ int index(int s, int a, int b) {
        if (s < 2) return s * n0 + a * NY + b;
        if (s < 4) return o2 + (s - 2) * n2 + a * NZ + b;
        return o4 + (s - 4) * n4 + a * NZ + b;
    }
};

template <int NX, int NY, int NZ, int EPB>
struct Cfg {
    using Dm = Dims<NX, NY, NZ>;
    using Tk = Tasks<NX, NY, NZ>;
    static constexpr int NP = Dm::NP, NPS = Dm::NPS, NFT = Tk::NFT;
    static constexpr int WORK = EPB * (NP > NFT ? NP : NFT);
    static constexpr int NT = pad_to(WORK, 32);
    // arrays staged by sweep A: rho, U(3), T, Jinv(9), cV, p_ref
    static constexpr int NIN_A = 16;
    // arrays staged by sweep B: rho_old, rho_new, U(3), T, p, [GU(9), GT(3)], Jinv(9), cV, rho_ref
    __host__ __device__ static constexpr int nin_b(bool visc) { return visc ? 30 : 18; }
    static constexpr size_t smemA(bool visc) {
        return sizeof(double) * ((size_t)NIN_A * EPB * NPS + 3 * EPB * NP + (size_t)EPB * NFT * (visc ? 8 : 1) + 3 * MAXN * MAXN) + 16;
    }
    // resident CTAs per SM aimed at: limited by 227 KB of shared memory and by >= 96 registers per thread
    static constexpr int minb(size_t smem, int regs = 96) {
        int bs = (int)((227 * 1024) / (smem + 1024));
        int br = 65536 / (NT * regs);
        int b = bs < br ? bs : br;
        return b < 1 ? 1 : (b > 4 ? 4 : b);
    }
    static constexpr int FS = trace_cs(Dm::NPF), TBS = trace_bs(Dm::NPF);
    static constexpr size_t smemB(bool visc) {
        return sizeof(double) * ((size_t)nin_b(visc) * EPB * NPS + 12 * EPB * NP + (size_t)EPB * NFT * 4 + (size_t)EPB * 6 * TBS +
                                 3 * MAXN * MAXN) + 16;
    }
};

// ---------------------------------------------------------------------------------------------------
// sweep A (v2)
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, int EPB, bool VISC, int MINB>
__global__ void __launch_bounds__(Cfg<NX, NY, NZ, EPB>::NT, MINB) sweepA_v2(const __grid_constant__ KParams P) {
    using C = Cfg<NX, NY, NZ, EPB>;
    using Tk = Tasks<NX, NY, NZ>;
    constexpr int NP = C::NP, NPS = C::NPS, NFT = C::NFT, NT = C::NT, NIN = C::NIN_A;
    constexpr int FS = VISC ? 8 : 1;   // doubles per face task: signed mass flux, dq[4], signed N[3]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sIn = reinterpret_cast<double*>(smem_raw);             // [NIN][EPB][NPS]
    double* sR = sIn + (size_t)NIN * EPB * NPS;                    // [3][EPB][NP]
    double* sF = sR + 3 * EPB * NP;                                // [FS][EPB][NFT]
    double* sD = sF + (size_t)EPB * NFT * FS;                      // [3][MAXN*MAXN]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sD + 3 * MAXN * MAXN);
    auto IN = [&](int arr, int e, int t) -> double& { return sIn[((size_t)arr * EPB + e) * NPS + t]; };

    const int tid = threadIdx.x;
    const uint32_t first = blockIdx.x * EPB;
    const int nvalid = (int)min((uint32_t)EPB, P.nB - first);
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)(nvalid * NIN * NPS * sizeof(double)));
        __syncwarp();
        for (int q = tid; q < nvalid * NIN; q += 32) {
            const int e = q / NIN, arr = q % NIN;
            const uint32_t elem = P.sched ? P.sched[first + e] : first + e;
            const double* src = (arr == 0) ? P.rho_old : (arr < 4) ? P.U_old[arr - 1] : (arr == 4) ? P.T_old
                              : (arr < 14) ? P.Jinv[arr - 5] : (arr == 14) ? P.cV : P.p_ref;
            bulk_g2s(&IN(arr, e, 0), src + (size_t)elem * NPS, NPS * sizeof(double), bar);
        }
    }
    for (int q = tid; q < 3 * MAXN * MAXN; q += NT) sD[q] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];

    // ---- node item ----
    const int ne = tid / NP, nt = tid % NP;
    const bool nodeOn = tid < EPB * NP && ne < nvalid;
    const uint32_t nelem = nodeOn ? (P.sched ? P.sched[first + ne] : first + ne) : 0u;
    const int i = nt / (NY * NZ), j = (nt / NZ) % NY, k = nt % NZ;
    const size_t idx = (size_t)nelem * NPS + nt;

    // ---- face task: decode, one 64-byte table record and the neighbour gathers are ISSUED before the element
    //      data has landed; nothing loaded here is consumed until after the wait ----
    const int fe = tid / NFT, ff = tid % NFT;
    const bool faceOn = tid < EPB * NFT && fe < nvalid;
    int fs = 0, fa = 0, fb = 0;
    double2 q0 = {0, 0}, q1 = {0, 0}, q2 = {0, 0}, q3 = {0, 0};
    double xr = 0, xu0 = 0, xu1 = 0, xu2 = 0, xT = 0;
    if (faceOn && P.probe != 2) {
        Tk::decode(ff, fs, fa, fb);
        const uint32_t felem = P.sched ? P.sched[first + fe] : first + fe;
        const double2* rp = reinterpret_cast<const double2*>(P.faceRec + ((size_t)felem * 6 + fs));
        q0 = rp[0]; q1 = rp[1]; q2 = rp[2]; q3 = rp[3];
        const unsigned long long om = (unsigned long long)__double_as_longlong(q0.x);
        const uint32_t fid = (uint32_t)(om >> 32) & FM_FID_MASK;
        const int n = (fs < 2) ? fa * NY + fb : fa * NZ + fb;
        const size_t oidx = (size_t)(uint32_t)om + (fid == FM_GHOST ? n : face_node<NX, NY, NZ>(fid, fa, fb));
        xr = P.rho_old[oidx];
        xu0 = P.U_old[0][oidx]; xu1 = P.U_old[1][oidx]; xu2 = P.U_old[2][oidx];
        xT = P.T_old[oidx];
    }

    mbar_wait(bar, 0);
    __syncthreads();            // sD visible
    if (P.probe) {
        // bandwidth probe (not a product path): touch every staged and gathered value, write the 14 outputs (+ traces if probe == 3)
        double acc = xr + xu0 + xu1 + xu2 + xT;
        if (tid < EPB * NP && tid / NP < nvalid) {
            for (int a = 0; a < NIN; a++) acc += IN(a, tid / NP, tid % NP);
            const size_t id2 = (size_t)(P.sched ? P.sched[first + tid / NP] : first + tid / NP) * NPS + tid % NP;
            P.rho_new[id2] = acc; P.p[id2] = acc;
            if (VISC) { for (int c = 0; c < 9; c++) P.GU[c][id2] = acc; for (int c = 0; c < 3; c++) P.GT[c][id2] = acc; }
        }
        if (P.probe == 3 && faceOn) {
            const uint32_t felem = P.sched ? P.sched[first + fe] : first + fe;
            double* dst = P.traceA + ((size_t)felem * 6 + fs) * C::TBS + ((fs < 2) ? fa * NY + fb : fa * NZ + fb);
            for (int c = 0; c < 7; c++) dst[c * C::FS] = acc;
        }
        return;
    }

    // ---- node: contravariant mass flux, theta ----
    double rho = 0, u0 = 0, u1 = 0, u2 = 0, th = 0, cV = 1, Jin[9];
    if (nodeOn) {
        rho = IN(0, ne, nt); u0 = IN(1, ne, nt); u1 = IN(2, ne, nt); u2 = IN(3, ne, nt);
        th = IN(4, ne, nt) + P.T0;
        cV = IN(14, ne, nt);
#pragma unroll
        for (int c = 0; c < 9; c++) Jin[c] = IN(5 + c, ne, nt) * cV;
        const double F0 = u0 * rho, F1 = u1 * rho, F2 = u2 * rho;
#pragma unroll
        for (int d = 0; d < 3; d++) sR[(d * EPB + ne) * NP + nt] = F0 * Jin[d] + F1 * Jin[3 + d] + F2 * Jin[6 + d];
        IN(4, ne, nt) = th;     // the T slot holds theta from here on (gradient + face tasks)
    }
    __syncthreads();

    double r_rho = 0, gU[9], gT[3];
    if (nodeOn) {
        double acc = 0;
#pragma unroll
        for (int ii = 0; ii < NX; ii++) acc += sR[(0 * EPB + ne) * NP + ii * NY * NZ + j * NZ + k] * sD[0 * MAXN * MAXN + ii * NX + i];
#pragma unroll
        for (int jj = 0; jj < NY; jj++) acc += sR[(1 * EPB + ne) * NP + i * NY * NZ + jj * NZ + k] * sD[1 * MAXN * MAXN + jj * NY + j];
#pragma unroll
        for (int kk = 0; kk < NZ; kk++) acc += sR[(2 * EPB + ne) * NP + i * NY * NZ + j * NZ + kk] * sD[2 * MAXN * MAXN + kk * NZ + k];
        r_rho = -acc;
        if (VISC) {
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const double* q = &IN(1 + f, ne, 0);
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

    // ---- face tasks ----
    if (faceOn) {
        const uint32_t meta = (uint32_t)((unsigned long long)__double_as_longlong(q0.x) >> 32);
        const bool own = meta & FM_OWNER;
        const double al = (meta & FM_HALF) ? 0.5 : 0.0;
        const double w = face_weight<NX, NY, NZ>(P, fs, fa, fb);
        const double N0 = q0.y * w, N1 = q1.x * w, N2 = q1.y * w;          // fN[k] = gFN * w_a w_b / 4
        const double nN = q2.x * N0 + q2.y * N1 + q3.x * N2;               // unit(fN).fN
        const double xth = xT + P.T0;
        const int fln = face_node<NX, NY, NZ>(fs, fa, fb);
        const double mr = IN(0, fe, fln), m0 = IN(1, fe, fln), m1 = IN(2, fe, fln), m2 = IN(3, fe, fln), mth = IN(4, fe, fln);
        const double rho_o = own ? mr : xr, rho_n = own ? xr : mr;
        const double uo0 = own ? m0 : xu0, uo1 = own ? m1 : xu1, uo2 = own ? m2 : xu2;
        const double un0 = own ? xu0 : m0, un1 = own ? xu1 : m1, un2 = own ? xu2 : m2;
        const double th_o = own ? mth : xth, th_n = own ? xth : mth;
        const double mo = sqrt(uo0 * uo0 + (uo1 * uo1 + uo2 * uo2)), mn = sqrt(un0 * un0 + (un1 * un1 + un2 * un2));
        const double co = sqrt(P.gamma * P.R * th_o), cn = sqrt(P.gamma * P.R * th_n);
        const double lam = ((mo * al + mn * (1 - al)) + (co * al + cn * (1 - al))) / 2;
        const double fo = rho_o * (uo0 * N0 + uo1 * N1 + uo2 * N2), fn = rho_n * (un0 * N0 + un1 * N1 + un2 * N2);
        const double flux = (fo * al + fn * (1 - al)) - lam * (rho_n - rho_o) * nN;
        double* out = &sF[(size_t)fe * NFT + ff];
        constexpr int CS = EPB * NFT;          // component stride
        out[0] = own ? flux : -flux;
        if (VISC) {
            const double sgn = own ? 1.0 : -1.0;
            out[1 * CS] = (uo0 * al + un0 * (1 - al)) - m0;
            out[2 * CS] = (uo1 * al + un1 * (1 - al)) - m1;
            out[3 * CS] = (uo2 * al + un2 * (1 - al)) - m2;
            out[4 * CS] = (th_o * al + th_n * (1 - al)) - mth;
            out[5 * CS] = sgn * N0; out[6 * CS] = sgn * N1; out[7 * CS] = sgn * N2;
        }
    }
    __syncthreads();

    if (nodeOn) {
#pragma unroll
        for (int s = 0; s < 6; s++) {
            if (!on_face(s, i, j, k, NX, NY, NZ)) continue;
            int a, b;
            face_slot<NX, NY, NZ>(s, i, j, k, a, b);
            const double* in = &sF[(size_t)ne * NFT + Tk::index(s, a, b)];
            constexpr int CS = EPB * NFT;
            r_rho += in[0];
            if (VISC) {
                const double q0_ = in[1 * CS], q1_ = in[2 * CS], q2_ = in[3 * CS], q3_ = in[4 * CS];
#pragma unroll
                for (int aa = 0; aa < 3; aa++) {
                    const double sn = in[(5 + aa) * CS];
                    gU[aa * 3 + 0] += sn * q0_;
                    gU[aa * 3 + 1] += sn * q1_;
                    gU[aa * 3 + 2] += sn * q2_;
                    gT[aa] += sn * q3_;
                }
            }
        }
        const double ap0 = (-1.0 / P.dt) * cV;
        const double rho_new = (r_rho + rho * ap0) / ap0;
        P.rho_new[idx] = rho_new;
        const double ppn = __dsub_rn(eos_pressure(P.P0, P.R, P.gamma, rho_new, th), IN(15, ne, nt));
        P.p[idx] = ppn;
        // the new values also go to shared memory (over the metric slots, which only this thread read) for the trace pass
        IN(14, ne, nt) = rho_new;
        IN(15, ne, nt) = ppn;
        if (VISC) {
            const double rcV = 1.0 / cV;            // r / cV (field.h:3359) as one reciprocal and 12 products
#pragma unroll
            for (int c = 0; c < 9; c++) { const double v = gU[c] * rcV; P.GU[c][idx] = v; IN(5 + c, ne, nt) = v; }
#pragma unroll
            for (int c = 0; c < 3; c++) { const double v = gT[c] * rcV; P.GT[c][idx] = v; sR[(c * EPB + ne) * NP + nt] = v; }
        }
    }
    __syncthreads();
    // this side's face traces for sweep B of the neighbours (and of the peers behind a partition boundary): dense face
    // tasks again, so every face block is written as contiguous runs
    if (faceOn) {
        const double w = face_weight<NX, NY, NZ>(P, fs, fa, fb);
        const double Nv[3] = {q0.y * w, q1.x * w, q1.y * w};
        const int fln = face_node<NX, NY, NZ>(fs, fa, fb);
        SideState q;
        q.rho_o = IN(0, fe, fln); q.rho_n = IN(14, fe, fln);
        q.u[0] = IN(1, fe, fln); q.u[1] = IN(2, fe, fln); q.u[2] = IN(3, fe, fln);
        q.th = IN(4, fe, fln);
        q.pp = IN(15, fe, fln);
        if (VISC) {
#pragma unroll
            for (int c = 0; c < 9; c++) q.gU[c] = IN(5 + c, fe, fln);
#pragma unroll
            for (int c = 0; c < 3; c++) q.gT[c] = sR[(c * EPB + fe) * NP + fln];
        }
        double out[7];
        side_trace(q, Nv, P.nu, P.iPr, P.gamma * P.R, VISC, out);
        constexpr int TS = C::FS;
        const uint32_t felem = P.sched ? P.sched[first + fe] : first + fe;
        const int n = (fs < 2) ? fa * NY + fb : fa * NZ + fb;
        double* dst = P.traceA + ((size_t)felem * 6 + fs) * C::TBS + n;
#pragma unroll
        for (int c = 0; c < 7; c++) dst[c * TS] = out[c];
    }
}

// ---------------------------------------------------------------------------------------------------
// sweep B (v2)
// ---------------------------------------------------------------------------------------------------
template <int NX, int NY, int NZ, int EPB, bool VISC, int MINB>
__global__ void __launch_bounds__(Cfg<NX, NY, NZ, EPB>::NT, MINB) sweepB_v2(const __grid_constant__ KParams P) {
    using C = Cfg<NX, NY, NZ, EPB>;
    using Tk = Tasks<NX, NY, NZ>;
    constexpr int NP = C::NP, NPS = C::NPS, NFT = C::NFT, NT = C::NT;
    constexpr int NIN = C::nin_b(VISC);
    // staged array slots
    constexpr int A_RO = 0, A_RN = 1, A_U = 2, A_T = 5, A_P = 6, A_GU = 7, A_GT = 16;
    constexpr int A_J = VISC ? 19 : 7, A_CV = A_J + 9, A_RR = A_CV + 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // bulk-copy destinations first: sIn and every trace block are multiples of 128 bytes
    double* sIn = reinterpret_cast<double*>(smem_raw);             // [NIN][EPB][NPS]
    double* sT = sIn + (size_t)NIN * EPB * NPS;                    // [EPB][6][7][FS] neighbour-side face traces
    double* sH = sT + (size_t)EPB * 6 * C::TBS;                    // [12][EPB][NP]
    double* sF = sH + 12 * EPB * NP;                               // [4][EPB][NFT]
    double* sD = sF + (size_t)EPB * NFT * 4;                       // [3][MAXN*MAXN]
    uint64_t* bar = reinterpret_cast<uint64_t*>(sD + 3 * MAXN * MAXN);
    auto IN = [&](int arr, int e, int t) -> double& { return sIn[((size_t)arr * EPB + e) * NPS + t]; };

    const int tid = threadIdx.x;
    const uint32_t first = blockIdx.x * EPB;
    const int nvalid = (int)min((uint32_t)EPB, P.nB - first);
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    constexpr int FS = C::FS, TBS = C::TBS;
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)(nvalid * (NIN * NPS + 6 * TBS) * sizeof(double)));
        __syncwarp();
        // neighbour traces: one bulk copy per face (the block id sits in the face record)
        for (int q = tid; q < nvalid * 6; q += 32) {
            const int e = q / 6, f = q % 6;
            const uint32_t elem = P.sched ? P.sched[first + e] : first + e;
            const uint64_t blk = P.faceRec[(size_t)elem * 6 + f].otherBlock;
            bulk_g2s(sT + ((size_t)e * 6 + f) * TBS, P.traceA + (size_t)blk * TBS, TBS * sizeof(double), bar);
        }
        for (int q = tid; q < nvalid * NIN; q += 32) {
            const int e = q / NIN, arr = q % NIN;
            const uint32_t elem = P.sched ? P.sched[first + e] : first + e;
            const double* src;
            if (arr == A_RO) src = P.rho_old;
            else if (arr == A_RN) src = P.rho_new;
            else if (arr < A_T) src = P.U_old[arr - A_U];
            else if (arr == A_T) src = P.T_old;
            else if (arr == A_P) src = P.p;
            else if (VISC && arr < A_GT) src = P.GU[arr - A_GU];
            else if (VISC && arr < A_J) src = P.GT[arr - A_GT];
            else if (arr < A_CV) src = P.Jinv[arr - A_J];
            else if (arr == A_CV) src = P.cV;
            else src = P.rho_ref;
            bulk_g2s(&IN(arr, e, 0), src + (size_t)elem * NPS, NPS * sizeof(double), bar);
        }
    }
    for (int q = tid; q < 3 * MAXN * MAXN; q += NT) sD[q] = P.D[q / (MAXN * MAXN)][q % (MAXN * MAXN)];

    const int ne = tid / NP, nt = tid % NP;
    const bool nodeOn = tid < EPB * NP && ne < nvalid;
    const uint32_t nelem = nodeOn ? (P.sched ? P.sched[first + ne] : first + ne) : 0u;
    const int i = nt / (NY * NZ), j = (nt / NZ) % NY, k = nt % NZ;
    const size_t idx = (size_t)nelem * NPS + nt;

    const int fe = tid / NFT, ff = tid % NFT;
    const bool faceOn = tid < EPB * NFT && fe < nvalid;
    int fs = 0, fa = 0, fb = 0;
    double2 q0 = {0, 0}, q1 = {0, 0}, q2 = {0, 0}, q3 = {0, 0};
    if (faceOn) {
        Tk::decode(ff, fs, fa, fb);
        const uint32_t felem = P.sched ? P.sched[first + fe] : first + fe;
        const double2* rp = reinterpret_cast<const double2*>(P.faceRec + ((size_t)felem * 6 + fs));
        q0 = rp[0]; q1 = rp[1]; q2 = rp[2]; q3 = rp[3];
    }

    mbar_wait(bar, 0);
    __syncthreads();
    if (P.probe) {
        // bandwidth probe (not a product path): touch every staged value and every gathered value, write the 4 outputs
        double acc = 0;
        if (faceOn) for (int c = 0; c < 7; c++) acc += sT[((size_t)fe * 6 + fs) * TBS + c * FS + ((fs < 2) ? fa * NY + fb : fa * NZ + fb)];
        if (nodeOn) {
            for (int a = 0; a < NIN; a++) acc += IN(a, ne, nt);
            P.U_new[0][idx] = acc; P.U_new[1][idx] = acc; P.U_new[2][idx] = acc; P.T_new[idx] = acc;
        } else if (acc == 1.2345e300) P.T_new[0] = acc;
        return;
    }

    double rho_o = 0, rho_nw = 1, u[3] = {0, 0, 0}, th = 0, cV = 1, rref = 0;
    if (nodeOn) {
        rho_o = IN(A_RO, ne, nt); rho_nw = IN(A_RN, ne, nt);
        u[0] = IN(A_U, ne, nt); u[1] = IN(A_U + 1, ne, nt); u[2] = IN(A_U + 2, ne, nt);
        th = IN(A_T, ne, nt) + P.T0;
        const double pp = IN(A_P, ne, nt);
        cV = IN(A_CV, ne, nt);
        rref = IN(A_RR, ne, nt);
        const double mu = VISC ? rho_o * P.nu : 0.0;
        double Jin[9];
#pragma unroll
        for (int c = 0; c < 9; c++) Jin[c] = IN(A_J + c, ne, nt) * cV;
        const double Fc[3] = {rho_o * u[0], rho_o * u[1], rho_o * u[2]};
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double fq[3];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                fq[b] = Fc[a] * u[b] + (a == b ? pp : 0.0);
                if (VISC) fq[b] -= mu * IN(A_GU + a * 3 + b, ne, nt);
            }
#pragma unroll
            for (int d = 0; d < 3; d++) sH[((a * 3 + d) * EPB + ne) * NP + nt] = fq[0] * Jin[d] + fq[1] * Jin[3 + d] + fq[2] * Jin[6 + d];
        }
        {
            double fq[3];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                fq[b] = Fc[b] * th;
                if (VISC) fq[b] -= (mu * P.iPr) * IN(A_GT + b, ne, nt);
            }
#pragma unroll
            for (int d = 0; d < 3; d++) sH[((9 + d) * EPB + ne) * NP + nt] = fq[0] * Jin[d] + fq[1] * Jin[3 + d] + fq[2] * Jin[6 + d];
        }
    }
    __syncthreads();

    double r[4] = {0, 0, 0, 0};
    if (nodeOn) {
#pragma unroll
        for (int a = 0; a < 4; a++) {
            double acc = 0;
#pragma unroll
            for (int ii = 0; ii < NX; ii++) acc += sH[((a * 3 + 0) * EPB + ne) * NP + ii * NY * NZ + j * NZ + k] * sD[0 * MAXN * MAXN + ii * NX + i];
#pragma unroll
            for (int jj = 0; jj < NY; jj++) acc += sH[((a * 3 + 1) * EPB + ne) * NP + i * NY * NZ + jj * NZ + k] * sD[1 * MAXN * MAXN + jj * NY + j];
#pragma unroll
            for (int kk = 0; kk < NZ; kk++) acc += sH[((a * 3 + 2) * EPB + ne) * NP + i * NY * NZ + j * NZ + kk] * sD[2 * MAXN * MAXN + kk * NZ + k];
            r[a] = -acc;
        }
    }

    if (faceOn) {
        const uint32_t meta = (uint32_t)((unsigned long long)__double_as_longlong(q0.x) >> 32);
        const uint32_t fid = meta & FM_FID_MASK;
        const bool own = meta & FM_OWNER;
        const double al = (meta & FM_HALF) ? 0.5 : 0.0;
        const double w = face_weight<NX, NY, NZ>(P, fs, fa, fb);
        const double N[3] = {q0.y * w, q1.x * w, q1.y * w};
        const double fu[3] = {q2.x, q2.y, q3.x};
        const double nN = fu[0] * N[0] + fu[1] * N[1] + fu[2] * N[2];
        // my side, from the staged element data
        const int fln = face_node<NX, NY, NZ>(fs, fa, fb);
        SideState me;
        me.rho_o = IN(A_RO, fe, fln); me.rho_n = IN(A_RN, fe, fln);
        me.u[0] = IN(A_U, fe, fln); me.u[1] = IN(A_U + 1, fe, fln); me.u[2] = IN(A_U + 2, fe, fln);
        me.th = IN(A_T, fe, fln) + P.T0;
        me.pp = IN(A_P, fe, fln);
        if (VISC) {
#pragma unroll
            for (int c = 0; c < 9; c++) me.gU[c] = IN(A_GU + c, fe, fln);
#pragma unroll
            for (int c = 0; c < 3; c++) me.gT[c] = IN(A_GT + c, fe, fln);
        }
        double mt[7];
        side_trace(me, N, P.nu, P.iPr, P.gamma * P.R, VISC, mt);
        // the other side: its face trace, slot = the same (a,b) in ITS face numbering (ghost blocks use mine)
        const int oslot = (fid == FM_GHOST) ? ((fs < 2) ? fa * NY + fb : fa * NZ + fb) : ((fid < 2) ? fa * NY + fb : fa * NZ + fb);
        const double* xt = sT + ((size_t)fe * 6 + fs) * TBS + oslot;
        const double wo = own ? al : 1 - al, wx = own ? 1 - al : al;
        const double lam = (mt[6] * wo + xt[6 * FS] * wx) / 2;
        const double sg = own ? 1.0 : -1.0;
        const double dqN = xt[4 * FS] - mt[4];
        const double dqT = xt[5 * FS] - mt[5];
        double* out = &sF[(size_t)fe * NFT + ff];
        constexpr int CS = EPB * NFT;
#pragma unroll
        for (int c = 0; c < 3; c++) out[c * CS] = sg * ((mt[c] * wo + xt[c * FS] * wx) - fu[c] * (lam * (sg * dqN)));
        out[3 * CS] = sg * ((mt[3] * wo + xt[3 * FS] * wx) - lam * (sg * dqT) * nN);
    }
    __syncthreads();

    if (!nodeOn) return;
#pragma unroll
    for (int s = 0; s < 6; s++) {
        if (!on_face(s, i, j, k, NX, NY, NZ)) continue;
        int a, b;
        face_slot<NX, NY, NZ>(s, i, j, k, a, b);
        const double* in = &sF[(size_t)ne * NFT + Tk::index(s, a, b)];
#pragma unroll
        for (int c = 0; c < 4; c++) r[c] += in[c * EPB * NFT];
    }
    const double ap0 = (-1.0 / P.dt) * cV;
    const double ap = ap0 * rho_nw;
    double g[3] = {P.g[0], P.g[1], P.g[2]};
    if (P.has_gfield) { g[0] = P.gfield[0][idx]; g[1] = P.gfield[1][idx]; g[2] = P.gfield[2][idx]; }
    const double drho = P.buoyancy ? (rho_nw - rref) : 0.0;
    const double rap = 1.0 / ap;                 // x = Su / ap (solve.cpp:563-570) as one reciprocal and 4 products
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double Su = (r[c] - (drho * g[c]) * cV) + (u[c] * rho_o) * ap0;
        P.U_new[c][idx] = Su * rap;
    }
    {
        const double Su = r[3] + (th * rho_o) * ap0;
        P.T_new[idx] = Su * rap - P.T0;
    }
}

}  // namespace v2
}  // namespace nsem
