// nsem_cuda.cu -- C ABI (include/nsem_c.h) over the sm_100a kernels in nsem_kernels.cuh.
//
// Owns all device memory.  Translates the reference's mesh/field layout (Mesh::initGeomMeshFields,
// src/field/field.cpp:171-280) into the device layout: SoA node arrays with element stride NPS, compact
// ghost cells (NPF slots per boundary face, stride GPS) and per-(element, local face) tables derived
// from gFaceID/gFOC/gFNC.  The derived face-node pairing is verified against the reference's FO/FN maps
// (dg.cpp:328-410) at upload time, so an orientation the kernels do not reproduce is rejected loudly.
#include "../../include/nsem_c.h"
#include "nsem_kernels.cuh"
#include "nsem_kernels_v2.cuh"
#include "nsem_kernels_v4.cuh"
#include "nsem_mortar.cuh"
#include "nsem_halo.cuh"
#include "nsem_amr.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#ifdef NSEM_WITH_NCCL
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>   // types only: the library is resolved at run time (see NcclApi) so that this .so carries no
                    // link-time NCCL dependency and shares whichever libnccl.so.2 the process already loaded
                    // (torch bundles its own; two different NCCL builds in one process do not mix)
#endif

#ifndef NSEM_V4_REGS_A
#define NSEM_V4_REGS_A 128
#endif
#ifndef NSEM_V4_REGS_B
#define NSEM_V4_REGS_B 128
#endif

using namespace nsem;

namespace {
std::string g_create_error;

#ifdef NSEM_WITH_NCCL
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // already in the process (torch)?
        if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define NSEM_SYM(field, name)                                                         \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                     \
    if (!field) { err = std::string("libnccl is missing ") + name; handle = nullptr; return false; }
        NSEM_SYM(GetUniqueId, "ncclGetUniqueId")
        NSEM_SYM(CommInitRank, "ncclCommInitRank")
        NSEM_SYM(CommDestroy, "ncclCommDestroy")
        NSEM_SYM(GroupStart, "ncclGroupStart")
        NSEM_SYM(GroupEnd, "ncclGroupEnd")
        NSEM_SYM(Send, "ncclSend")
        NSEM_SYM(Recv, "ncclRecv")
        NSEM_SYM(AllReduce, "ncclAllReduce")
        NSEM_SYM(GetErrorString, "ncclGetErrorString")
#undef NSEM_SYM
        return true;
    }
};
NcclApi g_nccl;
#endif

#define CUDA_TRY(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                      \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T>& h, cudaStream_t s) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};
}  // namespace

struct nsem_ctx {
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr, comm = nullptr;
    mutable std::string err;
    uint64_t launches = 0;
    bool use_v2 = false;      // bulk-async staged kernels (3-D); NSEM_KERNELS=v1 forces the plain-load kernels
    bool use_v4 = false;      // persistent software-pipelined kernels (3-D, default); NSEM_KERNELS=v2|v1 select the older generations
    bool pref_v2 = false, pref_v4 = false;   // what nsem_set_order selected; a non-conforming mesh falls back to v1
    bool tri = false;         // v4: metrics evaluated on the fly from the element's trilinear map (verified at upload)
    int numSMs = 148;
    double X[3][MAXN];        // LGL nodes

    int NX = 0, NY = 0, NZ = 0, NP = 0, NPF = 0, NPS = 0, GPS = 0;
    bool have_basis = false, have_mesh = false, have_params = false, have_state = false, have_ref = false, have_bcs = false;
    double D[3][MAXN * MAXN], W[3][MAXN];
    nsem_params prm{};

    uint32_t nB = 0, nG = 0, nF = 0, nAll = 0;
    size_t nNodes = 0;        // device nodes: nB*NPS + nG*GPS
    uint64_t ghostBase = 0;
    uint64_t nRefNodes = 0;   // n_cells_all * NP

    // host copies needed after upload_mesh
    std::vector<uint32_t> h_face_owner, h_face_neigh, h_bOwner;
    std::vector<uint8_t> h_bFid;

    // node arrays
    DevBuf<double> rho[2], U[2][3], T[2], S[2], p, GU[9], GT[3], Jinv[9], cV, rho_ref, p_ref, gfield[3], gh, diagPartial;
    // explicit scalar advection (apps/convection): the scalar lives in the rho slot, lambdaMax = cds(|U|)/2 (no sound speed: R = 0 in the kernels)
    bool convection = false;
    int conv_init = 0;                 // 0 = the wind is the uploaded U, 1 = LEVEQUE (re-evaluated at every step)
    double conv_etime = 1.0;           // end_step * dt: the period of the analytic wind
    long conv_step = 0;                // steps taken (Iteration::get_step() - 1)
    int conv_scheme = 0;               // 0 RUSANOV, 1 CDS, 2 UDS, 3 BLENDED (Controls::convection_scheme as the convection app's divf reads it)
    double conv_blend = 0.2;
    int ab_order = 1;                  // Adams-Bashforth order of the scalar's update (AB2..AB5 keep a residual history; 1 = one forward-Euler stage)
    int ab_stored = 0, ab_head = 0;    // entries of the history in use (MeshField::nstored), ring position of PREV(0)
    DevBuf<double> abHist[5];
    DevBuf<double> xyz[3];             // node coordinates (analytic wind only)
    bool speed_valid = false;  // S[cur] = |U| + c of the current state (kept by sweep B + the ghost update; recomputed after the state was set from outside)
    bool has_gh = false;
    int cur = 0;
    bool has_gfield = false;
    double sphere_radius = 0;
    // element-face tables
    DevBuf<uint32_t> faceOther, faceMeta, sched;
    DevBuf<double> faceVec, faceUnit;
    DevBuf<FaceRec> faceRec;
    DevBuf<ElemRec> elemRec;
    DevBuf<double> traceA, bVec;     // face traces of sweep A (v2), area vectors of the boundary faces
    bool has_sched = false;
    // non-conforming (mortar) faces: groups = coarse faces, subs = their sub-facets (nsem_mortar.cuh)
    DevBuf<MortarGroup> mortarGroups;
    DevBuf<MortarSub> mortarSubs;
    DevBuf<double> psiRef, psiCor, mortarA, mortarB;
    uint32_t nMortarGroups = 0, nMortarSubs = 0;
    // ghost tables
    DevBuf<uint32_t> ghostRef, bOwner;
    DevBuf<uint8_t> bFid;
    DevBuf<double> bUnit;
    DevBuf<uint8_t> bcKind[4];
    DevBuf<uint32_t> bcRec[4], bcPeer[4];
    DevBuf<double> bcFixed[4];
    DevBuf<BCRec> bcRecs;
    // staging for layout conversion
    DevBuf<double> stage;
    // pipelined transfers (nsem_upload_state_async / nsem_download_state_async): one staging buffer per direction, copies on their own
    // streams so that the download of one batch overlaps the upload of the next (PCIe is full duplex)
    DevBuf<double> stageIn, stageOut;          // [6][n_cells_all*NP]: rho, U(3), T, p in the reference layout
    DevBuf<double*> ptrTabAsync;               // [4 fields][3] device pointers for the conversion kernels
    DevBuf<int> compMapAsync;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t evH2D = nullptr, evScatter = nullptr, evGather = nullptr, evD2H = nullptr;
    bool asyncReady = false, scatterPending = false, d2hPending = false;
    DevBuf<double*> ptrTab;
    DevBuf<int> compMap;

    // host arrays the caller declares long-lived (nsem_pin_host) are page-locked so the PCIe copies run at full rate
    std::vector<std::pair<const void*, size_t>> pinned;
    bool pin(const void* p, size_t bytes) {
        for (auto& r : pinned)
            if (r.first == p && r.second >= bytes) return true;
        if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) {
            cudaGetLastError();     // pageable copies still work
            return false;
        }
        pinned.emplace_back(p, bytes);
        return true;
    }

    // timing
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};

#ifdef NSEM_WITH_NCCL
    ncclComm_t nccl = nullptr;
#endif
    struct Peer { int rank; uint32_t g0, nf; uint64_t off; };   // ghost cells [g0, g0+nf) are filled by this peer; off = slot offset
    std::vector<Peer> peers;
    uint64_t nSendSlots = 0;
    DevBuf<double> sendBuf;          // [16][nSendSlots]
    DevBuf<uint32_t> sendNodes;      // owner node (device index) per send slot
    // overlap of the halo exchange with the interior elements (mul(), field.h:2381-2415 does interior cells first too)
    DevBuf<uint32_t> schedInt, schedHalo;     // elements without / with an inter-partition face
    uint32_t nInt = 0, nHalo = 0;
    cudaEvent_t evA = nullptr, evCA = nullptr, evB = nullptr, evCB = nullptr;
    bool cbPending = false;                   // an exchange of U_new/T_new is in flight on the comm stream
    // halo over peer memory (nsem_halo.cuh): my receive window, the neighbours' windows mapped through CUDA IPC
    bool p2p = false;
    void* winBase = nullptr;
    uint64_t nRecvSlots = 0;
    struct Remote { void* base = nullptr; bool ipc = false; uint64_t nRecv = 0, offForMe = 0; uint32_t myIdx = 0; };
    std::vector<Remote> remotes;
    unsigned long long haloEpoch[HALO_KINDS] = {0, 0, 0};
    // halo fused into the persistent sweeps: tables for (kind 0 | 1) x parity, per ghost cell {neighbour, first slot in its window},
    // schedule with the partition-boundary elements first
    bool fused = false;
    DevBuf<HaloFuse> haloFuse;                // [2 kinds][2 parities]
    DevBuf<uint32_t> haloGhost, schedFused;
    bool overlap = false;                     // NSEM_OVERLAP=1: halo elements first, exchange overlapped with the interior (measured slower)
};

static int halo_check(nsem_ctx* c);       // halo over peer memory: did a neighbour fail to deliver? (defined with nsem_set_halo)

// ---------------------------------------------------------------------------------------------------------
// dispatch over the compiled (NX,NY,NZ) instantiations
// ---------------------------------------------------------------------------------------------------------
#ifdef NSEM_ONLY_ORDER4_3D      // kernel-variant experiments (profiles/): compile the bench instantiation only
#define NSEM_ORDERS(X) X(5, 5, 5)
#else
#define NSEM_ORDERS(X)                                                                            \
    X(2, 2, 2) X(3, 3, 3) X(4, 4, 4) X(5, 5, 5) X(6, 6, 6) X(7, 7, 7) X(8, 8, 8)                  \
    X(2, 1, 2) X(3, 1, 3) X(4, 1, 4) X(5, 1, 5) X(6, 1, 6) X(7, 1, 7) X(8, 1, 8)                  \
    X(2, 2, 1) X(3, 3, 1) X(4, 4, 1) X(5, 5, 1) X(6, 6, 1) X(7, 7, 1) X(8, 8, 1)                  \
    X(2, 1, 1) X(3, 1, 1) X(4, 1, 1) X(5, 1, 1) X(6, 1, 1) X(7, 1, 1) X(8, 1, 1)
#endif

template <int NX, int NY, int NZ>
struct Launch {
    using Dm = Dims<NX, NY, NZ>;
    // elements per block: aim at ~256 threads
    static constexpr int EPB = (Dm::TPE >= 256) ? 1 : (256 / Dm::TPE);

    static constexpr size_t smemA(bool visc) { return (size_t)EPB * (visc ? 7 : 3) * Dm::NP * sizeof(double); }
    static constexpr size_t smemB = (size_t)EPB * 12 * Dm::NP * sizeof(double);

    template <class K>
    static cudaError_t go(K kernel, size_t smem, const KParams& P, cudaStream_t s) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        const unsigned grid = (P.nB + EPB - 1) / EPB;
        kernel<<<grid, EPB * Dm::TPE, smem, s>>>(P);
        return cudaGetLastError();
    }
    static cudaError_t sweepA(const KParams& P, cudaStream_t s) {
        if (P.visc) return go(sweepA_kernel<NX, NY, NZ, EPB, true>, smemA(true), P, s);
        return go(sweepA_kernel<NX, NY, NZ, EPB, false>, smemA(false), P, s);
    }
    static cudaError_t sweepB(const KParams& P, cudaStream_t s) {
        if (P.visc) return go(sweepB_kernel<NX, NY, NZ, EPB, true>, smemB, P, s);
        return go(sweepB_kernel<NX, NY, NZ, EPB, false>, smemB, P, s);
    }
    // ---- v2: bulk-async staged, dense face tasks (3-D only) ----
    static constexpr int EPB2 = (Dm::NP >= 100) ? 1 : (Dm::NP >= 48 ? 2 : 4);
    using C2 = v2::Cfg<NX, NY, NZ, EPB2>;
    static constexpr bool has_v2 = (NX > 1 && NY > 1 && NZ > 1) && C2::smemB(true) <= 227 * 1024;
    template <class K>
    static cudaError_t go2(K kernel, size_t smem, const KParams& P, cudaStream_t s) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        const unsigned grid = (P.nB + EPB2 - 1) / EPB2;
        kernel<<<grid, C2::NT, smem, s>>>(P);
        return cudaGetLastError();
    }
    static cudaError_t sweepA2(const KParams& P, cudaStream_t s) {
        if constexpr (has_v2) {
            if (P.visc) return go2(v2::sweepA_v2<NX, NY, NZ, EPB2, true, C2::minb(C2::smemA(true))>, C2::smemA(true), P, s);
            return go2(v2::sweepA_v2<NX, NY, NZ, EPB2, false, C2::minb(C2::smemA(false))>, C2::smemA(false), P, s);
        } else return cudaErrorInvalidValue;
    }
    static cudaError_t sweepB2(const KParams& P, cudaStream_t s) {
        if constexpr (has_v2) {
            if (P.visc) return go2(v2::sweepB_v2<NX, NY, NZ, EPB2, true, C2::minb(C2::smemB(true), 128)>, C2::smemB(true), P, s);
            return go2(v2::sweepB_v2<NX, NY, NZ, EPB2, false, C2::minb(C2::smemB(false), 128)>, C2::smemB(false), P, s);
        } else return cudaErrorInvalidValue;
    }
    // ---- v4: persistent, software-pipelined (cubic 3-D orders) ----
    static constexpr bool has_v4 = (NX == NY && NY == NZ && NX > 1) && v4::CfgB<NX, NY, NZ, true, false>::smem <= 227 * 1024 &&
                                   v4::CfgA<NX, NY, NZ, true, false>::smem <= 227 * 1024 && v4::CfgA<NX, NY, NZ, true, false>::ok;
    // shared-memory carve-out of the v4 sweeps.  Sweep A gathers its neighbour values with 8-byte cp.async through L1, so it asks for no more
    // shared memory than its resident CTAs need (order 4: 4 x 41 KB = 71 % -> the 164 KB configuration instead of 228 KB: sweep A -11 %,
    // profiles/r1_variants.md); sweep B streams everything by bulk copies and keeps the maximum.  NSEM_V4_CARVEOUT_A / _B (percent of
    // 228 KB) override.
    static int carveout(const char* var, int default_pct) {
        const char* v = std::getenv(var);
        const int pct = v ? std::atoi(v) : default_pct;
        return (pct > 0 && pct <= 100) ? pct : (int)cudaSharedmemCarveoutMaxShared;
    }
    static constexpr int carve_pct(size_t smem, int minb) {
        // smallest percentage of 228 KB that holds minb CTAs (1 KB reserved per CTA)
        return (int)(((smem + 1024) * (size_t)minb * 100 + 228 * 1024 - 1) / (228 * 1024));
    }
    template <class K>
    static cudaError_t go4(K kernel, size_t smem, int nt, int minb, const KParams& P, int sms, cudaStream_t s, const char* carve_var, int carve_default) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout(carve_var, carve_default));
        if (e != cudaSuccess) return e;
        const uint64_t want = (uint64_t)sms * (uint64_t)minb;
        const unsigned grid = (unsigned)std::min<uint64_t>(P.nB, want);
        if (grid == 0) return cudaSuccess;
        kernel<<<grid, nt, smem, s>>>(P);
        return cudaGetLastError();
    }
    template <bool VISC, bool TRI>
    static cudaError_t sweepA4t(const KParams& P, int sms, cudaStream_t s) {
        using C4 = v4::CfgA<NX, NY, NZ, VISC, TRI>;
        constexpr int MB = C4::minb(NSEM_V4_REGS_A);
        if (P.mortarA) return go4(v4::sweepA_v4<NX, NY, NZ, VISC, TRI, MB, true>, C4::smem, C4::NT, MB, P, sms, s, "NSEM_V4_CARVEOUT_A", carve_pct(C4::smem, MB));     // non-conforming mesh
        return go4(v4::sweepA_v4<NX, NY, NZ, VISC, TRI, MB>, C4::smem, C4::NT, MB, P, sms, s, "NSEM_V4_CARVEOUT_A", carve_pct(C4::smem, MB));
    }
    template <bool VISC, bool TRI>
    static cudaError_t sweepB4t(const KParams& P, int sms, cudaStream_t s) {
        using C4 = v4::CfgB<NX, NY, NZ, VISC, TRI>;
        constexpr int MB = C4::minb(NSEM_V4_REGS_B);
        if (P.mortarB) return go4(v4::sweepB_v4<NX, NY, NZ, VISC, TRI, MB, true>, C4::smem, C4::NT, MB, P, sms, s, "NSEM_V4_CARVEOUT_B", 0);
        return go4(v4::sweepB_v4<NX, NY, NZ, VISC, TRI, MB>, C4::smem, C4::NT, MB, P, sms, s, "NSEM_V4_CARVEOUT_B", 0);
    }
    static cudaError_t sweepA4(const KParams& P, bool tri, int sms, cudaStream_t s) {
        if constexpr (has_v4) {
            if (P.visc) return tri ? sweepA4t<true, true>(P, sms, s) : sweepA4t<true, false>(P, sms, s);
            return tri ? sweepA4t<false, true>(P, sms, s) : sweepA4t<false, false>(P, sms, s);
        } else return cudaErrorInvalidValue;
    }
    static cudaError_t sweepB4(const KParams& P, bool tri, int sms, cudaStream_t s) {
        if constexpr (has_v4) {
            if (P.visc) return tri ? sweepB4t<true, true>(P, sms, s) : sweepB4t<true, false>(P, sms, s);
            return tri ? sweepB4t<false, true>(P, sms, s) : sweepB4t<false, false>(P, sms, s);
        } else return cudaErrorInvalidValue;
    }
    static cudaError_t ghost_trace(const GhostTraceParams& G, cudaStream_t s) {
        const uint64_t n = (uint64_t)G.nG * Dm::NPF;
        if (n == 0) return cudaSuccess;
        ghost_trace_kernel<NX, NY, NZ><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(G);
        return cudaGetLastError();
    }
    static cudaError_t bc(const BCParams& B, cudaStream_t s) {
        const uint64_t n = (uint64_t)B.nG * Dm::NPF;
        if (n == 0) return cudaSuccess;
        bc_kernel<NX, NY, NZ><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(B);
        return cudaGetLastError();
    }
};

static bool order_supported(int nx, int ny, int nz) {
#define X(a, b, c) if (nx == a && ny == b && nz == c) return true;
    NSEM_ORDERS(X)
#undef X
    return false;
}

static bool has_v2(int nx, int ny, int nz) {
#define X(a, b, c) if (nx == a && ny == b && nz == c) return Launch<a, b, c>::has_v2;
    NSEM_ORDERS(X)
#undef X
    return false;
}

static bool has_v4(int nx, int ny, int nz) {
#define X(a, b, c) if (nx == a && ny == b && nz == c) return Launch<a, b, c>::has_v4;
    NSEM_ORDERS(X)
#undef X
    return false;
}

static cudaError_t launch_sweepA(const nsem_ctx* c, const KParams& P) {
#define X(a, b, cc) if (c->NX == a && c->NY == b && c->NZ == cc) return c->use_v4 ? Launch<a, b, cc>::sweepA4(P, c->tri, c->numSMs, c->stream) : c->use_v2 ? Launch<a, b, cc>::sweepA2(P, c->stream) : Launch<a, b, cc>::sweepA(P, c->stream);
    NSEM_ORDERS(X)
#undef X
    return cudaErrorInvalidValue;
}
static cudaError_t launch_sweepB(const nsem_ctx* c, const KParams& P) {
#define X(a, b, cc) if (c->NX == a && c->NY == b && c->NZ == cc) return c->use_v4 ? Launch<a, b, cc>::sweepB4(P, c->tri, c->numSMs, c->stream) : c->use_v2 ? Launch<a, b, cc>::sweepB2(P, c->stream) : Launch<a, b, cc>::sweepB(P, c->stream);
    NSEM_ORDERS(X)
#undef X
    return cudaErrorInvalidValue;
}
static cudaError_t launch_ghost_trace(const nsem_ctx* c, const KParams& P) {
    if (!(c->use_v4 || c->use_v2)) return cudaSuccess;       // only the v2/v4 sweep B consumes face traces
    GhostTraceParams G;
    std::memset(&G, 0, sizeof G);
    G.nB = c->nB; G.nG = c->nG; G.ghostBase = c->ghostBase;
    G.T0 = P.T0; G.nu = P.nu; G.iPr = P.iPr; G.gammaR = P.gamma * P.R; G.visc = P.visc;
    std::memcpy(G.W, c->W, sizeof G.W);
    G.bFid = c->bFid.p; G.bVec = c->bVec.p;
    G.rho_old = P.rho_old; G.rho_new = P.rho_new; G.T_old = P.T_old; G.p = P.p;
    for (int d = 0; d < 3; d++) { G.U_old[d] = P.U_old[d]; G.GT[d] = P.GT[d]; }
    for (int d = 0; d < 9; d++) G.GU[d] = P.GU[d];
    G.traceA = c->traceA.p;
#define X(a, b, cc) if (c->NX == a && c->NY == b && c->NZ == cc) return Launch<a, b, cc>::ghost_trace(G, c->stream);
    NSEM_ORDERS(X)
#undef X
    return cudaErrorInvalidValue;
}
static cudaError_t launch_bc(const nsem_ctx* c, const BCParams& B) {
#define X(a, b, cc) if (c->NX == a && c->NY == b && c->NZ == cc) return Launch<a, b, cc>::bc(B, c->stream);
    NSEM_ORDERS(X)
#undef X
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------------
// life cycle
// ---------------------------------------------------------------------------------------------------------
extern "C" int nsem_create(int device, int rank, int nranks, const void* nccl_unique_id, nsem_ctx** out) {
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("nsem_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return 1;
    }
    if (device < 0) device = rank % ndev;      // one process per GPU of the node, ranks round-robin over the visible devices
    if (device >= ndev) {
        g_create_error = "nsem_create: device index out of range";
        return 1;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return 1;
    }
    nsem_ctx* c = new nsem_ctx();
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->numSMs = sms;
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->comm, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "nsem_create: cudaStreamCreate failed";
        delete c;
        return 1;
    }
    for (auto& ev : c->ev) cudaEventCreate(&ev);
    if (nranks > 1) {
#ifdef NSEM_WITH_NCCL
        if (!nccl_unique_id) {
            g_create_error = "nsem_create: nranks > 1 needs an ncclUniqueId";
            delete c;
            return 1;
        }
        if (!g_nccl.load()) {
            g_create_error = "nsem_create: " + g_nccl.err;
            delete c;
            return 1;
        }
        ncclUniqueId id;
        std::memcpy(&id, nccl_unique_id, sizeof(id));
        ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
        if (r != ncclSuccess) {
            g_create_error = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
            delete c;
            return 1;
        }
#else
        (void)nccl_unique_id;
        g_create_error = "nsem_create: library built without NCCL";
        delete c;
        return 1;
#endif
    }
    *out = c;
    return 0;
}

static void halo_p2p_release(nsem_ctx* c) {
    for (auto& r : c->remotes)
        if (r.base && r.ipc) cudaIpcCloseMemHandle(r.base);
    c->remotes.clear();
    if (c->winBase) cudaFree(c->winBase);
    c->winBase = nullptr;
    c->p2p = false;
    c->fused = false;
}

extern "C" void nsem_destroy(nsem_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
#ifdef NSEM_WITH_NCCL
    if (c->nccl) g_nccl.CommDestroy(c->nccl);
#endif
    halo_p2p_release(c);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& r : c->pinned) cudaHostUnregister(const_cast<void*>(r.first));
    for (cudaEvent_t e : {c->evH2D, c->evScatter, c->evGather, c->evD2H}) if (e) cudaEventDestroy(e);
    if (c->h2d) cudaStreamDestroy(c->h2d);
    if (c->d2h) cudaStreamDestroy(c->d2h);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm) cudaStreamDestroy(c->comm);
    delete c;
}

extern "C" const char* nsem_last_error(const nsem_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

extern "C" int nsem_device(const nsem_ctx* c) { return c ? c->device : -1; }

extern "C" int nsem_get_unique_id(void* out128) {
#ifdef NSEM_WITH_NCCL
    ncclUniqueId id;
    if (!g_nccl.load()) { g_create_error = g_nccl.err; return 1; }
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return 1;
    std::memcpy(out128, &id, sizeof(id));
    return 0;
#else
    (void)out128;
    return 1;
#endif
}

extern "C" uint64_t nsem_launch_count(const nsem_ctx* c) { return c->launches; }

// transport of the face-trace halo: "peer memory" (stores into the neighbours' windows over NVLink, nsem_halo.cuh), "nccl", or "none"
extern "C" const char* nsem_halo_info(const nsem_ctx* c) {
    if (c->nranks <= 1 || c->peers.empty()) return "none";
    return c->p2p ? "peer memory (CUDA IPC windows, NVLink stores + flags)" : "nccl send/recv";
}

extern "C" const char* nsem_kernel_info(const nsem_ctx* c) {
    if (c->use_v4 && c->nMortarGroups)
        return c->tri ? "v4 persistent pipelined, metrics on the fly (trilinear map verified) + mortar (non-conforming) face kernels"
                      : "v4 persistent pipelined, stored metrics + mortar (non-conforming) face kernels";
    if (c->use_v4) return c->tri ? "v4 persistent pipelined, metrics on the fly (trilinear map verified)" : "v4 persistent pipelined, stored metrics";
    if (c->use_v2) return "v2 bulk-async staged";
    if (c->nMortarGroups) return "v1 plain loads + mortar (non-conforming) face kernels";
    return "v1 plain loads";
}

extern "C" int nsem_sync(nsem_ctx* c) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->h2d) CUDA_TRY(c, cudaStreamSynchronize(c->h2d));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->comm));
    if (c->d2h) CUDA_TRY(c, cudaStreamSynchronize(c->d2h));
    return halo_check(c);
}

// ---------------------------------------------------------------------------------------------------------
// set-up
// ---------------------------------------------------------------------------------------------------------
extern "C" int nsem_set_order(nsem_ctx* c, int NPX, int NPY, int NPZ) {
    if (!order_supported(NPX, NPY, NPZ)) {
        char b[160];
        snprintf(b, sizeof b, "nsem_set_order: (%d,%d,%d) points per direction has no compiled kernel "
                 "(3-D n^3, 2-D n x 1 x n and n x n x 1 for n = 2..8)", NPX, NPY, NPZ);
        c->err = b;
        return 1;
    }
    c->NX = NPX; c->NY = NPY; c->NZ = NPZ;
    c->NP = NPX * NPY * NPZ;
    c->NPF = (NPX <= NPY && NPX <= NPZ) ? NPY * NPZ : ((NPY <= NPX && NPY <= NPZ) ? NPX * NPZ : NPX * NPY);
    c->NPS = pad_to(c->NP, 16);
    c->GPS = pad_to(c->NPF, 4);
    c->have_basis = c->have_mesh = c->have_state = c->have_ref = c->have_bcs = false;
    const char* kv = std::getenv("NSEM_KERNELS");
    c->use_v2 = has_v2(NPX, NPY, NPZ) && !(kv && std::strcmp(kv, "v1") == 0);
    c->use_v4 = has_v4(NPX, NPY, NPZ) && !(kv && (std::strcmp(kv, "v1") == 0 || std::strcmp(kv, "v2") == 0));
    c->pref_v2 = c->use_v2; c->pref_v4 = c->use_v4;
    c->tri = false;
    return 0;
}

extern "C" int nsem_set_basis(nsem_ctx* c, const double* const dpsi[3], const double* const wgl[3]) {
    if (!c->NP) { c->err = "nsem_set_basis: call nsem_set_order first"; return 1; }
    const int n[3] = {c->NX, c->NY, c->NZ};
    std::memset(c->D, 0, sizeof c->D);
    std::memset(c->W, 0, sizeof c->W);
    for (int d = 0; d < 3; d++) {
        for (int q = 0; q < n[d] * n[d]; q++) c->D[d][q] = dpsi[d][q];
        for (int q = 0; q < n[d]; q++) c->W[d][q] = wgl[d][q];
    }
    // Legendre-Gauss-Lobatto nodes (dg.cpp:53-99 computes the same points; here they are only used to evaluate the
    // trilinear map of straight-edged elements, and that evaluation is verified against the uploaded Jinv)
    std::memset(c->X, 0, sizeof c->X);
    for (int d = 0; d < 3; d++) {
        const int N = n[d] - 1;
        if (N < 1) continue;
        for (int q = 0; q <= N; q++) {
            double x = -std::cos(M_PI * q / N);
            for (int it = 0; it < 100; it++) {
                double p0 = 1, p1 = x;
                for (int m = 2; m <= N; m++) { const double p2 = ((2 * m - 1) * x * p1 - (m - 1) * p0) / m; p0 = p1; p1 = p2; }
                const double dx = (N == 1) ? 0.0 : (x * p1 - p0) / ((N + 1) * p1);
                x -= dx;
                if (std::fabs(dx) < 1e-16) break;
            }
            c->X[d][q] = x;
        }
        c->X[d][0] = -1.0; c->X[d][N] = 1.0;
        for (int q = 0; q <= N / 2; q++) {          // symmetrise
            const double v = 0.5 * (c->X[d][N - q] - c->X[d][q]);
            c->X[d][q] = -v; c->X[d][N - q] = v;
        }
    }
    c->have_basis = true;
    return 0;
}

extern "C" int nsem_set_params(nsem_ctx* c, const nsem_params* p) {
    if (!(p->dt > 0) || !(p->cv > 0) || !(p->cp > p->cv) || !(p->P0 > 0) || !(p->Pr > 0)) {
        c->err = "nsem_set_params: need dt > 0, cp > cv > 0, P0 > 0, Pr > 0";
        return 1;
    }
    c->prm = *p;
    c->have_params = true;
    return 0;
}

// local node of face `fid` at face coordinates (a,b)
static inline int h_face_node(const nsem_ctx* c, int fid, int a, int b) {
    const int NX = c->NX, NY = c->NY, NZ = c->NZ;
    if (fid < 2) return a * NY * NZ + b * NZ + (fid == 0 ? 0 : NZ - 1);
    if (fid < 4) return a * NY * NZ + (fid == 2 ? 0 : NY - 1) * NZ + b;
    return (fid == 4 ? 0 : NX - 1) * NY * NZ + a * NZ + b;
}

extern "C" int nsem_upload_mesh(nsem_ctx* c, const nsem_mesh* m) {
    if (!c->have_basis) { c->err = "nsem_upload_mesh: call nsem_set_order and nsem_set_basis first"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int NX = c->NX, NY = c->NY, NZ = c->NZ, NP = c->NP, NPF = c->NPF, NPS = c->NPS, GPS = c->GPS;
    const uint32_t nB = m->n_cells_real, nAll = m->n_cells_all, nF = m->n_faces, nG = nAll - nB;
    if ((uint64_t)nB * NPS + (uint64_t)nG * GPS >= 0xffffffffull) {
        c->err = "nsem_upload_mesh: more than 2^32 device nodes on one GPU";
        return 1;
    }
    c->nB = nB; c->nG = nG; c->nF = nF; c->nAll = nAll;
    c->ghostBase = (uint64_t)nB * NPS;
    c->nNodes = (size_t)nB * NPS + (size_t)nG * GPS;
    c->nRefNodes = (uint64_t)nAll * NP;
    const uint32_t sentinel = (uint32_t)c->nRefNodes;

    // ---- per (element, local face id) tables ----
    std::vector<uint32_t> fOther((size_t)nB * 6, 0), fMeta((size_t)nB * 6, FM_ABSENT);
    std::vector<double> fVec((size_t)nB * 18, 0.0), fUnit((size_t)nB * 18, 0.0);
    std::vector<uint32_t> ghostRef((size_t)nG * NPF, 0xffffffffu), bOwner(nG, 0);
    std::vector<uint8_t> bFid(nG, 0);
    std::vector<double> bUnit((size_t)nG * 3, 0.0);
    // face id of a face as seen from a real cell
    auto local_id = [&](uint32_t cell, uint32_t face) -> int {
        for (uint32_t f = m->face_begin[cell]; f < m->face_end[cell]; f++)
            if (m->all_faces[f] == face) return (int)m->face_id[f];
        return -1;
    };
    // ---- non-conforming (mortar) faces: groups = coarse faces with their sub-facets (nsem_mortar.cuh) ----
    bool anyMortar = false;
    if (m->face_mortar)
        for (uint32_t f = 0; f < nF && !anyMortar; f++) anyMortar = m->face_mortar[f] != 0;
    if (anyMortar) {
        bool have = m->cC && m->face_center;
        for (int q = 0; q < 6; q++) have = have && m->psi_ref[q] && m->psi_cor[q];
        if (!have) { c->err = "nsem_upload_mesh: a non-conforming mesh needs cC, face_center, psi_ref and psi_cor"; return 1; }
    }
    // the element sweeps that know the FM_MORTAR branch: the persistent ones (v4, cubic 3-D orders; NSEM_MORTAR_V1=1 keeps them off) and
    // the plain-load ones (v1, every order)
    {
        const char* mv1 = std::getenv("NSEM_MORTAR_V1");
        const bool v4ok = !(mv1 && std::strcmp(mv1, "1") == 0);
        c->use_v2 = c->pref_v2 && !anyMortar;
        c->use_v4 = c->pref_v4 && (!anyMortar || v4ok);
    }
    std::vector<MortarGroup> mGroups;
    std::vector<std::vector<MortarSub>> mSubs;          // per group, in the coarse cell's face order
    std::vector<int32_t> mBlock(anyMortar ? (size_t)nB * 6 : 0, -1), mGroupOf(anyMortar ? (size_t)nB * 6 : 0, -1);
    uint32_t nMortarBlocks = 0;
    auto mortar_block = [&](uint32_t cell, int sid) -> uint32_t {
        int32_t& b = mBlock[(size_t)cell * 6 + sid];
        if (b < 0) b = (int32_t)nMortarBlocks++;
        return (uint32_t)b;
    };
    auto face_dims = [&](int fid, int& d1, int& d2, int& n1, int& n2) {
        if (fid < 2) { d1 = 0; d2 = 1; n1 = NX; n2 = NY; }
        else if (fid < 4) { d1 = 0; d2 = 2; n1 = NX; n2 = NZ; }
        else { d1 = 1; d2 = 2; n1 = NY; n2 = NZ; }
    };
    for (uint32_t ci = 0; ci < nB; ci++) {
        for (uint32_t f = m->face_begin[ci]; f < m->face_end[ci]; f++) {
            const uint32_t face = m->all_faces[f];
            const int sid = (int)m->face_id[f];
            if (sid < 0 || sid > 5) { c->err = "nsem_upload_mesh: face id outside 0..5"; return 1; }
            if (anyMortar && m->face_mortar[face] != 0) {
                const uint32_t fm = m->face_mortar[face];
                const uint32_t fo = m->face_owner[face], fn = m->face_neigh[face];
                const uint32_t fine = (fm == 1) ? fo : fn, coarse = (fm == 1) ? fn : fo;      // field.h:2158-2159
                if (fm > 2 || fine >= nB || coarse >= nB || (ci != fine && ci != coarse)) {
                    c->err = "nsem_upload_mesh: inconsistent gFMC/gFOC/gFNC on a non-conforming face";
                    return 1;
                }
                const size_t e6 = (size_t)ci * 6 + sid;
                const bool own = (fo == ci);
                int d1, d2, n1, n2;
                face_dims(sid, d1, d2, n1, n2);
                // my side of the reference's node maps on this facet: slot (a,b) <-> my face node (a,b)  (dg.cpp:372-404)
                for (int a = 0; a < n1; a++)
                    for (int b = 0; b < n2; b++) {
                        const size_t k = (size_t)face * NPF + (size_t)a * n2 + b;
                        const uint32_t mine = ci * (uint32_t)NP + (uint32_t)h_face_node(c, sid, a, b);
                        if ((own ? m->FO[k] : m->FN[k]) != mine) {
                            c->err = "nsem_upload_mesh: a non-conforming facet does not follow the tensor-product node pairing of dg.cpp:372-404";
                            return 1;
                        }
                    }
                fOther[e6] = mortar_block(ci, sid);
                fMeta[e6] = FM_MORTAR | FM_ABSENT | FM_HALF | (own ? FM_OWNER : 0u);
                if (ci == fine) continue;        // the sub-facet record is written from the coarse side
                // ---- coarse side: one group per (coarse cell, local face), sub-facets in the cell's face order ----
                int32_t& gi = mGroupOf[e6];
                if (gi < 0) {
                    gi = (int32_t)mGroups.size();
                    MortarGroup g;
                    std::memset(&g, 0, sizeof g);
                    g.coarse = ci * (uint32_t)NPS; g.fid_c = (uint32_t)sid; g.block = fOther[e6];
                    mGroups.push_back(g);
                    mSubs.emplace_back();
                }
                const int fidf = local_id(fine, face);
                if (fidf < 0) { c->err = "nsem_upload_mesh: face not found in its fine cell"; return 1; }
                int e1, e2, m1, m2;
                face_dims(fidf, e1, e2, m1, m2);
                if (m1 != n1 || m2 != n2) { c->err = "nsem_upload_mesh: the two sides of a non-conforming facet have different node extents"; return 1; }
                // half of the coarse face this sub-facet covers along each face axis (field.h:2174-2197)
                double cco[3], ccn[3] = {0, 0, 0};
                for (int d = 0; d < 3; d++) cco[d] = m->face_center[(size_t)face * 3 + d];
                int nch = 0;
                for (uint32_t r = m->face_begin[ci]; r < m->face_end[ci]; r++)
                    if ((int)m->face_id[r] == sid) {
                        for (int d = 0; d < 3; d++) ccn[d] += m->face_center[(size_t)m->all_faces[r] * 3 + d];
                        nch++;
                    }
                for (int d = 0; d < 3; d++) ccn[d] /= (double)nch;
                const double* v0 = m->cC + ((size_t)ci * NP + h_face_node(c, sid, 0, 0)) * 3;
                const double* v1 = m->cC + ((size_t)ci * NP + h_face_node(c, sid, n1 - 1, 0)) * 3;
                const double* v2 = m->cC + ((size_t)ci * NP + h_face_node(c, sid, 0, n2 - 1)) * 3;
                auto dot3 = [](const double* a, const double* o, const double* b) {
                    return ((a[0] - o[0]) * (b[0] - o[0]) + (a[1] - o[1]) * (b[1] - o[1])) + (a[2] - o[2]) * (b[2] - o[2]);
                };
                const uint32_t h1 = (dot3(ccn, v0, v1) >= dot3(cco, v0, v1)) ? 0u : 1u;
                const uint32_t h2 = (dot3(ccn, v0, v2) >= dot3(cco, v0, v2)) ? 0u : 1u;
                MortarSub sb;
                std::memset(&sb, 0, sizeof sb);
                sb.fine = fine * (uint32_t)NPS; sb.fid_f = (uint32_t)fidf;
                sb.flags = h1 | (h2 << 1) | (fm == 1 ? 4u : 0u);
                sb.block = mortar_block(fine, fidf);
                const double* N = m->face_normal + (size_t)face * 3;
                const double mg = std::sqrt(N[0] * N[0] + (N[1] * N[1] + N[2] * N[2]));
                for (int d = 0; d < 3; d++) { sb.vec[d] = N[d]; sb.unit[d] = N[d] / mg; }
                mSubs[gi].push_back(sb);
                continue;
            }
            const size_t e6 = (size_t)ci * 6 + sid;
            if ((fMeta[e6] & FM_FID_MASK) != FM_ABSENT || (fMeta[e6] & FM_MORTAR)) {
                c->err = "nsem_upload_mesh: two faces with the same local id on one element that are not flagged in face_mortar";
                return 1;
            }
            const bool own = (m->face_owner[face] == ci);
            const uint32_t other = own ? m->face_neigh[face] : m->face_owner[face];
            uint32_t meta = own ? FM_OWNER : 0u;
            // fI at the first used slot of the face
            double alpha = -1;
            for (int n = 0; n < NPF; n++)
                if (m->FO[(size_t)face * NPF + n] != sentinel) { alpha = m->fI[(size_t)face * NPF + n]; break; }
            if (alpha == 0.5) meta |= FM_HALF;
            else if (alpha != 0.0) { c->err = "nsem_upload_mesh: fI must be 0.5 or 0 on a DG mesh"; return 1; }
            if (other < nB) {
                const int oid = local_id(other, face);
                if (oid < 0) { c->err = "nsem_upload_mesh: face not found in its neighbour cell"; return 1; }
                meta |= (uint32_t)oid;
                fOther[e6] = other * (uint32_t)NPS;
            } else {
                if (!own) { c->err = "nsem_upload_mesh: boundary cell owns a face"; return 1; }
                const uint32_t g = other - nB;
                meta |= FM_GHOST;
                fOther[e6] = (uint32_t)(c->ghostBase + (uint64_t)g * GPS);
                bOwner[g] = ci;
                bFid[g] = (uint8_t)sid;
            }
            fMeta[e6] = meta;
            const double* N = m->face_normal + (size_t)face * 3;
            const double mg = std::sqrt(N[0] * N[0] + (N[1] * N[1] + N[2] * N[2]));
            for (int d = 0; d < 3; d++) {
                fVec[e6 * 3 + d] = N[d];
                fUnit[e6 * 3 + d] = N[d] / mg;
            }
            if (other >= nB)
                for (int d = 0; d < 3; d++) bUnit[(size_t)(other - nB) * 3 + d] = N[d] / mg;

            // ---- verify the derived node pairing + weights against the reference maps ----
            const int na = (sid < 2) ? NX : (sid < 4 ? NX : NY);
            const int nb = (sid < 2) ? NY : NZ;
            for (int a = 0; a < na; a++)
                for (int b = 0; b < nb; b++) {
                    const int n = (sid < 2) ? a * NY + b : a * NZ + b;
                    const size_t k = (size_t)face * NPF + n;
                    const uint32_t mine = ci * (uint32_t)NP + (uint32_t)h_face_node(c, sid, a, b);
                    const uint32_t refMine = own ? m->FO[k] : m->FN[k];
                    const uint32_t refOther = own ? m->FN[k] : m->FO[k];
                    bool ok = (refMine == mine);
                    if (other < nB) {
                        const uint32_t theirs = other * (uint32_t)NP + (uint32_t)h_face_node(c, (int)(meta & FM_FID_MASK), a, b);
                        ok = ok && (refOther == theirs);
                    } else if (ok) {
                        ghostRef[(size_t)(other - nB) * NPF + n] = refOther;
                    }
                    if (!ok) {
                        char bmsg[200];
                        snprintf(bmsg, sizeof bmsg, "nsem_upload_mesh: face %u (cell %u, local id %d) does not follow the "
                                 "tensor-product node pairing of dg.cpp:372-404", face, ci, sid);
                        c->err = bmsg;
                        return 1;
                    }
                    const double w = (sid < 2) ? c->W[0][a] * c->W[1][b] / 4 : (sid < 4 ? c->W[0][a] * c->W[2][b] / 4 : c->W[1][a] * c->W[2][b] / 4);
                    for (int d = 0; d < 3; d++) {
                        const double want = m->fN[k * 3 + d], got = N[d] * w;
                        if (std::fabs(want - got) > 1e-12 * (std::fabs(want) + mg * w)) {
                            c->err = "nsem_upload_mesh: fN != face_normal * w_a*w_b/4";
                            return 1;
                        }
                    }
                }
        }
    }
    c->h_face_owner.assign(m->face_owner, m->face_owner + nF);
    c->h_face_neigh.assign(m->face_neigh, m->face_neigh + nF);
    c->h_bOwner = bOwner;
    c->h_bFid = bFid;

    cudaStream_t s = c->stream;
    {
        std::vector<MortarSub> subs;
        for (size_t g = 0; g < mGroups.size(); g++) {
            mGroups[g].sub0 = (uint32_t)subs.size();
            mGroups[g].nsub = (uint32_t)mSubs[g].size();
            subs.insert(subs.end(), mSubs[g].begin(), mSubs[g].end());
        }
        c->nMortarGroups = (uint32_t)mGroups.size();
        c->nMortarSubs = (uint32_t)subs.size();
        CUDA_TRY(c, c->mortarGroups.upload(mGroups, s));
        CUDA_TRY(c, c->mortarSubs.upload(subs, s));
        std::vector<double> pr(anyMortar ? 6 * MAXN * MAXN : 0, 0.0), pc(pr.size(), 0.0);
        if (anyMortar) {
            const int nd[3] = {NX, NY, NZ};
            for (int q = 0; q < 6; q++)
                for (int e = 0; e < nd[q / 2] * nd[q / 2]; e++) {
                    pr[(size_t)q * MAXN * MAXN + e] = m->psi_ref[q][e];
                    pc[(size_t)q * MAXN * MAXN + e] = m->psi_cor[q][e];
                }
        }
        CUDA_TRY(c, c->psiRef.upload(pr, s));
        CUDA_TRY(c, c->psiCor.upload(pc, s));
        CUDA_TRY(c, c->mortarA.alloc((size_t)nMortarBlocks * MORTAR_NA * MORTAR_MAXF));
        CUDA_TRY(c, c->mortarB.alloc((size_t)nMortarBlocks * MORTAR_NB * MORTAR_MAXF));
        if (nMortarBlocks) {
            CUDA_TRY(c, cudaMemsetAsync(c->mortarA.p, 0, c->mortarA.n * sizeof(double), s));
            CUDA_TRY(c, cudaMemsetAsync(c->mortarB.p, 0, c->mortarB.n * sizeof(double), s));
        }
        CUDA_TRY(c, cudaStreamSynchronize(s));
    }
    CUDA_TRY(c, c->faceOther.upload(fOther, s));
    CUDA_TRY(c, c->faceMeta.upload(fMeta, s));
    CUDA_TRY(c, c->faceVec.upload(fVec, s));
    CUDA_TRY(c, c->faceUnit.upload(fUnit, s));
    {
        std::vector<FaceRec> rec((size_t)nB * 6);
        for (size_t q = 0; q < rec.size(); q++) {
            rec[q].other = fOther[q];
            rec[q].meta = fMeta[q];
            for (int d = 0; d < 3; d++) { rec[q].vec[d] = fVec[q * 3 + d]; rec[q].unit[d] = fUnit[q * 3 + d]; }
            const uint32_t fid = fMeta[q] & FM_FID_MASK;
            if (fid == FM_GHOST) rec[q].otherBlock = (uint64_t)nB * 6 + (fOther[q] - (uint32_t)c->ghostBase) / (uint32_t)GPS;
            else if (fid == FM_ABSENT) rec[q].otherBlock = 0;
            else rec[q].otherBlock = (uint64_t)(fOther[q] / (uint32_t)NPS) * 6 + fid;
        }
        const size_t TBS = (size_t)trace_bs(NPF);
        CUDA_TRY(c, c->traceA.alloc(((size_t)nB * 6 + nG) * TBS));
        CUDA_TRY(c, cudaMemsetAsync(c->traceA.p, 0, ((size_t)nB * 6 + nG) * TBS * sizeof(double), s));
        std::vector<double> bv((size_t)nG * 3, 0.0);
        for (uint32_t ci2 = 0; ci2 < nB; ci2++)
            for (int sid = 0; sid < 6; sid++) {
                const size_t e6 = (size_t)ci2 * 6 + sid;
                if ((fMeta[e6] & FM_FID_MASK) == FM_GHOST) {
                    const uint32_t g = (fOther[e6] - (uint32_t)c->ghostBase) / (uint32_t)GPS;
                    for (int d = 0; d < 3; d++) bv[(size_t)g * 3 + d] = fVec[e6 * 3 + d];
                }
            }
        CUDA_TRY(c, c->bVec.upload(bv, s));
        CUDA_TRY(c, c->faceRec.upload(rec, s));
        CUDA_TRY(c, cudaStreamSynchronize(s));

        // ---- v4: element records = the six face records + the element's trilinear map ----
        // dg.cpp:176-325 places the nodes of a straight-edged hexahedron by trilinear interpolation of its corners, so
        // J = dx/dxi is a closed form of 7 coefficient vectors.  They are recovered from the uploaded Jinv at the 8 corner
        // nodes and the closed form is then checked against the uploaded Jinv and cV at EVERY node; a mesh that fails
        // (curved edges, spherical shells) runs the stored-metric instantiation instead.
        c->tri = false;
        if (c->use_v4) {
            std::vector<ElemRec> er(nB);
            static const int rm9[9] = {0, 4, 8, 1, 5, 2, 3, 7, 6};      // AoS component -> row-major a*3+d
            const char* mv = std::getenv("NSEM_METRICS");
            const bool want_tri = !(mv && std::strcmp(mv, "stored") == 0);
            const char* av = std::getenv("NSEM_AFFINE");
            const bool want_affine = !(av && std::strcmp(av, "0") == 0);
            const unsigned nth = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            std::vector<int> bad(nth, 0);
            auto work = [&](unsigned t) {
                const uint32_t e0 = (uint32_t)((uint64_t)nB * t / nth), e1 = (uint32_t)((uint64_t)nB * (t + 1) / nth);
                for (uint32_t e = e0; e < e1; e++) {
                    ElemRec& R = er[e];
                    for (int f = 0; f < 6; f++) R.face[f] = rec[(size_t)e * 6 + f];
                    for (int q = 0; q < 7; q++) for (int a = 0; a < 3; a++) R.c[q][a] = 0.0;
                    R.vol = 0.0;
                    if (!want_tri || bad[t]) continue;
                    auto loadM = [&](int node, double M[9]) {
                        const double* src = m->Jinv + ((size_t)e * NP + node) * 9;
                        for (int comp = 0; comp < 9; comp++) M[rm9[comp]] = src[comp];
                    };
                    double acc[7][3] = {};
                    bool ok = true;
                    for (int v = 0; v < 8 && ok; v++) {
                        const int sx = (v & 1) ? 1 : -1, sy = (v & 2) ? 1 : -1, sz = (v & 4) ? 1 : -1;
                        const int node = ((sx > 0) ? NX - 1 : 0) * NY * NZ + ((sy > 0) ? NY - 1 : 0) * NZ + ((sz > 0) ? NZ - 1 : 0);
                        double M[9], Ci[9];
                        loadM(node, M);
                        for (int a = 0; a < 3; a++)
                            for (int d = 0; d < 3; d++) {
                                const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, d1 = (d + 1) % 3, d2 = (d + 2) % 3;
                                Ci[a * 3 + d] = M[a1 * 3 + d1] * M[a2 * 3 + d2] - M[a1 * 3 + d2] * M[a2 * 3 + d1];
                            }
                        const double det = M[0] * Ci[0] + M[1] * Ci[1] + M[2] * Ci[2];
                        if (!(std::fabs(det) > 0)) { ok = false; break; }
                        // J = (M^-1)^T = cofactor(M) / det(M)
                        double J[9];
                        for (int q = 0; q < 9; q++) J[q] = Ci[q] / det;
                        for (int a = 0; a < 3; a++) {
                            const double j0 = J[a * 3 + 0], j1 = J[a * 3 + 1], j2 = J[a * 3 + 2];
                            acc[0][a] += j0 / 8;                                   // c100
                            acc[1][a] += j1 / 8;                                   // c010
                            acc[2][a] += j2 / 8;                                   // c001
                            acc[3][a] += (sy * j0 + sx * j1) / 16;                 // c110
                            acc[4][a] += (sz * j0 + sx * j2) / 16;                 // c101
                            acc[5][a] += (sz * j1 + sy * j2) / 16;                 // c011
                            acc[6][a] += (sy * sz * j0 + sx * sz * j1 + sx * sy * j2) / 24;   // c111
                        }
                    }
                    if (!ok) { bad[t] = 1; continue; }
                    for (int q = 0; q < 7; q++) for (int a = 0; a < 3; a++) R.c[q][a] = acc[q][a];
                    R.vol = m->cV[(size_t)e * NP] / (((c->W[0][0] * c->W[1][0]) * c->W[2][0]) / 8);
                    for (int node = 0; node < NP && ok; node++) {
                        const int i = node / (NY * NZ), j = (node / NZ) % NY, k = node % NZ;
                        const double x0 = c->X[0][i], x1 = c->X[1][j], x2 = c->X[2][k];
                        double M[9], J[9];
                        loadM(node, M);
                        for (int a = 0; a < 3; a++) {
                            J[a * 3 + 0] = R.c[0][a] + R.c[3][a] * x1 + R.c[4][a] * x2 + R.c[6][a] * (x1 * x2);
                            J[a * 3 + 1] = R.c[1][a] + R.c[3][a] * x0 + R.c[5][a] * x2 + R.c[6][a] * (x0 * x2);
                            J[a * 3 + 2] = R.c[2][a] + R.c[4][a] * x0 + R.c[5][a] * x1 + R.c[6][a] * (x0 * x1);
                        }
                        for (int a = 0; a < 3; a++)
                            for (int b = 0; b < 3; b++) {
                                const double v = J[a * 3] * M[b * 3] + J[a * 3 + 1] * M[b * 3 + 1] + J[a * 3 + 2] * M[b * 3 + 2];
                                if (!(std::fabs(v - (a == b ? 1.0 : 0.0)) <= 1e-11)) ok = false;
                            }
                        const double cv = R.vol * (((c->W[0][i] * c->W[1][j]) * c->W[2][k]) / 8);
                        const double want = m->cV[(size_t)e * NP + node];
                        if (!(std::fabs(cv - want) <= 1e-13 * std::fabs(want))) ok = false;
                    }
                    if (!ok) { bad[t] = 1; continue; }
                    // parallelepiped (all mixed terms vanish): Jinv*cV = A * (w_i w_j w_k / 8) with one matrix per element
                    if (want_affine) {
                        double scale = 0, mixed = 0;
                        for (int q = 0; q < 3; q++) for (int a = 0; a < 3; a++) scale = std::max(scale, std::fabs(R.c[q][a]));
                        for (int q = 3; q < 7; q++) for (int a = 0; a < 3; a++) mixed = std::max(mixed, std::fabs(R.c[q][a]));
                        if (mixed <= 1e-13 * scale) {
                            double J[9], Cf[9];
                            for (int a = 0; a < 3; a++) for (int d = 0; d < 3; d++) J[a * 3 + d] = R.c[d][a];
                            for (int a = 0; a < 3; a++)
                                for (int d = 0; d < 3; d++) {
                                    const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, d1 = (d + 1) % 3, d2 = (d + 2) % 3;
                                    Cf[a * 3 + d] = J[a1 * 3 + d1] * J[a2 * 3 + d2] - J[a1 * 3 + d2] * J[a2 * 3 + d1];
                                }
                            const double det = J[0] * Cf[0] + J[1] * Cf[1] + J[2] * Cf[2];
                            for (int q = 0; q < 9; q++) R.c[3 + q / 3][q % 3] = Cf[q] * (R.vol / det);
                            for (int a = 0; a < 3; a++) R.c[6][a] = 0.0;
                            R.face[0].meta |= FM_AFFINE;
                        }
                    }
                }
            };
            {
                std::vector<std::thread> th;
                for (unsigned t = 1; t < nth; t++) th.emplace_back(work, t);
                work(0);
                for (auto& x : th) x.join();
            }
            bool all_ok = want_tri;
            for (unsigned t = 0; t < nth; t++) if (bad[t]) all_ok = false;
            c->tri = all_ok;
            CUDA_TRY(c, c->elemRec.upload(er, s));
            CUDA_TRY(c, cudaStreamSynchronize(s));
        }
    }
    CUDA_TRY(c, c->ghostRef.upload(ghostRef, s));
    CUDA_TRY(c, c->bOwner.upload(bOwner, s));
    CUDA_TRY(c, c->bFid.upload(bFid, s));
    CUDA_TRY(c, c->bUnit.upload(bUnit, s));

    // ---- node arrays ----
    const size_t nN = c->nNodes;
    for (int b = 0; b < 2; b++) {
        CUDA_TRY(c, c->rho[b].alloc(nN));
        CUDA_TRY(c, c->T[b].alloc(nN));
        CUDA_TRY(c, c->S[b].alloc(nN));
        CUDA_TRY(c, cudaMemsetAsync(c->S[b].p, 0, nN * 8, s));
        for (int d = 0; d < 3; d++) CUDA_TRY(c, c->U[b][d].alloc(nN));
        CUDA_TRY(c, cudaMemsetAsync(c->rho[b].p, 0, nN * 8, s));
        CUDA_TRY(c, cudaMemsetAsync(c->T[b].p, 0, nN * 8, s));
        for (int d = 0; d < 3; d++) CUDA_TRY(c, cudaMemsetAsync(c->U[b][d].p, 0, nN * 8, s));
    }
    CUDA_TRY(c, c->p.alloc(nN));
    CUDA_TRY(c, cudaMemsetAsync(c->p.p, 0, nN * 8, s));
    for (int q = 0; q < 9; q++) { CUDA_TRY(c, c->GU[q].alloc(nN)); CUDA_TRY(c, cudaMemsetAsync(c->GU[q].p, 0, nN * 8, s)); }
    for (int q = 0; q < 3; q++) { CUDA_TRY(c, c->GT[q].alloc(nN)); CUDA_TRY(c, cudaMemsetAsync(c->GT[q].p, 0, nN * 8, s)); }
    CUDA_TRY(c, c->rho_ref.alloc(nN));
    CUDA_TRY(c, c->p_ref.alloc(nN));
    CUDA_TRY(c, cudaMemsetAsync(c->rho_ref.p, 0, nN * 8, s));
    CUDA_TRY(c, cudaMemsetAsync(c->p_ref.p, 0, nN * 8, s));

    // metrics: cV for all nodes, Jinv (AoS XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX -> row-major arrays) for real nodes
    {
        std::vector<double> h(nN, 1.0);
        for (uint32_t ci = 0; ci < nB; ci++)
            for (int t = 0; t < NP; t++) h[(size_t)ci * NPS + t] = m->cV[(size_t)ci * NP + t];
        for (uint32_t g = 0; g < nG; g++)
            for (int n = 0; n < NPF; n++) {
                const uint32_t ref = ghostRef[(size_t)g * NPF + n];
                if (ref != 0xffffffffu) h[c->ghostBase + (size_t)g * GPS + n] = m->cV[ref];
            }
        CUDA_TRY(c, c->cV.upload(h, s));
        CUDA_TRY(c, cudaStreamSynchronize(s));
        static const int rm[9] = {0, 4, 8, 1, 5, 2, 3, 7, 6};      // AoS component -> row-major a*3+d
        std::vector<double> hj(c->tri ? 0 : (size_t)nB * NPS);
        for (int comp = 0; comp < 9; comp++) {
            if (c->tri) { c->Jinv[rm[comp]].release(); continue; }      // metrics on the fly: Jinv is never streamed
            std::fill(hj.begin(), hj.end(), 0.0);
            for (uint32_t ci = 0; ci < nB; ci++)
                for (int t = 0; t < NP; t++) hj[(size_t)ci * NPS + t] = m->Jinv[((size_t)ci * NP + t) * 9 + comp];
            CUDA_TRY(c, c->Jinv[rm[comp]].upload(hj, s));
            CUDA_TRY(c, cudaStreamSynchronize(s));
        }
    }
    // BC tables default to NONE
    for (int f = 0; f < 4; f++) {
        std::vector<uint8_t> k0(nG, 0);
        std::vector<uint32_t> z(nG, 0);
        CUDA_TRY(c, c->bcKind[f].upload(k0, s));
        CUDA_TRY(c, c->bcRec[f].upload(z, s));
        CUDA_TRY(c, c->bcPeer[f].upload(z, s));
    }
    CUDA_TRY(c, cudaStreamSynchronize(s));
    c->has_sched = false;
    c->have_mesh = true;
    c->have_state = c->have_ref = c->have_bcs = false;
    c->cur = 0;
    return 0;
}

extern "C" int nsem_set_schedule(nsem_ctx* c, const uint32_t* order, uint32_t n) {
    if (!c->have_mesh) { c->err = "nsem_set_schedule: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!order) { c->has_sched = false; return 0; }
    if (n != c->nB) { c->err = "nsem_set_schedule: order must list every real element once"; return 1; }
    std::vector<uint8_t> seen(n, 0);
    for (uint32_t q = 0; q < n; q++) {
        if (order[q] >= n || seen[order[q]]) { c->err = "nsem_set_schedule: not a permutation"; return 1; }
        seen[order[q]] = 1;
    }
    std::vector<uint32_t> h(order, order + n);
    CUDA_TRY(c, c->sched.upload(h, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->has_sched = true;
    return 0;
}

extern "C" int nsem_set_bcs(nsem_ctx* c, const nsem_bc* bcs, uint32_t nb) {
    if (!c->have_mesh) { c->err = "nsem_set_bcs: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const uint32_t nG = c->nG, nB = c->nB;
    const int NPF = c->NPF;
    std::vector<uint8_t> kind[4];
    std::vector<uint32_t> rec[4], peer[4];
    std::vector<double> fixedv[4];
    std::vector<BCRec> recs;
    for (int f = 0; f < 4; f++) { kind[f].assign(nG, 0); rec[f].assign(nG, 0); peer[f].assign(nG, 0); }
    recs.push_back(BCRec{});
    for (uint32_t q = 0; q < nb; q++) {
        const nsem_bc& b = bcs[q];
        if (b.field < 0 || b.field >= 4) { c->err = "nsem_set_bcs: bad field id"; return 1; }
        if (b.kind < NSEM_BC_NEUMANN || b.kind > NSEM_BC_UNLISTED) { c->err = "nsem_set_bcs: unsupported BC kind"; return 1; }
        if (b.kind == NSEM_BC_UNLISTED && b.field != NSEM_F_RHO) { c->err = "nsem_set_bcs: only rho may go without a condition on a patch (NSEM_BC_UNLISTED)"; return 1; }
        const int comps = (b.field == NSEM_F_U) ? 3 : 1;
        BCRec r{};
        for (int d = 0; d < 3; d++) { r.value[d] = b.value[d]; r.tvalue[d] = b.tvalue[d]; }
        r.shape = b.shape; r.tshape = b.tshape; r.zMin = b.zMin;
        recs.push_back(r);
        const uint32_t ri = (uint32_t)recs.size() - 1;
        if ((b.kind == NSEM_BC_FIXED || b.kind == NSEM_BC_UNLISTED) && fixedv[b.field].empty()) fixedv[b.field].assign((size_t)nG * NPF * comps, 0.0);
        for (uint32_t j = 0; j < b.n_faces; j++) {
            const uint32_t face = b.faces[j];
            if (face >= c->nF || c->h_face_neigh[face] < nB) {
                c->err = "nsem_set_bcs: patch face is not a boundary face of this partition";
                return 1;
            }
            const uint32_t g = c->h_face_neigh[face] - nB;
            kind[b.field][g] = (uint8_t)b.kind;
            rec[b.field][g] = ri;
            if (b.kind == NSEM_BC_CYCLIC) {
                if (!b.peer_faces) { c->err = "nsem_set_bcs: CYCLIC needs peer_faces"; return 1; }
                const uint32_t pf = b.peer_faces[j];
                if (pf >= c->nF || c->h_face_neigh[pf] < nB) { c->err = "nsem_set_bcs: CYCLIC peer is not a boundary face"; return 1; }
                peer[b.field][g] = c->h_face_neigh[pf] - nB;
            }
            if (b.kind == NSEM_BC_FIXED || b.kind == NSEM_BC_UNLISTED) {
                if (!b.fixed) { c->err = "nsem_set_bcs: FIXED needs values, UNLISTED the volume ratios"; return 1; }
                for (int n = 0; n < NPF * comps; n++)
                    fixedv[b.field][((size_t)g * NPF) * comps + n] = b.fixed[((size_t)j * NPF) * comps + n];
            }
        }
    }
    // every ghost cell needs a condition for every field (a patch without one keeps stale ghost values in the
    // reference, field.h:2602-2611; the shipped euler cases always define all four)
    for (int f = 0; f < 4; f++)
        for (uint32_t g = 0; g < nG; g++)
            if (kind[f][g] == 0) {
                char bmsg[160];
                snprintf(bmsg, sizeof bmsg, "nsem_set_bcs: boundary face of element %u has no condition for field %d", c->h_bOwner[g], f);
                c->err = bmsg;
                return 1;
            }
    cudaStream_t s = c->stream;
    for (int f = 0; f < 4; f++) {
        CUDA_TRY(c, c->bcKind[f].upload(kind[f], s));
        CUDA_TRY(c, c->bcRec[f].upload(rec[f], s));
        CUDA_TRY(c, c->bcPeer[f].upload(peer[f], s));
        CUDA_TRY(c, c->bcFixed[f].upload(fixedv[f], s));
    }
    CUDA_TRY(c, c->bcRecs.upload(recs, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    c->have_bcs = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// state transfer (reference AoS <-> device SoA)
// ---------------------------------------------------------------------------------------------------------
static int to_device(nsem_ctx* c, const double* host, int comps, double* const dst[3]) {
    const size_t bytes = (size_t)c->nRefNodes * comps * sizeof(double);
    if (c->stage.n < (size_t)c->nRefNodes * 3) CUDA_TRY(c, c->stage.alloc((size_t)c->nRefNodes * 3));
    if (!c->ptrTab.p) { CUDA_TRY(c, c->ptrTab.alloc(3)); CUDA_TRY(c, c->compMap.alloc(3)); }
    const int cm[3] = {0, 1, 2};
    CUDA_TRY(c, cudaMemcpyAsync(c->stage.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->ptrTab.p, dst, comps * sizeof(double*), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->compMap.p, cm, comps * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const uint64_t n = (uint64_t)c->nB * c->NP + (uint64_t)c->nG * c->NPF;
    scatter_to_device<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->stage.p, comps, c->ptrTab.p, c->compMap.p, c->nB, c->NP,
                                                                         c->NPS, c->nG, c->NPF, c->GPS, c->ghostRef.p, c->ghostBase);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}
// Writes the real nodes and every ghost node that a boundary face touches; the remaining nodes of the ghost cells
// (never read by any operator) are returned as 0.
static int from_device(nsem_ctx* c, double* host, int comps, const double* const src[3]) {
    const size_t bytes = (size_t)c->nRefNodes * comps * sizeof(double);
    const size_t realBytes = (size_t)c->nB * c->NP * comps * sizeof(double);
    if (c->stage.n < (size_t)c->nRefNodes * 3) CUDA_TRY(c, c->stage.alloc((size_t)c->nRefNodes * 3));
    if (!c->ptrTab.p) { CUDA_TRY(c, c->ptrTab.alloc(3)); CUDA_TRY(c, c->compMap.alloc(3)); }
    const int cm[3] = {0, 1, 2};
    CUDA_TRY(c, cudaMemsetAsync(reinterpret_cast<char*>(c->stage.p) + realBytes, 0, bytes - realBytes, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->ptrTab.p, src, comps * sizeof(double*), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->compMap.p, cm, comps * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const uint64_t n = (uint64_t)c->nB * c->NP + (uint64_t)c->nG * c->NPF;
    gather_from_device<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->stage.p, comps, (const double* const*)c->ptrTab.p, c->compMap.p,
                                                                          c->nB, c->NP, c->NPS, c->nG, c->NPF, c->GPS, c->ghostRef.p, c->ghostBase);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(host, c->stage.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int nsem_pin_host(nsem_ctx* c, const void* p, uint64_t bytes) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->pin(p, (size_t)bytes)) { c->err = "nsem_pin_host: cudaHostRegister failed"; return 1; }
    return 0;
}

extern "C" int nsem_upload_state(nsem_ctx* c, const double* rho, const double* U, const double* T, const double* p) {
    if (!c->have_mesh) { c->err = "nsem_upload_state: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int k = c->cur;
    double* d1[3] = {c->rho[k].p, nullptr, nullptr};
    if (to_device(c, rho, 1, d1)) return 1;
    double* d3[3] = {c->U[k][0].p, c->U[k][1].p, c->U[k][2].p};
    if (to_device(c, U, 3, d3)) return 1;
    d1[0] = c->T[k].p;
    if (to_device(c, T, 1, d1)) return 1;
    if (p) { d1[0] = c->p.p; if (to_device(c, p, 1, d1)) return 1; }
    c->have_state = true;
    c->speed_valid = false;
    return 0;
}

extern "C" int nsem_download_state(nsem_ctx* c, double* rho, double* U, double* T, double* p) {
    if (!c->have_state) { c->err = "nsem_download_state: no state"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (halo_check(c)) return 1;
    const int k = c->cur;
    const double* s1[3] = {c->rho[k].p, nullptr, nullptr};
    if (rho && from_device(c, rho, 1, s1)) return 1;
    const double* s3[3] = {c->U[k][0].p, c->U[k][1].p, c->U[k][2].p};
    if (U && from_device(c, U, 3, s3)) return 1;
    s1[0] = c->T[k].p;
    if (T && from_device(c, T, 1, s1)) return 1;
    s1[0] = c->p.p;
    if (p && from_device(c, p, 1, s1)) return 1;
    return 0;
}

static int join_comm_fwd(nsem_ctx* c);      // = join_comm (defined with the step): make the compute stream see an in-flight halo exchange

// Operator-level view for unit parity: what gradf<strong>(U) and gradf<strong>(T) + fillBCs(r, fIndex) (field.h:3328-3362,
// 2731-2769) produced in the last step, in the reference's AoS layouts.  The fused sweeps keep these per-unit-volume gradients in HBM between
// sweep A and sweep B, so no extra kernel is needed.
extern "C" int nsem_download_gradients(nsem_ctx* c, double* gradU, double* gradT) {
    if (!c->have_state) { c->err = "nsem_download_gradients: no state"; return 1; }
    if (!(c->prm.diffusion && c->prm.viscosity != 0.0)) { c->err = "nsem_download_gradients: the gradients are only evaluated when diffusion is on"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (join_comm_fwd(c)) return 1;
    if (gradU) {
        // Tensor AoS order XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX (tensor.h:452-454) <- row-major G[a*3+b] = d_a U_b
        const int rm[9] = {0, 4, 8, 1, 5, 2, 3, 7, 6};
        std::vector<double> tmp((size_t)c->nRefNodes * 3);
        for (int g = 0; g < 3; g++) {
            const double* src[3] = {c->GU[rm[3 * g]].p, c->GU[rm[3 * g + 1]].p, c->GU[rm[3 * g + 2]].p};
            if (from_device(c, tmp.data(), 3, src)) return 1;
            for (uint64_t i = 0; i < c->nRefNodes; i++)
                for (int k = 0; k < 3; k++) gradU[i * 9 + 3 * g + k] = tmp[i * 3 + k];
        }
    }
    if (gradT) {
        const double* src[3] = {c->GT[0].p, c->GT[1].p, c->GT[2].p};
        if (from_device(c, gradT, 3, src)) return 1;
    }
    return 0;
}

// ---- pipelined transfers -----------------------------------------------------------------------------------
static int async_setup(nsem_ctx* c) {
    if (c->asyncReady && c->stageIn.n >= (size_t)c->nRefNodes * 6) return 0;
    CUDA_TRY(c, c->stageIn.alloc((size_t)c->nRefNodes * 6));
    CUDA_TRY(c, c->stageOut.alloc((size_t)c->nRefNodes * 6));
    if (!c->ptrTabAsync.p) { CUDA_TRY(c, c->ptrTabAsync.alloc(12)); CUDA_TRY(c, c->compMapAsync.alloc(3)); }
    if (!c->h2d) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&c->evH2D, &c->evScatter, &c->evGather, &c->evD2H}) CUDA_TRY(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    const int cm[3] = {0, 1, 2};
    CUDA_TRY(c, cudaMemcpyAsync(c->compMapAsync.p, cm, sizeof cm, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->asyncReady = true;
    c->scatterPending = c->d2hPending = false;
    return 0;
}

// field q of {rho, U, T, p}: components and offset (in nodes) inside a staging buffer
static const int kAsyncComps[4] = {1, 3, 1, 1};
static const int kAsyncOff[4] = {0, 1, 4, 5};

extern "C" int nsem_upload_state_async(nsem_ctx* c, const double* rho, const double* U, const double* T, const double* p) {
    if (!c->have_mesh) { c->err = "nsem_upload_state_async: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (async_setup(c)) return 1;
    if (join_comm_fwd(c)) return 1;
    const int k = c->cur;
    const double* host[4] = {rho, U, T, p};
    double* dst[4][3] = {{c->rho[k].p, nullptr, nullptr}, {c->U[k][0].p, c->U[k][1].p, c->U[k][2].p}, {c->T[k].p, nullptr, nullptr}, {c->p.p, nullptr, nullptr}};
    // the staging buffer is free once the conversion kernels of the previous upload have read it
    if (c->scatterPending) CUDA_TRY(c, cudaStreamWaitEvent(c->h2d, c->evScatter, 0));
    for (int q = 0; q < 4; q++)
        if (host[q])
            CUDA_TRY(c, cudaMemcpyAsync(c->stageIn.p + (size_t)kAsyncOff[q] * c->nRefNodes, host[q], (size_t)c->nRefNodes * kAsyncComps[q] * sizeof(double),
                                        cudaMemcpyHostToDevice, c->h2d));
    CUDA_TRY(c, cudaEventRecord(c->evH2D, c->h2d));
    // conversion on the compute stream: after everything already enqueued there (a pending download's gather reads the same state arrays)
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->evH2D, 0));
    CUDA_TRY(c, cudaMemcpyAsync(c->ptrTabAsync.p, dst, sizeof dst, cudaMemcpyHostToDevice, c->stream));
    const uint64_t n = (uint64_t)c->nB * c->NP + (uint64_t)c->nG * c->NPF;
    for (int q = 0; q < 4; q++)
        if (host[q]) {
            scatter_to_device<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->stageIn.p + (size_t)kAsyncOff[q] * c->nRefNodes, kAsyncComps[q],
                                                                                 c->ptrTabAsync.p + q * 3, c->compMapAsync.p, c->nB, c->NP, c->NPS, c->nG,
                                                                                 c->NPF, c->GPS, c->ghostRef.p, c->ghostBase);
            c->launches++;
        }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(c->evScatter, c->stream));
    c->scatterPending = true;
    c->have_state = true;
    c->speed_valid = false;
    return 0;
}

extern "C" int nsem_download_state_async(nsem_ctx* c, double* rho, double* U, double* T, double* p) {
    if (!c->have_state) { c->err = "nsem_download_state_async: no state"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (async_setup(c)) return 1;
    if (join_comm_fwd(c)) return 1;
    const int k = c->cur;
    double* host[4] = {rho, U, T, p};
    const double* src[4][3] = {{c->rho[k].p, nullptr, nullptr}, {c->U[k][0].p, c->U[k][1].p, c->U[k][2].p}, {c->T[k].p, nullptr, nullptr}, {c->p.p, nullptr, nullptr}};
    // the outgoing staging buffer is free once the previous download's copies have left it
    if (c->d2hPending) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->evD2H, 0));
    CUDA_TRY(c, cudaMemcpyAsync(c->ptrTabAsync.p, src, sizeof src, cudaMemcpyHostToDevice, c->stream));
    const uint64_t n = (uint64_t)c->nB * c->NP + (uint64_t)c->nG * c->NPF;
    for (int q = 0; q < 4; q++)
        if (host[q]) {
            double* st = c->stageOut.p + (size_t)kAsyncOff[q] * c->nRefNodes;
            const size_t bytes = (size_t)c->nRefNodes * kAsyncComps[q] * sizeof(double), realBytes = (size_t)c->nB * c->NP * kAsyncComps[q] * sizeof(double);
            CUDA_TRY(c, cudaMemsetAsync(reinterpret_cast<char*>(st) + realBytes, 0, bytes - realBytes, c->stream));
            gather_from_device<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(st, kAsyncComps[q], (const double* const*)(c->ptrTabAsync.p + q * 3),
                                                                                  c->compMapAsync.p, c->nB, c->NP, c->NPS, c->nG, c->NPF, c->GPS,
                                                                                  c->ghostRef.p, c->ghostBase);
            c->launches++;
        }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(c->evGather, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->d2h, c->evGather, 0));
    for (int q = 0; q < 4; q++)
        if (host[q])
            CUDA_TRY(c, cudaMemcpyAsync(host[q], c->stageOut.p + (size_t)kAsyncOff[q] * c->nRefNodes, (size_t)c->nRefNodes * kAsyncComps[q] * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->d2h));
    CUDA_TRY(c, cudaEventRecord(c->evD2H, c->d2h));
    c->d2hPending = true;
    return 0;
}

extern "C" int nsem_upload_ref(nsem_ctx* c, const double* rho_ref, const double* p_ref, const double* g) {
    if (!c->have_mesh) { c->err = "nsem_upload_ref: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    double* d1[3] = {c->rho_ref.p, nullptr, nullptr};
    if (to_device(c, rho_ref, 1, d1)) return 1;
    d1[0] = c->p_ref.p;
    if (to_device(c, p_ref, 1, d1)) return 1;
    c->has_gfield = false;
    if (g) {
        for (int d = 0; d < 3; d++) CUDA_TRY(c, c->gfield[d].alloc(c->nNodes));
        double* d3[3] = {c->gfield[0].p, c->gfield[1].p, c->gfield[2].p};
        if (to_device(c, g, 3, d3)) return 1;
        c->has_gfield = true;
    }
    c->have_ref = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------------------
static void fill_kparams(const nsem_ctx* c, KParams& P) {
    std::memset(&P, 0, sizeof P);
    const int k = c->cur, o = 1 - c->cur;
    P.nB = c->nB; P.nG = c->nG; P.ghostBase = c->ghostBase;
    const nsem_params& q = c->prm;
    P.P0 = q.P0; P.T0 = q.T0; P.R = c->convection ? 0.0 : q.cp - q.cv; P.gamma = q.cp / q.cv; P.nu = q.viscosity; P.iPr = 1 / q.Pr; P.dt = q.dt;
    for (int d = 0; d < 3; d++) P.g[d] = q.gravity[d];
    P.mrdt = -1.0 / q.dt; P.mdt = -q.dt;
    P.buoyancy = q.buoyancy;
    // mu = rho*viscosity when diffusion is on (euler.cpp:189-190); viscosity == 0 gives the same fluxes
    P.visc = (q.diffusion && q.viscosity != 0.0) ? 1 : 0;
    P.has_gfield = c->has_gfield ? 1 : 0;
    { const char* pr = std::getenv("NSEM_PROBE"); P.probe = pr ? std::atoi(pr) : 0; }
    // sweep A walks runs of consecutive elements per CTA (RunIter): on a structured mesh the k-face neighbours are then the CTA's own previous
    // / next element and their gathered values hit L2 (sweep A's DRAM reads fall from 21 to 11 KB per element; 1 % of the step under the
    // power cap, profiles/r2_variants.md).  32 per run when every CTA still gets at least four runs, shorter runs on small meshes.
    {
        const char* rn = std::getenv("NSEM_RUN");
        const uint32_t ctas = (uint32_t)c->numSMs * 4u;
        P.run = rn ? (uint32_t)std::max(1, std::atoi(rn)) : std::min(32u, std::max(1u, c->nB / (4u * ctas)));
    }
    std::memcpy(P.D, c->D, sizeof P.D);
    std::memcpy(P.W, c->W, sizeof P.W);
    std::memcpy(P.X, c->X, sizeof P.X);
    P.sms = c->numSMs;
    P.elemRec = c->elemRec.p;
    P.rho_old = c->rho[k].p; P.rho_new = c->rho[o].p;
    P.T_old = c->T[k].p; P.T_new = c->T[o].p;
    P.S_old = c->S[k].p; P.S_new = c->S[o].p;
    for (int d = 0; d < 3; d++) { P.U_old[d] = c->U[k][d].p; P.U_new[d] = c->U[o][d].p; P.gfield[d] = c->gfield[d].p; }
    P.p = c->p.p;
    for (int d = 0; d < 9; d++) { P.GU[d] = c->GU[d].p; P.Jinv[d] = c->Jinv[d].p; }
    for (int d = 0; d < 3; d++) P.GT[d] = c->GT[d].p;
    P.cV = c->cV.p; P.rho_ref = c->rho_ref.p; P.p_ref = c->p_ref.p;
    P.faceOther = c->faceOther.p; P.faceMeta = c->faceMeta.p; P.faceVec = c->faceVec.p; P.faceUnit = c->faceUnit.p;
    P.sched = c->has_sched ? c->sched.p : nullptr;
    P.faceRec = c->faceRec.p;
    P.traceA = c->traceA.p;
    P.mortarA = c->mortarA.p; P.mortarB = c->mortarB.p;
    // stage-slot order of the v4 sweeps (CfgA / CfgB)
    {
        int q = 0;
        P.srcA[q++] = P.rho_old; for (int d = 0; d < 3; d++) P.srcA[q++] = P.U_old[d];
        P.srcA[q++] = P.T_old; P.srcA[q++] = P.p_ref; P.srcA[q++] = P.S_old;
        for (int d = 0; d < 9; d++) P.srcA[q++] = P.Jinv[d];
        P.srcA[q++] = P.cV;
        q = 0;
        P.srcB[q++] = P.rho_old; P.srcB[q++] = P.rho_new; for (int d = 0; d < 3; d++) P.srcB[q++] = P.U_old[d];
        P.srcB[q++] = P.T_old; P.srcB[q++] = P.p; P.srcB[q++] = P.S_old; P.srcB[q++] = P.rho_ref;
        if (P.visc) { for (int d = 0; d < 9; d++) P.srcB[q++] = P.GU[d]; for (int d = 0; d < 3; d++) P.srcB[q++] = P.GT[d]; }
        for (int d = 0; d < 9; d++) P.srcB[q++] = P.Jinv[d];
        P.srcB[q++] = P.cV;
    }
}
static void fill_bcparams(const nsem_ctx* c, const KParams& P, BCParams& B, int phase) {
    std::memset(&B, 0, sizeof B);
    B.nB = c->nB; B.nG = c->nG; B.ghostBase = c->ghostBase;
    B.P0 = P.P0; B.T0 = P.T0; B.R = P.R; B.gamma = P.gamma; B.visc = P.visc; B.phase = phase;
    B.bOwner = c->bOwner.p; B.bFid = c->bFid.p; B.bUnit = c->bUnit.p;
    for (int f = 0; f < 4; f++) { B.kind[f] = c->bcKind[f].p; B.rec[f] = c->bcRec[f].p; B.peer[f] = c->bcPeer[f].p; B.fixedv[f] = c->bcFixed[f].p; }
    B.recs = c->bcRecs.p;
    B.rho_new = P.rho_new; B.rho_old = P.rho_old; B.p = P.p; B.T_old = P.T_old; B.p_ref = P.p_ref;
    for (int d = 0; d < 9; d++) B.GU[d] = P.GU[d];
    for (int d = 0; d < 3; d++) { B.GT[d] = P.GT[d]; B.U_new[d] = P.U_new[d]; }
    B.T_new = P.T_new;
    B.S_new = P.S_new;
}

static int halo_exchange(nsem_ctx* c, double* const* arrays, int nf, cudaStream_t s, int kind, unsigned long long pushed_epoch);

// non-conforming faces: phase 0 before sweep A (old state), phase 1 before sweep B (rho_new, p', gradients of real cells)
static cudaError_t launch_mortar(const nsem_ctx* c, const KParams& P, int phase) {
    if (c->nMortarGroups == 0) return cudaSuccess;
    MortarParams M;
    std::memset(&M, 0, sizeof M);
    M.nGroups = c->nMortarGroups;
    M.NX = c->NX; M.NY = c->NY; M.NZ = c->NZ; M.visc = P.visc;
    M.conv_scheme = P.conv_scheme; M.blend = P.blend;
    M.T0 = P.T0; M.nu = P.nu; M.iPr = P.iPr; M.gammaR = P.gamma * P.R;
    std::memcpy(M.W, c->W, sizeof M.W);
    M.groups = c->mortarGroups.p; M.subs = c->mortarSubs.p;
    M.psiRef = c->psiRef.p; M.psiCor = c->psiCor.p;
    M.rho_old = P.rho_old; M.rho_new = P.rho_new; M.T_old = P.T_old; M.p = P.p;
    for (int d = 0; d < 3; d++) { M.U_old[d] = P.U_old[d]; M.GT[d] = P.GT[d]; }
    for (int d = 0; d < 9; d++) M.GU[d] = P.GU[d];
    M.outA = c->mortarA.p; M.outB = c->mortarB.p;
    if (phase == 0) mortarA_kernel<<<c->nMortarGroups, MORTAR_MAXF, 0, c->stream>>>(M);
    else mortarB_kernel<<<c->nMortarGroups, MORTAR_MAXF, 0, c->stream>>>(M);
    return cudaGetLastError();
}

// S[cur] after the state was set from outside a step (upload, regrid transfer, restart, halo exchange of the initial state)
static int ensure_speed(nsem_ctx* c) {
    if (c->speed_valid) return 0;
    const int k = c->cur;
    const uint64_t n = c->nNodes;
    speed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->prm.T0, c->convection ? 0.0 : (c->prm.cp / c->prm.cv) * (c->prm.cp - c->prm.cv), c->U[k][0].p,
                                                                    c->U[k][1].p, c->U[k][2].p, c->T[k].p, c->S[k].p);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    c->speed_valid = true;
    return 0;
}

// One step with the halo exchanges on the comm stream overlapped with the interior elements (the reference's mul()
// also does interior cells while the halo is in flight, field.h:2381-2415).  Elements that touch an
// inter-partition face go first in each sweep, their traces are packed and exchanged while the interior runs:
//   compute: wait exch(U,T) of the previous step | A(halo) -> evA | A(interior) | bcA | wait exch(A fields) | B(halo) -> evB | B(interior) | bcB
//   comm   :                     wait evA: pack + exchange(rho_new,p,grads) -> evCA      wait evB: pack + exchange(U_new,T_new) -> evCB
static int one_step_overlapped(nsem_ctx* c) {
    KParams P;
    BCParams B;
    fill_kparams(c, P);
    if (ensure_speed(c)) return 1;
    KParams PI = P, PH = P;
    PI.sched = c->schedInt.p; PI.nB = c->nInt;
    PH.sched = c->schedHalo.p; PH.nB = c->nHalo;
    cudaStream_t s = c->stream, cs = c->comm;
    if (c->cbPending) CUDA_TRY(c, cudaStreamWaitEvent(s, c->evCB, 0));
    if (c->nHalo) CUDA_TRY(c, launch_sweepA(c, PH));
    CUDA_TRY(c, cudaEventRecord(c->evA, s));
    CUDA_TRY(c, cudaStreamWaitEvent(cs, c->evA, 0));
    {
        double* arr[14] = {P.rho_new, P.p};
        int nf = 2;
        if (P.visc) { for (int q = 0; q < 9; q++) arr[nf++] = P.GU[q]; for (int q = 0; q < 3; q++) arr[nf++] = P.GT[q]; }
        if (halo_exchange(c, arr, nf, cs, 0, 0)) return 1;
    }
    CUDA_TRY(c, cudaEventRecord(c->evCA, cs));
    if (c->nInt) CUDA_TRY(c, launch_sweepA(c, PI));
    fill_bcparams(c, P, B, 0);
    CUDA_TRY(c, launch_bc(c, B));
    CUDA_TRY(c, cudaStreamWaitEvent(s, c->evCA, 0));
    CUDA_TRY(c, launch_ghost_trace(c, P));
    if (c->nHalo) CUDA_TRY(c, launch_sweepB(c, PH));
    CUDA_TRY(c, cudaEventRecord(c->evB, s));
    CUDA_TRY(c, cudaStreamWaitEvent(cs, c->evB, 0));
    {
        double* arr[5] = {P.U_new[0], P.U_new[1], P.U_new[2], P.T_new, P.S_new};
        if (halo_exchange(c, arr, 5, cs, 1, 0)) return 1;
    }
    CUDA_TRY(c, cudaEventRecord(c->evCB, cs));
    c->cbPending = true;
    if (c->nInt) CUDA_TRY(c, launch_sweepB(c, PI));
    B.phase = 1;
    CUDA_TRY(c, launch_bc(c, B));
    c->launches += (c->nInt ? 2 : 0) + (c->nHalo ? 2 : 0) + (c->nG ? 2 : 0);
    c->cur ^= 1;
    return 0;
}
// make the compute stream see the last in-flight halo exchange (before downloads, timing stops, serial steps)
static int join_comm(nsem_ctx* c) {
    if (c->cbPending) {
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->evCB, 0));
        c->cbPending = false;
    }
    return 0;
}

static int join_comm_fwd(nsem_ctx* c) { return join_comm(c); }

static int one_step(nsem_ctx* c, bool timed, double* acc) {
    if (!timed && !c->peers.empty() && c->overlap && !c->p2p && c->nMortarGroups == 0) return one_step_overlapped(c);
    if (join_comm(c)) return 1;
    KParams P;
    BCParams B;
    fill_kparams(c, P);
    if (ensure_speed(c)) return 1;
    // halo fused into the persistent sweeps: they store the partition-boundary values into the neighbours' windows themselves, boundary
    // elements first, so the transfer runs under the interior elements (the reference's mul() does interior cells while its halo is in
    // flight too, field.h:2381-2415)
    const bool fuse = c->fused && c->use_v4 && !c->has_sched && !c->peers.empty();
    unsigned long long epA = 0, epB = 0;
    if (fuse) {
        epA = ++c->haloEpoch[0];
        epB = ++c->haloEpoch[1];
        // NSEM_HALO_FIRST=0: mesh order (every CTA reports when it runs out of work; the stores still overlap the sweep, the flags come last)
        static const bool halo_first = !(std::getenv("NSEM_HALO_FIRST") && std::strcmp(std::getenv("NSEM_HALO_FIRST"), "0") == 0);
        P.sched = halo_first ? c->schedFused.p : nullptr;
        P.nHalo = halo_first ? c->nHalo : c->nB;
        P.haloGhost = c->haloGhost.p;
    }
    if (timed) cudaEventRecord(c->ev[0], c->stream);
    CUDA_TRY(c, launch_mortar(c, P, 0));
    if (fuse) { P.halo = c->haloFuse.p + (0 * 2 + (int)(epA & 1ull)); P.haloEpoch = epA; }
    CUDA_TRY(c, launch_sweepA(c, P));
    if (timed) cudaEventRecord(c->ev[1], c->stream);
    fill_bcparams(c, P, B, 0);
    CUDA_TRY(c, launch_bc(c, B));
    if (!c->peers.empty()) {
        double* arr[14] = {P.rho_new, P.p};
        int nf = 2;
        if (P.visc) { for (int q = 0; q < 9; q++) arr[nf++] = P.GU[q]; for (int q = 0; q < 3; q++) arr[nf++] = P.GT[q]; }
        if (halo_exchange(c, arr, nf, c->stream, 0, epA)) return 1;
    }
    CUDA_TRY(c, launch_ghost_trace(c, P));
    if (timed) cudaEventRecord(c->ev[2], c->stream);
    CUDA_TRY(c, launch_mortar(c, P, 1));
    if (fuse) { P.halo = c->haloFuse.p + (1 * 2 + (int)(epB & 1ull)); P.haloEpoch = epB; }
    CUDA_TRY(c, launch_sweepB(c, P));
    if (timed) cudaEventRecord(c->ev[3], c->stream);
    B.phase = 1;
    CUDA_TRY(c, launch_bc(c, B));
    if (!c->peers.empty()) {
        double* arr[5] = {P.U_new[0], P.U_new[1], P.U_new[2], P.T_new, P.S_new};
        if (halo_exchange(c, arr, 5, c->stream, 1, epB)) return 1;
    }
    if (timed) cudaEventRecord(c->ev[4], c->stream);
    c->launches += 2 + (c->nG ? 2 : 0) + ((c->nG && (c->use_v4 || c->use_v2)) ? 1 : 0) + (c->nMortarGroups ? 2 : 0);
    c->cur ^= 1;
    if (timed) {
        CUDA_TRY(c, cudaEventSynchronize(c->ev[4]));
        for (int q = 0; q < 4; q++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev[q], c->ev[q + 1]);
            acc[q] += ms;
        }
    }
    return 0;
}

static int check_ready(nsem_ctx* c, const char* who) {
    if (!(c->have_mesh && c->have_params && c->have_state && c->have_ref && (c->have_bcs || c->nG == 0))) {
        c->err = std::string(who) + ": mesh, params, BCs, reference state and state must all be set first";
        return 1;
    }
    return 0;
}

// Launch-bound meshes (the reference's own examples are 10^2..10^4 elements: a step is five to nine launches of a few microseconds each): two
// steps -- one period of the ping-pong buffers -- are captured into a CUDA graph once per call and replayed, so the host issues one graph launch
// per two steps instead of up to eighteen kernel launches.  One partition only (the halo kernels carry the exchange's epoch as an argument).
// NSEM_GRAPH=0 keeps plain launches.
static int steps_by_graph(nsem_ctx* c, int nsteps, int& done) {
    done = 0;
    static const bool on = !(std::getenv("NSEM_GRAPH") && std::strcmp(std::getenv("NSEM_GRAPH"), "0") == 0);
    if (!on || !c->peers.empty() || c->nranks > 1 || nsteps < 8 || c->nB > 32768u) return 0;
    // the first step runs eagerly: lazy set-up (the wave-speed array, kernel attributes) stays outside the capture
    if (one_step(c, false, nullptr)) return 1;
    done = 1;
    const uint64_t l0 = c->launches;
    const int cur0 = c->cur;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    const int rc = one_step(c, false, nullptr) || one_step(c, false, nullptr);
    const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    const uint64_t per_pair = c->launches - l0;
    c->launches = l0;                         // nothing ran yet; the captured pair leaves the buffers where they were
    if (rc || e != cudaSuccess || c->cur != cur0) {
        if (graph) cudaGraphDestroy(graph);
        if (!rc) c->err = std::string("nsem_euler_step: graph capture failed: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return 1;
    }
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        // no graph on this driver/configuration: plain launches do the rest
        cudaGetLastError();
        cudaGraphDestroy(graph);
        return 0;
    }
    int rcl = 0;
    for (; done + 2 <= nsteps; done += 2) {
        if (cudaGraphLaunch(exec, c->stream) != cudaSuccess) { c->err = "nsem_euler_step: cudaGraphLaunch failed"; rcl = 1; break; }
        c->launches += per_pair;
    }
    // the executable graph must outlive its launches
    const cudaError_t es = cudaStreamSynchronize(c->stream);
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    if (es != cudaSuccess) { c->err = std::string("nsem_euler_step: ") + cudaGetErrorString(es); return 1; }
    return rcl;
}

extern "C" int nsem_euler_step(nsem_ctx* c, int nsteps) {
    if (check_ready(c, "nsem_euler_step")) return 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int s = 0;
    if (steps_by_graph(c, nsteps, s)) return 1;
    for (; s < nsteps; s++)
        if (one_step(c, false, nullptr)) return 1;
    return join_comm(c);
}

extern "C" int nsem_time_steps(nsem_ctx* c, int nsteps, double* ms, double* per_kernel_ms) {
    if (check_ready(c, "nsem_time_steps")) return 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    double acc[4] = {0, 0, 0, 0};
    cudaEvent_t t0, t1;
    CUDA_TRY(c, cudaEventCreate(&t0));
    CUDA_TRY(c, cudaEventCreate(&t1));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (per_kernel_ms) {
        // per-kernel events serialise host and device; the total below is taken in a second, untimed-inside pass
        for (int s = 0; s < nsteps; s++)
            if (one_step(c, true, acc)) return 1;
        for (int q = 0; q < 4; q++) per_kernel_ms[q] = acc[q];
    }
    CUDA_TRY(c, cudaEventRecord(t0, c->stream));
    for (int s = 0; s < nsteps; s++)
        if (one_step(c, false, nullptr)) return 1;
    if (join_comm(c)) return 1;
    CUDA_TRY(c, cudaEventRecord(t1, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(t1));
    float tot = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&tot, t0, t1));
    *ms = tot;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// operator-level entry points: the reference's operators one at a time, on the state the context holds, for unit parity
// against the oracle.  They run the same kernels as the step (KParams::op_mode makes a sweep store its divf residual instead of the
// update) and use the `new` halves of the ping-pong buffers as scratch; the current state is not touched.
// ---------------------------------------------------------------------------------------------------------
static int op_begin(nsem_ctx* c, const char* who, KParams& P) {
    if (check_ready(c, who)) return 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (join_comm(c)) return 1;
    fill_kparams(c, P);
    return ensure_speed(c);
}
// sweep A (+ the ghost update and, on several partitions, the halo after it) exactly as the step runs them
static int op_run_first_half(nsem_ctx* c, const KParams& P) {
    BCParams B;
    CUDA_TRY(c, launch_mortar(c, P, 0));
    CUDA_TRY(c, launch_sweepA(c, P));
    fill_bcparams(c, P, B, 0);
    CUDA_TRY(c, launch_bc(c, B));
    c->launches += 1 + (c->nG ? 1 : 0) + (c->nMortarGroups ? 1 : 0);
    if (!c->peers.empty()) {
        double* arr[14] = {P.rho_new, P.p};
        int nf = 2;
        if (P.visc) { for (int q = 0; q < 9; q++) arr[nf++] = P.GU[q]; for (int q = 0; q < 3; q++) arr[nf++] = P.GT[q]; }
        if (halo_exchange(c, arr, nf, c->stream, 0, 0)) return 1;
    }
    return 0;
}

// divf<weak> (field.h:3417-3478: volume term + rusanov faces + div_flux) of the three equations of euler.cpp:195-258 on the current state:
// r_rho = divf(rho U, rho, lambda); r_U = divf(Fc (x) U + I p' - mu grad U, rho_new U, lambda); r_T = divf(Fc theta - mu/Pr grad theta,
// rho_new theta, lambda) -- the residuals BEFORE src/addTemporal/Solve.  Reference layouts (scalar / AoS vector over n_cells_all * NP
// nodes; ghost entries are not meaningful).  Any pointer may be NULL.
extern "C" int nsem_op_divf_weak(nsem_ctx* c, double* r_rho, double* r_U, double* r_T) {
    KParams P;
    if (op_begin(c, "nsem_op_divf_weak", P)) return 1;
    if (r_rho) {
        KParams Q = P;
        Q.op_mode = 1;
        CUDA_TRY(c, launch_mortar(c, Q, 0));
        CUDA_TRY(c, launch_sweepA(c, Q));
        c->launches += 1 + (c->nMortarGroups ? 1 : 0);
        const double* src[3] = {P.rho_new, nullptr, nullptr};
        if (from_device(c, r_rho, 1, src)) return 1;
    }
    if (r_U || r_T) {
        if (op_run_first_half(c, P)) return 1;
        CUDA_TRY(c, launch_ghost_trace(c, P));
        CUDA_TRY(c, launch_mortar(c, P, 1));
        KParams Q = P;
        Q.op_mode = 1;
        CUDA_TRY(c, launch_sweepB(c, Q));
        c->launches += 2 + (c->nMortarGroups ? 1 : 0);
        if (r_U) { const double* src[3] = {P.U_new[0], P.U_new[1], P.U_new[2]}; if (from_device(c, r_U, 3, src)) return 1; }
        if (r_T) { const double* src[3] = {P.T_new, nullptr, nullptr}; if (from_device(c, r_T, 1, src)) return 1; }
    }
    return 0;
}

// rusanov(rho U, rho, lambda) . fN (field.h:2928-2943 contracted as div_flux does, :3093-3114) of the mass equation on every element face
// node, in the face OWNER's frame: flux[(elem*6 + local face) * NPF + slot], slot = the face-node index of dg.cpp:372-404.  Both elements
// of an interior face report the same number (that is what makes the scheme conservative).  Mortar faces report 0 (their flux lives on
// the sub-facets, nsem_mortar.cuh).  
extern "C" int nsem_op_rusanov(nsem_ctx* c, double* flux) {
    KParams P;
    if (op_begin(c, "nsem_op_rusanov", P)) return 1;
    const size_t n = (size_t)c->nB * 6 * c->NPF;
    DevBuf<double> d;
    CUDA_TRY(c, d.alloc(n));
    CUDA_TRY(c, cudaMemsetAsync(d.p, 0, n * sizeof(double), c->stream));
    P.op_mode = 1;
    P.op_flux = d.p;
    // the persistent and the plain-load sweep A report the flux; the bulk-staged generation (NSEM_KERNELS=v2, stored metrics) hands over to
    // the plain-load one
    const bool v2 = c->use_v2;
    if (!c->use_v4) c->use_v2 = false;
    cudaError_t e = launch_mortar(c, P, 0);
    if (e == cudaSuccess) e = launch_sweepA(c, P);
    c->use_v2 = v2;
    CUDA_TRY(c, e);
    c->launches += 1 + (c->nMortarGroups ? 1 : 0);
    CUDA_TRY(c, cudaMemcpyAsync(flux, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// cds(f) (field.h:2881-2893) of a scalar cell field given in the reference layout (n_cells_all * NP values, ghost cells included), same
// output layout as nsem_op_rusanov
extern "C" int nsem_op_cds(nsem_ctx* c, const double* field, double* out) {
    KParams P;
    if (op_begin(c, "nsem_op_cds", P)) return 1;
    double* dst[3] = {P.rho_new, nullptr, nullptr};
    if (to_device(c, field, 1, dst)) return 1;
    const size_t n = (size_t)c->nB * 6 * c->NPF;
    DevBuf<double> d;
    CUDA_TRY(c, d.alloc(n));
    cudaError_t e = cudaErrorInvalidValue;
#define X(a, b, cc) if (c->NX == a && c->NY == b && c->NZ == cc) { op_cds_kernel<a, b, cc><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(P, P.rho_new, d.p); e = cudaGetLastError(); }
    NSEM_ORDERS(X)
#undef X
    CUDA_TRY(c, e);
    c->launches++;
    CUDA_TRY(c, cudaMemcpyAsync(out, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// gradf<strong>(U), gradf<strong>(theta) + fillBCs(r, fIndex) (field.h:3328-3362, 2731-2769) of the CURRENT state (nsem_download_gradients
// returns what the last step left behind)
extern "C" int nsem_op_gradf_strong(nsem_ctx* c, double* grad_U, double* grad_T) {
    KParams P;
    if (op_begin(c, "nsem_op_gradf_strong", P)) return 1;
    if (!P.visc) { c->err = "nsem_op_gradf_strong: the gradients are only evaluated when diffusion is on"; return 1; }
    if (op_run_first_half(c, P)) return 1;
    return nsem_download_gradients(c, grad_U, grad_T);
}

// applyExplicitBCs (field.h:2586-2727) on one field given in the reference layout: the ghost entries of `values` are replaced by what
// the field's boundary conditions make of the owner values.  field = NSEM_F_RHO, NSEM_F_U (3 components) or NSEM_F_T (the condition acts
// on theta = T + T0 and T0 is subtracted again, euler.cpp:258,286); NSEM_F_P is tied to the equation of state inside the step and has
// no stand-alone form.
extern "C" int nsem_op_apply_bcs(nsem_ctx* c, int field, double* values) {
    KParams P;
    if (op_begin(c, "nsem_op_apply_bcs", P)) return 1;
    if (!(field == NSEM_F_RHO || field == NSEM_F_U || field == NSEM_F_T)) { c->err = "nsem_op_apply_bcs: field must be NSEM_F_RHO, NSEM_F_U or NSEM_F_T"; return 1; }
    BCParams B;
    fill_bcparams(c, P, B, field == NSEM_F_RHO ? 0 : 1);
    B.visc = 0;
    B.S_new = nullptr;
    double* d3[3] = {P.U_new[0], P.U_new[1], P.U_new[2]};
    double* d1[3] = {field == NSEM_F_RHO ? P.rho_new : P.T_new, nullptr, nullptr};
    const int comps = field == NSEM_F_U ? 3 : 1;
    if (to_device(c, values, comps, field == NSEM_F_U ? d3 : d1)) return 1;
    CUDA_TRY(c, launch_bc(c, B));
    if (c->nG) c->launches++;
    return from_device(c, values, comps, field == NSEM_F_U ? d3 : d1);
}

// ---------------------------------------------------------------------------------------------------------
// explicit scalar advection: apps/convection/convection.cpp:113-137 on the kernels of the euler path.
//     M = divf(Fc * T, false, &F, &T, &lambdaMax); addTemporal<1>(M); Solve(M)      with Fc = U, lambdaMax = cds(mag(U)) / 2
// is the mass equation of the euler step with the transported scalar in the place of rho and no sound speed in lambdaMax, so the scalar
// is kept in the rho slot of the state (nsem_upload_state(ctx, scalar, U, NULL...), boundary conditions under NSEM_F_RHO) and a step is
// sweep A (inviscid instantiation, R = 0) + the ghost update of rho + the halo; the other outputs of the sweep are not used.
// problem_init: 0 = the wind is the uploaded U; 1 = LEVEQUE, 2 / 3 = LAURITZEN_0 / LAURITZEN_1 (spherical meshes, nsem_set_sphere), re-evaluated
// on the device at time step * dt before every step (convection.cpp:55-82,114-121; needs nsem_upload_coords).  etime = end_step * dt; first_step = the step the next call starts with.
// ---------------------------------------------------------------------------------------------------------
extern "C" int nsem_set_convection(nsem_ctx* c, int problem_init, double etime, long first_step) {
    if (problem_init < 0 || problem_init > 3) { c->err = "nsem_set_convection: problem_init must be 0 (NONE), 1 (LEVEQUE), 2 (LAURITZEN_0) or 3 (LAURITZEN_1)"; return 1; }
    if (problem_init >= 1 && !c->xyz[0].p) { c->err = "nsem_set_convection: the analytic winds need the node coordinates (nsem_upload_coords)"; return 1; }
    if (problem_init >= 2 && !(c->sphere_radius > 0)) { c->err = "nsem_set_convection: the Lauritzen winds need the radius of the sphere (nsem_set_sphere)"; return 1; }
    c->convection = true;
    c->conv_init = problem_init;
    c->conv_etime = etime;
    c->conv_step = first_step - 1;
    c->speed_valid = false;
    return 0;
}

// Controls::convection_scheme for the scalar's face value (divf, field.h:3427-3437): 0 RUSANOV (default), 1 CDS, 2 UDS, 3 BLENDED with
// Controls::blend_factor.  examples/transport/wave2d runs BLENDED 0.6.
extern "C" int nsem_set_convection_scheme(nsem_ctx* c, int scheme, double blend_factor) {
    if (scheme < 0 || scheme > 3) { c->err = "nsem_set_convection_scheme: 0 RUSANOV, 1 CDS, 2 UDS or 3 BLENDED"; return 1; }
    c->conv_scheme = scheme;
    c->conv_blend = blend_factor;
    return 0;
}

// Controls::time_scheme AB1..AB5 for nsem_convection_step (order 1 = the one forward-Euler stage that BDF1, AB1 and RK1-RK4 all are on this
// path); the residual history starts over (a new field after a regrid or a restart, field.h:3885-3895)
extern "C" int nsem_set_ab_order(nsem_ctx* c, int order) {
    if (order < 1 || order > 5) { c->err = "nsem_set_ab_order: order must be 1..5"; return 1; }
    c->ab_order = order;
    c->ab_stored = 0;
    c->ab_head = 0;
    return 0;
}

// Mesh::sphere_radius of a cubed-sphere mesh (mesh.cpp:32): only the Lauritzen winds read it, everything else of a spherical case is in
// the uploaded geometry and the per-node gravity
extern "C" int nsem_set_sphere(nsem_ctx* c, double radius) {
    if (!(radius > 0)) { c->err = "nsem_set_sphere: radius must be positive"; return 1; }
    c->sphere_radius = radius;
    return 0;
}

// Mesh::cC (node coordinates, AoS, n_cells_all * NP entries, ghost cells included)
extern "C" int nsem_upload_coords(nsem_ctx* c, const double* cC) {
    if (!c->have_mesh) { c->err = "nsem_upload_coords: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    for (int d = 0; d < 3; d++) CUDA_TRY(c, c->xyz[d].alloc(c->nNodes));
    double* dst[3] = {c->xyz[0].p, c->xyz[1].p, c->xyz[2].p};
    return to_device(c, cC, 3, dst);
}

extern "C" int nsem_convection_step(nsem_ctx* c, int nsteps) {
    if (!c->convection) { c->err = "nsem_convection_step: call nsem_set_convection first"; return 1; }
    if (check_ready(c, "nsem_convection_step")) return 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (join_comm(c)) return 1;
    for (int st = 0; st < nsteps; st++) {
        const int k = c->cur;
        c->conv_step++;
        if (c->conv_init == 1) {
            const uint64_t n = c->nNodes;
            wind_leveque_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, (double)c->conv_step * c->prm.dt, c->conv_etime, c->xyz[0].p, c->xyz[1].p,
                                                                                   c->U[k][0].p, c->U[k][1].p, c->U[k][2].p);
            c->launches++;
            c->speed_valid = false;
        } else if (c->conv_init >= 2) {
            const uint64_t n = c->nNodes;
            wind_lauritzen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->conv_init - 2, (double)c->conv_step * c->prm.dt, c->conv_etime,
                                                                                     c->sphere_radius, c->xyz[0].p, c->xyz[1].p, c->xyz[2].p, c->U[k][0].p,
                                                                                     c->U[k][1].p, c->U[k][2].p);
            c->launches++;
            c->speed_valid = false;
        }
        if (ensure_speed(c)) return 1;                 // S = |U| (R = 0)
        KParams P;
        BCParams B;
        fill_kparams(c, P);
        P.visc = 0; P.buoyancy = 0;
        P.conv_scheme = c->conv_scheme; P.blend = c->conv_blend;
        const bool plain = (c->conv_scheme != 0);          // the other face values live in the plain-load sweep only
        const bool keep_v2 = c->use_v2, keep_v4 = c->use_v4;
        if (plain) { c->use_v2 = false; c->use_v4 = false; }
        struct Restore { nsem_ctx* c; bool v2, v4; ~Restore() { c->use_v2 = v2; c->use_v4 = v4; } } restore{c, keep_v2, keep_v4};
        CUDA_TRY(c, launch_mortar(c, P, 0));
        if (c->ab_order > 1) {
            // AB2..AB5: the sweep leaves the residual, the update combines it with the stored ones (field.h:3789-3806, 3885-3905)
            const bool v2 = c->use_v2;
            if (!c->use_v4) c->use_v2 = false;         // the bulk-staged generation has no residual mode; the plain-load kernels do
            P.op_mode = 1;
            const cudaError_t e = launch_sweepA(c, P);
            c->use_v2 = v2;
            CUDA_TRY(c, e);
            ABParams A;
            std::memset(&A, 0, sizeof A);
            const bool first = (c->ab_stored == 0);
            for (int j = 0; j < c->ab_order; j++)
                if (c->abHist[j].n < c->nNodes) { CUDA_TRY(c, c->abHist[j].alloc(c->nNodes)); }
            if (!first) c->ab_head = (c->ab_head + c->ab_order - 1) % c->ab_order;      // updateStore: the oldest entry's buffer takes PREV(0)
            c->ab_stored = first ? 1 : c->ab_stored + 1;
            A.nB = c->nB; A.NP = c->NP; A.NPS = c->NPS; A.order = c->ab_order; A.use = std::min(c->ab_order, c->ab_stored); A.first = first ? 1 : 0;
            A.dt = c->prm.dt; A.cV = c->cV.p; A.q_old = P.rho_old; A.r = P.rho_new;
            for (int j = 0; j < c->ab_order; j++) A.prev[j] = c->abHist[(c->ab_head + j) % c->ab_order].p;
            const uint64_t nn = (uint64_t)c->nB * c->NPS;
            ab_update_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, c->stream>>>(A);
            c->launches++;
        } else
        CUDA_TRY(c, launch_sweepA(c, P));
        fill_bcparams(c, P, B, 0);
        CUDA_TRY(c, launch_bc(c, B));
        c->launches += 1 + (c->nG ? 1 : 0) + (c->nMortarGroups ? 1 : 0);
        if (!c->peers.empty()) {
            double* arr[2] = {P.rho_new, P.p};
            if (halo_exchange(c, arr, 2, c->stream, 0, 0)) return 1;
        }
        // the new scalar becomes the current one; U, T and S stay where they are
        std::swap(c->rho[0].p, c->rho[1].p);
        std::swap(c->rho[0].n, c->rho[1].n);
    }
    return 0;
}

// ASYNC_COMM (field.h:2255-2324) on the current state: the halo of rho, U, T, p and the reference state (= nsem_exchange_state_halos)
extern "C" int nsem_op_halo(nsem_ctx* c) { return nsem_exchange_state_halos(c); }

// Map the neighbours' receive windows (nsem_halo.cuh).  Collective over the ranks that share faces: the IPC handles and the slot offsets
// travel once, through NCCL point-to-point; if ANY rank cannot map a neighbour (no peer access, handles refused) every rank keeps the NCCL
// transport -- the choice is agreed with an all-reduce so that no rank waits for a flag nobody will set.  NSEM_HALO=nccl forces NCCL.
static int halo_p2p_setup(nsem_ctx* c) {
    halo_p2p_release(c);
    for (auto& e : c->haloEpoch) e = 0;
#ifdef NSEM_WITH_NCCL
    const char* mode = std::getenv("NSEM_HALO");
    int want = !(mode && std::strcmp(mode, "nccl") == 0) && (int)c->peers.size() <= HALO_MAX_PEERS;
    const int np = (int)c->peers.size();
    c->nRecvSlots = c->nSendSlots;             // the same faces in both directions
    struct Msg { cudaIpcMemHandle_t h; uint64_t nRecv, offForYou, raw; uint32_t yourIdx, pid; };
    std::vector<Msg> out(np), in(np);
    if (want) {
        const size_t bytes = HALO_HEADER_BYTES + halo_window_doubles(c->nRecvSlots) * sizeof(double);
        if (cudaMalloc(&c->winBase, bytes) != cudaSuccess) { cudaGetLastError(); c->winBase = nullptr; want = 0; }
        else CUDA_TRY(c, cudaMemsetAsync(c->winBase, 0, HALO_HEADER_BYTES, c->stream));
    }
    cudaIpcMemHandle_t h;
    std::memset(&h, 0, sizeof h);
    if (want && cudaIpcGetMemHandle(&h, c->winBase) != cudaSuccess) { cudaGetLastError(); want = 0; }
    for (int p = 0; p < np; p++) out[p] = Msg{h, c->nRecvSlots, c->peers[p].off, (uint64_t)(uintptr_t)c->winBase, (uint32_t)p, (uint32_t)getpid()};
    DevBuf<unsigned char> dOut, dIn;
    DevBuf<int> dFlag;
    CUDA_TRY(c, dOut.alloc((size_t)np * sizeof(Msg)));
    CUDA_TRY(c, dIn.alloc((size_t)np * sizeof(Msg)));
    CUDA_TRY(c, dFlag.alloc(1));
    CUDA_TRY(c, cudaMemcpyAsync(dOut.p, out.data(), (size_t)np * sizeof(Msg), cudaMemcpyHostToDevice, c->stream));
    ncclResult_t r = g_nccl.GroupStart();
    for (int p = 0; p < np && r == ncclSuccess; p++) {
        r = g_nccl.Send(dOut.p + (size_t)p * sizeof(Msg), sizeof(Msg), ncclChar, c->peers[p].rank, c->nccl, c->stream);
        if (r == ncclSuccess) r = g_nccl.Recv(dIn.p + (size_t)p * sizeof(Msg), sizeof(Msg), ncclChar, c->peers[p].rank, c->nccl, c->stream);
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) { c->err = std::string("halo set-up: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2); return 1; }
    CUDA_TRY(c, cudaMemcpyAsync(in.data(), dIn.p, (size_t)np * sizeof(Msg), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    int ok = want;
    c->remotes.assign(np, nsem_ctx::Remote{});
    for (int p = 0; p < np && ok; p++) {
        nsem_ctx::Remote& rm = c->remotes[p];
        rm.nRecv = in[p].nRecv; rm.offForMe = in[p].offForYou; rm.myIdx = in[p].yourIdx;
        if (in[p].raw == 0) { ok = 0; break; }
        if (in[p].pid == (uint32_t)getpid()) rm.base = reinterpret_cast<void*>((uintptr_t)in[p].raw);      // same address space
        else if (cudaIpcOpenMemHandle(&rm.base, in[p].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) rm.ipc = true;
        else { cudaGetLastError(); rm.base = nullptr; ok = 0; }
    }
    // every rank of the communicator takes part (ranks without neighbours included): the transport is one decision
    CUDA_TRY(c, cudaMemcpyAsync(dFlag.p, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (g_nccl.AllReduce(dFlag.p, dFlag.p, 1, ncclInt, ncclMin, c->nccl, c->stream) != ncclSuccess) { c->err = "halo set-up: ncclAllReduce failed"; return 1; }
    CUDA_TRY(c, cudaMemcpyAsync(&ok, dFlag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (ok) c->p2p = !c->peers.empty();
    else halo_p2p_release(c);
    if (c->p2p) {
        // tables of the fused variant (the v4 sweeps store into the windows themselves; opt-in, see below)
        const char* fz = std::getenv("NSEM_HALO_FUSED");
        std::vector<HaloFuse> hf(4);
        char* base = static_cast<char*>(c->winBase);
        for (int kind = 0; kind < 2; kind++)
            for (int parity = 0; parity < 2; parity++) {
                HaloFuse& H = hf[kind * 2 + parity];
                std::memset(&H, 0, sizeof H);
                H.npeers = np;
                H.counter = reinterpret_cast<unsigned int*>(base + HALO_KINDS * HALO_MAX_PEERS * 8);
                for (int p = 0; p < np; p++) {
                    const nsem_ctx::Remote& rm = c->remotes[p];
                    char* rb = static_cast<char*>(rm.base);
                    H.win[p] = reinterpret_cast<double*>(rb + HALO_HEADER_BYTES) + halo_region_offset(kind, parity, rm.nRecv) + rm.offForMe;
                    H.stride[p] = rm.nRecv;
                    H.flag[p] = reinterpret_cast<unsigned long long*>(rb) + (size_t)kind * HALO_MAX_PEERS + rm.myIdx;
                }
            }
        std::vector<uint32_t> hg((size_t)c->nG * 2, 0xffffffffu);
        for (int p = 0; p < np; p++)
            for (uint32_t j = 0; j < c->peers[p].nf; j++) {
                hg[(size_t)(c->peers[p].g0 + j) * 2] = (uint32_t)p;
                hg[(size_t)(c->peers[p].g0 + j) * 2 + 1] = j * (uint32_t)c->GPS;
            }
        std::vector<uint32_t> order;
        {
            std::vector<uint8_t> isHalo(c->nB, 0);
            for (const auto& P : c->peers)
                for (uint32_t j = 0; j < P.nf; j++) isHalo[c->h_bOwner[P.g0 + j]] = 1;
            for (uint32_t e = 0; e < c->nB; e++) if (isHalo[e]) order.push_back(e);
            for (uint32_t e = 0; e < c->nB; e++) if (!isHalo[e]) order.push_back(e);
        }
        CUDA_TRY(c, c->haloFuse.upload(hf, c->stream));
        CUDA_TRY(c, c->haloGhost.upload(hg, c->stream));
        CUDA_TRY(c, c->schedFused.upload(order, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        // measured on 2 B200 (profiles/r2k_halo_fused_vs_push_2gpu.log): the fused variant is bit-identical but 5-8 % SLOWER than the two
        // small kernels (the boundary-first schedule costs the sweeps their streaming order, the strided remote stores ride on their
        // critical path, and the transfer it hides is only 0.06 ms), so it is opt-in: NSEM_HALO_FUSED=1
        c->fused = (fz && std::strcmp(fz, "1") == 0);
    }
#endif
    return 0;
}

// a neighbour that never arrived (halo_pull_kernel gave up): reported where the host synchronises anyway
static int halo_check(nsem_ctx* c) {
    if (!c->p2p) return 0;
    int e = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&e, static_cast<char*>(c->winBase) + HALO_KINDS * HALO_MAX_PEERS * 8 + 8, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (e) { c->err = "halo exchange: neighbour #" + std::to_string(e - 1) + " (rank " + std::to_string(c->peers[e - 1].rank) + ") never delivered its face traces"; return 1; }
    return 0;
}

// Diagnostics of the peer-memory halo: the time (ms) this rank's compute stream has spent, since the last call, inside halo_pull_kernel
// waiting for its neighbours' flags -- [0] after sweep A, [1] after sweep B, [2] state exchanges.  It is the skew between partitions
// (a neighbour that is slower, or started later), not transfer time.  Zeros when the transport is NCCL.
extern "C" int nsem_halo_wait_ms(nsem_ctx* c, double out[3]) {
    out[0] = out[1] = out[2] = 0.0;
    if (!c->p2p) return 0;
    CUDA_TRY(c, cudaSetDevice(c->device));
    unsigned long long ns[3] = {0, 0, 0};
    char* w = static_cast<char*>(c->winBase) + HALO_KINDS * HALO_MAX_PEERS * 8 + 16;
    CUDA_TRY(c, cudaMemcpyAsync(ns, w, sizeof ns, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(w, 0, sizeof ns, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; k++) out[k] = (double)ns[k] * 1e-6;
    return 0;
}

extern "C" int nsem_set_halo(nsem_ctx* c, const nsem_halo_peer* peers, uint32_t n_peers) {
    if (!c->have_mesh) { c->err = "nsem_set_halo: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->peers.clear();
    c->nSendSlots = 0;
    if (n_peers == 0) return c->nranks > 1 ? halo_p2p_setup(c) : 0;      // the transport is agreed by all ranks, neighbours or not
    if (c->nranks <= 1) { c->err = "nsem_set_halo: context was created for a single rank"; return 1; }
    const int NPF = c->NPF, GPS = c->GPS, NPS = c->NPS;
    std::vector<uint32_t> nodes;
    for (uint32_t q = 0; q < n_peers; q++) {
        const nsem_halo_peer& hp = peers[q];
        if (hp.peer_rank < 0 || hp.peer_rank >= c->nranks || hp.peer_rank == c->rank || hp.n_faces == 0) {
            c->err = "nsem_set_halo: bad peer entry";
            return 1;
        }
        nsem_ctx::Peer P{hp.peer_rank, 0, hp.n_faces, (uint64_t)nodes.size()};
        for (uint32_t j = 0; j < hp.n_faces; j++) {
            const uint32_t face = hp.faces[j];
            if (face >= c->nF || c->h_face_neigh[face] < c->nB) { c->err = "nsem_set_halo: face is not a boundary face"; return 1; }
            const uint32_t g = c->h_face_neigh[face] - c->nB;
            if (j == 0) P.g0 = g;
            else if (g != P.g0 + j) {
                c->err = "nsem_set_halo: ghost cells of an interMesh patch are not contiguous (addBoundaryCells creates them patch by patch)";
                return 1;
            }
            const int fid = c->h_bFid[g];
            const int NX = c->NX, NY = c->NY, NZ = c->NZ;
            for (int n = 0; n < GPS; n++) {
                uint32_t nd = 0xffffffffu;
                int a, b, cnt;
                if (fid < 2) { a = n / NY; b = n % NY; cnt = NX * NY; }
                else if (fid < 4) { a = n / NZ; b = n % NZ; cnt = NX * NZ; }
                else { a = n / NZ; b = n % NZ; cnt = NY * NZ; }
                if (n < NPF && n < cnt) nd = c->h_bOwner[g] * (uint32_t)NPS + (uint32_t)h_face_node(c, fid, a, b);
                nodes.push_back(nd);
            }
        }
        c->peers.push_back(P);
    }
    c->nSendSlots = nodes.size();
    CUDA_TRY(c, c->sendNodes.upload(nodes, c->stream));
    CUDA_TRY(c, c->sendBuf.alloc((size_t)16 * nodes.size()));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (!c->evA)
        for (cudaEvent_t* e : {&c->evA, &c->evCA, &c->evB, &c->evCB}) CUDA_TRY(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    {
        std::vector<uint8_t> isHalo(c->nB, 0);
        for (const auto& P : c->peers)
            for (uint32_t j = 0; j < P.nf; j++) isHalo[c->h_bOwner[P.g0 + j]] = 1;
        std::vector<uint32_t> li, lh;
        for (uint32_t e = 0; e < c->nB; e++) (isHalo[e] ? lh : li).push_back(e);
        c->nInt = (uint32_t)li.size();
        c->nHalo = (uint32_t)lh.size();
        CUDA_TRY(c, c->schedInt.upload(li, c->stream));
        CUDA_TRY(c, c->schedHalo.upload(lh, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    const char* ov = std::getenv("NSEM_OVERLAP");
    c->overlap = (ov && std::strcmp(ov, "1") == 0);
    c->cbPending = false;
    return halo_p2p_setup(c);
}

// pack the owner-side face values of `nf` arrays and exchange them with every peer on stream `s`:
// send from the packed buffer, receive straight into the ghost region [ghostBase + g0*GPS, +nf*GPS) of each array
static int halo_exchange(nsem_ctx* c, double* const* arrays, int nf, cudaStream_t s, int kind, unsigned long long pushed_epoch) {
    if (c->peers.empty()) return 0;
    if (nf > 16 || nf > halo_kind_fields(kind)) { c->err = "halo_exchange: too many fields"; return 1; }
    if (c->p2p) {
        // two kernels: stores into the neighbours' windows over NVLink + flags, then wait for the neighbours' flags and fill the ghost cells
        // (pushed_epoch != 0: the sweep that produced the arrays has already stored and flagged them, only the second kernel runs)
        const int np = (int)c->peers.size();
        const unsigned long long epoch = pushed_epoch ? pushed_epoch : ++c->haloEpoch[kind];
        const int parity = (int)(epoch & 1ull);
        char* base = static_cast<char*>(c->winBase);
        HaloPushParams H;
        std::memset(&H, 0, sizeof H);
        H.nfields = nf; H.npeers = np; H.nslots = c->nSendSlots; H.node = c->sendNodes.p; H.epoch = epoch;
        H.counter = reinterpret_cast<unsigned int*>(base + HALO_KINDS * HALO_MAX_PEERS * 8);
        for (int f = 0; f < nf; f++) H.src[f] = arrays[f];
        HaloPullParams R;
        std::memset(&R, 0, sizeof R);
        R.nfields = nf; R.npeers = np; R.nslots = c->nRecvSlots; R.stride = c->nRecvSlots; R.ghostBase = c->ghostBase; R.epoch = epoch;
        R.win = reinterpret_cast<const double*>(base + HALO_HEADER_BYTES) + halo_region_offset(kind, parity, c->nRecvSlots);
        R.flag = reinterpret_cast<const unsigned long long*>(base) + (size_t)kind * HALO_MAX_PEERS;
        R.error = reinterpret_cast<int*>(base + HALO_KINDS * HALO_MAX_PEERS * 8 + 8);
        R.wait_ns = reinterpret_cast<unsigned long long*>(base + HALO_KINDS * HALO_MAX_PEERS * 8 + 16) + kind;
        { const char* t = std::getenv("NSEM_HALO_TIMEOUT_S"); R.timeout_ns = (unsigned long long)((t ? std::atof(t) : 30.0) * 1e9); }
        for (int f = 0; f < nf; f++) R.dst[f] = arrays[f];
        for (int p = 0; p < np; p++) {
            const nsem_ctx::Remote& rm = c->remotes[p];
            char* rb = static_cast<char*>(rm.base);
            H.off[p] = R.off[p] = c->peers[p].off;
            H.win[p] = reinterpret_cast<double*>(rb + HALO_HEADER_BYTES) + halo_region_offset(kind, parity, rm.nRecv) + rm.offForMe;
            H.stride[p] = rm.nRecv;
            H.flag[p] = reinterpret_cast<unsigned long long*>(rb) + (size_t)kind * HALO_MAX_PEERS + rm.myIdx;
            R.ghostOff[p] = (uint64_t)c->peers[p].g0 * c->GPS;
        }
        H.off[np] = R.off[np] = c->nSendSlots;
        if (!pushed_epoch) { halo_push_kernel<<<(unsigned)((H.nslots + 255) / 256), 256, 0, s>>>(H); c->launches++; }
        halo_pull_kernel<<<(unsigned)((R.nslots + 255) / 256), 256, 0, s>>>(R);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
        return 0;
    }
#ifdef NSEM_WITH_NCCL
    PackParams H;
    std::memset(&H, 0, sizeof H);
    H.nfields = nf; H.nslots = c->nSendSlots; H.node = c->sendNodes.p; H.dst = c->sendBuf.p;
    for (int f = 0; f < nf; f++) H.src[f] = arrays[f];
    halo_pack_kernel<<<(unsigned)((H.nslots + 255) / 256), 256, 0, s>>>(H);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    ncclResult_t r = g_nccl.GroupStart();
    for (const auto& P : c->peers) {
        const size_t cnt = (size_t)P.nf * c->GPS;
        for (int f = 0; f < nf && r == ncclSuccess; f++) {
            r = g_nccl.Send(c->sendBuf.p + (size_t)f * c->nSendSlots + P.off, cnt, ncclDouble, P.rank, c->nccl, s);
            if (r == ncclSuccess) r = g_nccl.Recv(arrays[f] + c->ghostBase + (size_t)P.g0 * c->GPS, cnt, ncclDouble, P.rank, c->nccl, s);
        }
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) {
        c->err = std::string("NCCL halo exchange: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2);
        return 1;
    }
    return 0;
#else
    (void)arrays; (void)nf; (void)s;
    c->err = "library built without NCCL";
    return 1;
#endif
}

extern "C" int nsem_exchange_state_halos(nsem_ctx* c) {
    if (!(c->have_mesh && c->have_state && c->have_ref)) { c->err = "nsem_exchange_state_halos: upload state and reference first"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int k = c->cur;
    double* arr[8] = {c->rho[k].p, c->U[k][0].p, c->U[k][1].p, c->U[k][2].p, c->T[k].p, c->p.p, c->rho_ref.p, c->p_ref.p};
    if (halo_exchange(c, arr, 8, c->stream, 2, 0)) return 1;
    c->speed_valid = false;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int nsem_upload_geopotential(nsem_ctx* c, const double* gh) {
    if (!c->have_mesh) { c->err = "nsem_upload_geopotential: no mesh"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, c->gh.alloc(c->nNodes));
    double* d1[3] = {c->gh.p, nullptr, nullptr};
    if (to_device(c, gh, 1, d1)) return 1;
    c->has_gh = true;
    return 0;
}

extern "C" int nsem_diagnostics(nsem_ctx* c, double out[6]) {
    if (check_ready(c, "nsem_diagnostics")) return 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (join_comm(c)) return 1;
    const int nblk = 148 * 4;
    if (c->diagPartial.n < (size_t)(nblk + 1) * 6) CUDA_TRY(c, c->diagPartial.alloc((size_t)(nblk + 1) * 6));
    DiagParams D;
    std::memset(&D, 0, sizeof D);
    const int k = c->cur;
    D.nB = c->nB; D.NP = c->NP; D.NPS = c->NPS;
    D.P0 = c->prm.P0; D.T0 = c->prm.T0; D.R = c->prm.cp - c->prm.cv; D.cp = c->prm.cp; D.cv = c->prm.cv; D.dt = c->prm.dt;
    D.rho = c->rho[k].p; D.T = c->T[k].p; D.p = c->p.p; D.p_ref = c->p_ref.p; D.cV = c->cV.p;
    for (int d = 0; d < 3; d++) D.U[d] = c->U[k][d].p;
    D.gh = c->has_gh ? c->gh.p : nullptr;
    D.partial = c->diagPartial.p; D.nparts = nblk;
    diag_kernel<<<nblk, 256, 0, c->stream>>>(D);
    diag_fold_kernel<<<1, 256, 0, c->stream>>>(D);
    c->launches += 2;
    CUDA_TRY(c, cudaGetLastError());
    double* res = c->diagPartial.p + (size_t)nblk * 6;      // {max, min, sum courant, mass, energy, volume}
    double count = (double)c->nB * c->NP;
#ifdef NSEM_WITH_NCCL
    if (c->nranks > 1) {
        // reduce_max / reduce_min / reduce_sum over ranks (MP::allreduce, mp.h:93-104)
        if (c->stage.n < 8) CUDA_TRY(c, c->stage.alloc(8));
        CUDA_TRY(c, cudaMemcpyAsync(c->stage.p, &count, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        ncclResult_t r = g_nccl.GroupStart();
        if (r == ncclSuccess) r = g_nccl.AllReduce(res + 0, res + 0, 1, ncclDouble, ncclMax, c->nccl, c->stream);
        if (r == ncclSuccess) r = g_nccl.AllReduce(res + 1, res + 1, 1, ncclDouble, ncclMin, c->nccl, c->stream);
        if (r == ncclSuccess) r = g_nccl.AllReduce(res + 2, res + 2, 4, ncclDouble, ncclSum, c->nccl, c->stream);
        if (r == ncclSuccess) r = g_nccl.AllReduce(c->stage.p, c->stage.p, 1, ncclDouble, ncclSum, c->nccl, c->stream);
        ncclResult_t r2 = g_nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) { c->err = std::string("nsem_diagnostics: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2); return 1; }
        CUDA_TRY(c, cudaMemcpyAsync(&count, c->stage.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
#endif
    double h[6];
    CUDA_TRY(c, cudaMemcpyAsync(h, res, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2] / count; out[3] = h[3]; out[4] = h[4]; out[5] = h[5];
    return 0;
}

// Sum of a host array over the ranks of the context's communicator, in place (MP::allreduce, mp.h:93-104): what the host side needs
// around a regrid on several partitions -- assembling the whole-domain state from the parts (every rank fills its own cells, zeros
// elsewhere: the sum is exact) and handing a fresh communicator id from rank 0 to everybody.  dtype 0 = double, 1 = unsigned byte.
extern "C" int nsem_allreduce_host(nsem_ctx* c, void* buf, uint64_t count, int dtype) {
    if (dtype != 0 && dtype != 1) { c->err = "nsem_allreduce_host: dtype must be 0 (double) or 1 (unsigned byte)"; return 1; }
    if (c->nranks <= 1 || count == 0) return 0;
#ifdef NSEM_WITH_NCCL
    if (!c->nccl) { c->err = "nsem_allreduce_host: the context has no communicator"; return 1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (join_comm(c)) return 1;
    const size_t bytes = (size_t)count * (dtype == 0 ? sizeof(double) : 1);
    const size_t words = (bytes + sizeof(double) - 1) / sizeof(double);
    DevBuf<double> tmp;
    CUDA_TRY(c, tmp.alloc(words));
    CUDA_TRY(c, cudaMemcpyAsync(tmp.p, buf, bytes, cudaMemcpyHostToDevice, c->stream));
    const ncclResult_t r = g_nccl.AllReduce(tmp.p, tmp.p, (size_t)count, dtype == 0 ? ncclDouble : ncclUint8, ncclSum, c->nccl, c->stream);
    if (r != ncclSuccess) { c->err = std::string("nsem_allreduce_host: ") + g_nccl.GetErrorString(r); tmp.release(); return 1; }
    CUDA_TRY(c, cudaMemcpyAsync(buf, tmp.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    tmp.release();
    return 0;
#else
    c->err = "nsem_allreduce_host: built without NCCL";
    return 1;
#endif
}

// ---------------------------------------------------------------------------------------------------------
// AMR: device-resident field transfer at a regrid (MeshField::refineField, field.h:1863-2015) and the restart
// branch of the solver set-up (euler.cpp:150-162); kernels in nsem_amr.cuh
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr uint32_t kAmrRemoved = 1u << 31;      // Constants::MAX_INT (tensor.h:455): cellMap entry of a cell that no longer exists

// which half of the parent a child covers along each element axis (field.h:1899-1909, 1949-1959): `nodes` = coordinates of the
// NP nodes of the cell whose nodes are the inputs (the child when merging, the parent when splitting)
uint32_t amr_child_half(const double* nodes, int NX, int NY, int NZ, const double* ccc, const double* ccp) {
    const double* v0 = nodes;
    const double* vv[3] = {nodes + (size_t)(NX - 1) * NY * NZ * 3, nodes + (size_t)(NY - 1) * NZ * 3, nodes + (size_t)(NZ - 1) * 3};
    uint32_t bits = 0;
    for (int d = 0; d < 3; d++) {
        double da = 0.0, db = 0.0;
        for (int k = 0; k < 3; k++) {
            const double e = vv[d][k] - v0[k];
            da += (ccc[k] - v0[k]) * e;
            db += (ccp[k] - v0[k]) * e;
        }
        if (!(da <= db)) bits |= 1u << d;
    }
    return bits;
}

// flat family lists [n, first, kid_0 .. kid_{n-1}]* -> offsets of the families; false when malformed
bool amr_parse(const uint32_t* m, uint32_t len, std::vector<uint32_t>& starts) {
    uint32_t i = 0;
    while (i < len) {
        if (i + 2 > len) return false;
        const uint32_t n = m[i];
        if (n == 0 || n > (uint32_t)amr::MAXKIDS || i + 2 + n > len) return false;
        starts.push_back(i);
        i += n + 2;
    }
    return true;
}
}  // namespace

extern "C" int nsem_refine_state(nsem_ctx* o, const nsem_regrid* r, nsem_ctx* c) {
    if (!o || !c || !r) { if (c) c->err = "nsem_refine_state: null argument"; return 1; }
    if (!(o->have_mesh && o->have_state)) { c->err = "nsem_refine_state: the old context has no state"; return 1; }
    if (!c->have_mesh) { c->err = "nsem_refine_state: upload the regridded mesh into the new context first"; return 1; }
    if (o == c) { c->err = "nsem_refine_state: old and new context must differ"; return 1; }
    if (o->device != c->device) { c->err = "nsem_refine_state: both contexts must live on the same device"; return 1; }
    if (o->NX != c->NX || o->NY != c->NY || o->NZ != c->NZ) { c->err = "nsem_refine_state: polynomial orders differ"; return 1; }
    if (r->n_cells_new != c->nB) { c->err = "nsem_refine_state: n_cells_new does not match the mesh of the new context"; return 1; }
    if (r->n_cell_map < o->nB) { c->err = "nsem_refine_state: cell_map shorter than the old mesh"; return 1; }
    if (!r->cell_map || !r->old_cV || !r->old_cC || !r->new_cV || !r->new_cC || !r->old_node_cC || (r->n_refine_map && !r->refine_map) ||
        (r->n_coarse_map && !r->coarse_map)) { c->err = "nsem_refine_state: missing array"; return 1; }
    for (int q = 0; q < 6; q++)
        if (!r->psi_ref[q] || !r->psi_cor[q]) { c->err = "nsem_refine_state: psi_ref / psi_cor tables are required"; return 1; }
    const int NX = o->NX, NY = o->NY, NZ = o->NZ, NP = o->NP;
    const uint32_t nOld = o->nB, nNew = c->nB;

    // ---- task tables on the host ----
    std::vector<uint8_t> written(nNew, 0);
    std::vector<uint32_t> copyOld, copyNew;
    for (uint32_t i = 0; i < nOld; i++) {
        const uint32_t id = r->cell_map[i];
        if (id == kAmrRemoved) continue;
        if (id >= nNew) { c->err = "nsem_refine_state: cell_map entry out of range"; return 1; }
        copyOld.push_back(i); copyNew.push_back(id);
        written[id] = 1;
    }
    std::vector<uint32_t> sStarts, mStarts;
    if (!amr_parse(r->refine_map, r->n_refine_map, sStarts) || !amr_parse(r->coarse_map, r->n_coarse_map, mStarts)) {
        c->err = "nsem_refine_state: malformed refine_map / coarse_map (families of 1..8 children expected)"; return 1;
    }
    std::vector<amr::Family> merges(mStarts.size()), splits(sStarts.size());
    for (size_t f = 0; f < mStarts.size(); f++) {
        const uint32_t* m = r->coarse_map + mStarts[f];
        amr::Family& F = merges[f];
        std::memset(&F, 0, sizeof F);
        if (m[1] >= r->n_cell_map || r->cell_map[m[1]] >= nNew) { c->err = "nsem_refine_state: coarse_map parent out of range"; return 1; }
        F.base = r->cell_map[m[1]];
        F.n = m[0];
        for (uint32_t j = 0; j < F.n; j++) {
            const uint32_t id1 = m[2 + j];
            if (id1 >= nOld) { c->err = "nsem_refine_state: coarse_map child out of range"; return 1; }
            F.kid[j] = id1;
            F.cv[j] = r->old_cV[id1];
            F.half[j] = amr_child_half(r->old_node_cC + (size_t)id1 * NP * 3, NX, NY, NZ, r->old_cC + (size_t)id1 * 3, r->new_cC + (size_t)F.base * 3);
        }
        written[F.base] = 1;
    }
    for (size_t f = 0; f < sStarts.size(); f++) {
        const uint32_t* m = r->refine_map + sStarts[f];
        amr::Family& F = splits[f];
        std::memset(&F, 0, sizeof F);
        if (m[1] >= nOld) { c->err = "nsem_refine_state: refine_map parent out of range"; return 1; }
        F.base = m[1];
        F.n = m[0];
        F.cvBase = r->old_cV[F.base];
        for (uint32_t j = 0; j < F.n; j++) {
            if (m[2 + j] >= r->n_cell_map || r->cell_map[m[2 + j]] >= nNew) { c->err = "nsem_refine_state: refine_map child out of range"; return 1; }
            const uint32_t id1 = r->cell_map[m[2 + j]];
            F.kid[j] = id1;
            F.cv[j] = r->new_cV[id1];
            F.half[j] = amr_child_half(r->old_node_cC + (size_t)F.base * NP * 3, NX, NY, NZ, r->new_cC + (size_t)id1 * 3, r->old_cC + (size_t)F.base * 3);
            written[id1] = 1;
        }
    }
    for (uint32_t i = 0; i < nNew; i++)
        if (!written[i]) { c->err = "nsem_refine_state: new cell " + std::to_string(i) + " is neither copied, split from nor merged into"; return 1; }
    std::vector<double> psiR(6 * 64, 0.0), psiC(6 * 64, 0.0), wnode(NP);
    const int nd[3] = {NX, NY, NZ};
    for (int q = 0; q < 6; q++) {
        const int n = nd[q / 2];
        std::memcpy(psiR.data() + q * 64, r->psi_ref[q], sizeof(double) * n * n);
        std::memcpy(psiC.data() + q * 64, r->psi_cor[q], sizeof(double) * n * n);
    }
    for (int i = 0, q = 0; i < NX; i++)
        for (int j = 0; j < NY; j++)
            for (int k = 0; k < NZ; k++, q++) wnode[q] = ((o->W[0][i] * o->W[1][j]) * o->W[2][k]) / 8;

    // ---- device ----
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (join_comm(o)) { c->err = o->err; return 1; }
    CUDA_TRY(c, cudaStreamSynchronize(o->stream));
    cudaStream_t s = c->stream;
    DevBuf<double> dPsiR, dPsiC, dW;
    DevBuf<amr::Family> dMerge, dSplit;
    DevBuf<uint32_t> dCo, dCn;
    CUDA_TRY(c, dPsiR.upload(psiR, s)); CUDA_TRY(c, dPsiC.upload(psiC, s)); CUDA_TRY(c, dW.upload(wnode, s));
    CUDA_TRY(c, dMerge.upload(merges, s)); CUDA_TRY(c, dSplit.upload(splits, s));
    CUDA_TRY(c, dCo.upload(copyOld, s)); CUDA_TRY(c, dCn.upload(copyNew, s));
    amr::Params A;
    std::memset(&A, 0, sizeof A);
    A.NX = NX; A.NY = NY; A.NZ = NZ; A.NP = NP;
    A.npsSrc = (uint32_t)o->NPS; A.npsDst = (uint32_t)c->NPS;
    const int ko = o->cur, kn = c->cur;
    const double* in[amr::NCOMP] = {o->rho[ko].p, o->U[ko][0].p, o->U[ko][1].p, o->U[ko][2].p, o->T[ko].p, o->p.p};
    double* out[amr::NCOMP] = {c->rho[kn].p, c->U[kn][0].p, c->U[kn][1].p, c->U[kn][2].p, c->T[kn].p, c->p.p};
    for (int f = 0; f < amr::NCOMP; f++) { A.in[f] = in[f]; A.out[f] = out[f]; }
    A.wnode = dW.p;
    A.copyOld = dCo.p; A.copyNew = dCn.p; A.nCopy = (uint32_t)copyOld.size();
    const unsigned threads = (unsigned)((NP + 31) / 32 * 32);
    const size_t smem = (size_t)(amr::NCOMP * NP + 2 * amr::NCOMP) * sizeof(double);
    if (A.nCopy) {
        const uint64_t n = (uint64_t)A.nCopy * NP;
        amr::copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(A);
        c->launches++;
    }
    if (!merges.empty()) {
        A.psi = dPsiC.p; A.fam = dMerge.p;
        amr::merge_kernel<<<(unsigned)merges.size(), threads, smem, s>>>(A);
        c->launches++;
    }
    if (!splits.empty()) {
        A.psi = dPsiR.p; A.fam = dSplit.p;
        amr::split_kernel<<<(unsigned)splits.size(), threads, smem, s>>>(A);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(s));       // the task tables are freed on return
    c->have_state = true;
    c->speed_valid = false;
    return 0;
}

extern "C" int nsem_restart_state(nsem_ctx* c) {
    if (check_ready(c, "nsem_restart_state")) return 1;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (join_comm(c)) return 1;
    KParams P;
    BCParams B;
    fill_kparams(c, P);
    const int k = c->cur;
    const uint64_t nReal = (uint64_t)c->nB * c->NP;
    amr::pressure_from_density_kernel<<<(unsigned)((nReal + 255) / 256), 256, 0, c->stream>>>(nReal, c->NP, c->NPS, P.P0, P.T0, P.R, P.gamma, c->rho[k].p,
                                                                                           c->T[k].p, c->p_ref.p, c->p.p);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    // ghost cells from the CURRENT state: the BC pass of sweep A (rho, p from the owner's EOS) without the gradients,
    // then the BC pass of sweep B (U, T)
    fill_bcparams(c, P, B, 0);
    B.visc = 0;
    B.rho_new = c->rho[k].p; B.T_old = c->T[k].p; B.p = c->p.p;
    for (int d = 0; d < 3; d++) B.U_new[d] = c->U[k][d].p;
    B.T_new = c->T[k].p;
    CUDA_TRY(c, launch_bc(c, B));
    B.phase = 1;
    CUDA_TRY(c, launch_bc(c, B));
    if (c->nG) c->launches += 2;
    c->speed_valid = false;
    if (!c->peers.empty()) return nsem_exchange_state_halos(c);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}
