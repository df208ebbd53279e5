// nsem_halo.cuh -- face-trace halo over NVLink peer memory (ASYNC_COMM, src/field/field.h:2255-2324; MP::isend/irecieve/waitall,
// src/mp/mp.h:117-131), without a communication library on the data path.
//
// Every rank owns ONE receive window (a single cudaMalloc, exported with cudaIpcGetMemHandle and mapped by its neighbours): a small header
// of arrival flags followed by regions [exchange kind][parity][field][receive slot].  An exchange is two small kernels on the compute
// stream:
//   halo_push_kernel  packs the owner-side face values of the exchanged arrays and STORES THEM STRAIGHT INTO THE NEIGHBOURS' WINDOWS over
//                     NVLink (one coalesced 8-byte store per value, laid out like the receiver's ghost cells), then the last block to finish
//                     publishes the exchange's epoch in every neighbour's flag word (fence.sys + st.release.sys);
//   halo_pull_kernel  waits until every neighbour's flag has reached the epoch (ld.acquire.sys, bounded spin) and copies the window into
//                     the ghost regions of the arrays.
// No rendezvous, no per-field messages, no host involvement per step: the 14 + 5 arrays of a step travel as two kernels per rank instead
// of 2 x 19 x (number of neighbours) library point-to-point operations.  Ordering between exchanges needs no extra handshake: a neighbour
// can only produce the data of exchange e + 2 of a kind after it has consumed what this rank sent after ITS exchange e (the sweeps
// alternate A, B, A, B, ...), so two parities per kind are enough.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nsem {

constexpr int HALO_MAX_PEERS = 32;
constexpr int HALO_KINDS = 3;                       // 0 = after sweep A (rho_new, p', gradients), 1 = after sweep B (U, T, S), 2 = state (set-up, restart)
constexpr int HALO_HEADER_BYTES = 4096;             // flags[HALO_KINDS][HALO_MAX_PEERS] (u64), then the push counter and the error word
__host__ __device__ constexpr int halo_kind_fields(int kind) { return kind == 0 ? 14 : (kind == 1 ? 5 : 8); }
// offset (in doubles, after the header) of region (kind, parity) in a window whose owner receives nRecv slots
__host__ __device__ constexpr uint64_t halo_region_offset(int kind, int parity, uint64_t nRecv) {
    uint64_t f = 0;
    for (int k = 0; k < kind; k++) f += 2 * (uint64_t)halo_kind_fields(k);
    return (f + (uint64_t)parity * halo_kind_fields(kind)) * nRecv;
}
__host__ __device__ constexpr uint64_t halo_window_doubles(uint64_t nRecv) { return halo_region_offset(HALO_KINDS, 0, nRecv); }

struct HaloPushParams {
    int nfields, npeers;
    uint64_t nslots;                                // send slots of all neighbours together
    const uint32_t* node;                           // [nslots] owner node (device index) or 0xffffffff for padding slots
    const double* src[16];
    uint64_t off[HALO_MAX_PEERS + 1];               // first send slot of every neighbour (off[npeers] = nslots)
    double* win[HALO_MAX_PEERS];                    // neighbour's region (kind, parity), at the slot where this rank's block starts
    uint64_t stride[HALO_MAX_PEERS];                // doubles between two fields there (= the neighbour's receive slots)
    unsigned long long* flag[HALO_MAX_PEERS];       // neighbour's arrival flag for (kind, this rank)
    unsigned long long epoch;
    unsigned int* counter;                          // local: blocks that have finished their stores
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) halo_push_kernel(const __grid_constant__ HaloPushParams H) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < H.nslots) {
        int p = 0;
        while (p + 1 < H.npeers && q >= H.off[p + 1]) p++;
        const uint64_t local = q - H.off[p];
        const uint32_t nd = H.node[q];
        double* w = H.win[p] + local;
        const uint64_t st = H.stride[p];
        for (int f = 0; f < H.nfields; f++) w[(uint64_t)f * st] = (nd != 0xffffffffu) ? H.src[f][nd] : 0.0;
    }
    // publish: every block's stores are ordered before its counter increment; the last block orders all of them before the flags
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int done = atomicAdd(H.counter, 1u);
        if (done == gridDim.x - 1) {
            *H.counter = 0;
            __threadfence_system();
            for (int p = 0; p < H.npeers; p++) st_release_sys(H.flag[p], H.epoch);
        }
    }
}

struct HaloPullParams {
    int nfields, npeers;
    uint64_t nslots;                                // receive slots
    const double* win;                              // my region (kind, parity)
    uint64_t stride;                                // = nslots
    double* dst[16];
    uint64_t ghostBase;
    uint64_t off[HALO_MAX_PEERS + 1];               // first receive slot of every neighbour
    uint64_t ghostOff[HALO_MAX_PEERS];              // g0 * GPS: where that neighbour's ghost cells start behind ghostBase
    const unsigned long long* flag;                 // my flags for this kind [HALO_MAX_PEERS]
    unsigned long long epoch;
    unsigned long long timeout_ns;
    int* error;                                     // set when a neighbour never arrived (the host reports it at the next synchronisation)
    unsigned long long* wait_ns;                    // += the time block 0 spent waiting for the neighbours' flags (diagnostics: nsem_halo_wait_ms)
};

__global__ void __launch_bounds__(256) halo_pull_kernel(const __grid_constant__ HaloPullParams H) {
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = 1;
        const unsigned long long t0 = global_timer_ns();
        for (int p = 0; p < H.npeers; p++) {
            while (ld_acquire_sys(H.flag + p) < H.epoch) {
                __nanosleep(200);
                if (global_timer_ns() - t0 > H.timeout_ns) { ok = 0; *H.error = 1 + p; break; }
            }
            if (!ok) break;
        }
        if (blockIdx.x == 0) atomicAdd(H.wait_ns, global_timer_ns() - t0);
    }
    __syncthreads();
    if (!ok) return;
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= H.nslots) return;
    int p = 0;
    while (p + 1 < H.npeers && q >= H.off[p + 1]) p++;
    const uint64_t g = H.ghostBase + H.ghostOff[p] + (q - H.off[p]);
    // the window was written by another GPU: read it past L1 (the acquire above ordered the loads after the flag)
    for (int f = 0; f < H.nfields; f++) H.dst[f][g] = __ldcg(H.win + (uint64_t)f * H.stride + q);
}

}  // namespace nsem
