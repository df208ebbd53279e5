// nsem_amr.cuh -- device-resident AMR field transfer: MeshField::refineField (src/field/field.h:1863-2015) for the
// resident state (rho, U, T, p), from the SoA arrays of the context that holds the old mesh into the SoA arrays of a
// context that holds the regridded mesh.  The reference round-trips every field through files at a regrid
// (Prepare::refineMesh, field.cpp:625-955); here the state never leaves the device.
//
// The floating-point operation order of the reference is kept exactly (explicit mul/add intrinsics, no contraction):
//   * a copied cell is copied (field.h:1878-1884);
//   * a merged cell is the volume-weighted projection of its children, every output node accumulating
//     ((P*cV_child)*fx*fy*fz) over (child, input node) in order, divided by the summed volume (field.h:1887-1934);
//   * a split cell's children are interpolated, every output node accumulating (P*fx*fy*fz) over the input nodes in
//     order, then the whole family is scaled by |integral over the parent| / |integral over the children|; both integrals
//     are SEQUENTIAL sums in the reference (over children, input nodes, output nodes), so one thread per field component
//     walks them in that order (field.h:1937-2000).
// One CTA per family; regrids are rare (every amr_step * write_interval steps), the kernels are sized for clarity.
#pragma once
#include <cstdint>

#include "nsem_kernels.cuh"

namespace amr {

constexpr int MAXKIDS = 8;           // refineMesh splits a hexahedron in up to three directions (dir = 7, field.cpp:884)
constexpr int NCOMP = 6;             // rho | U0 U1 U2 | T | p

struct Family {
    uint32_t base;                   // split: the old cell;   merge: the new cell
    uint32_t n;                      // children
    uint32_t kid[MAXKIDS];           // split: new cells;      merge: old cells
    uint32_t half[MAXKIDS];          // bit d: the child covers the upper half of the parent along element axis d
    double cv[MAXKIDS];              // split: gCV of the new cells;  merge: gCV of the old cells
    double cvBase;                   // split: gCV of the old cell
};

struct Params {
    int NX, NY, NZ, NP;
    uint32_t npsSrc, npsDst;         // element stride of the node arrays
    const double* in[NCOMP];
    double* out[NCOMP];
    const double* psi;               // [6][8*8] psiRef (split) or psiCor (merge), table d*2+half, [in*n + out]
    const double* wnode;             // [NP] ((w_i*w_j)*w_k)/8
    const Family* fam;
    const uint32_t* copyOld;         // copy pairs
    const uint32_t* copyNew;
    uint32_t nCopy;
};

__device__ __forceinline__ void node_ijk(int q, int NY, int NZ, int& i, int& j, int& k) {
    i = q / (NY * NZ);
    j = (q / NZ) % NY;
    k = q % NZ;
}

__global__ void copy_kernel(const Params A) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)A.nCopy * A.NP) return;
    const uint32_t c = (uint32_t)(gid / A.NP);
    const int q = (int)(gid % A.NP);
    const size_t si = (size_t)A.copyOld[c] * A.npsSrc + q, di = (size_t)A.copyNew[c] * A.npsDst + q;
#pragma unroll
    for (int f = 0; f < NCOMP; f++) A.out[f][di] = A.in[f][si];
}

// dynamic shared memory: NCOMP*NP doubles (values of the cell whose nodes are the inputs)
__global__ void merge_kernel(const Params A) {
    extern __shared__ double sP[];
    const Family& F = A.fam[blockIdx.x];
    const int NP = A.NP, t = threadIdx.x;
    int io = 0, jo = 0, ko = 0;
    if (t < NP) node_ijk(t, A.NY, A.NZ, io, jo, ko);
    double acc[NCOMP];
#pragma unroll
    for (int f = 0; f < NCOMP; f++) acc[f] = 0.0;
    double vol = 0.0;
    for (uint32_t j = 0; j < F.n; j++) {
        __syncthreads();
        if (t < NP) {
#pragma unroll
            for (int f = 0; f < NCOMP; f++) sP[f * NP + t] = __dmul_rn(A.in[f][(size_t)F.kid[j] * A.npsSrc + t], F.cv[j]);
        }
        __syncthreads();
        const double* fxT = A.psi + (0 * 2 + (F.half[j] & 1)) * 64;
        const double* fyT = A.psi + (1 * 2 + ((F.half[j] >> 1) & 1)) * 64;
        const double* fzT = A.psi + (2 * 2 + ((F.half[j] >> 2) & 1)) * 64;
        if (t < NP) {
            for (int q = 0; q < NP; q++) {
                int ii, jj, kk;
                node_ijk(q, A.NY, A.NZ, ii, jj, kk);
                const double fx = fxT[ii * A.NX + io], fy = fyT[jj * A.NY + jo], fz = fzT[kk * A.NZ + ko];
#pragma unroll
                for (int f = 0; f < NCOMP; f++)
                    acc[f] = __dadd_rn(acc[f], __dmul_rn(__dmul_rn(__dmul_rn(sP[f * NP + q], fx), fy), fz));
            }
        }
        vol = __dadd_rn(vol, F.cv[j]);
    }
    if (t < NP) {
#pragma unroll
        for (int f = 0; f < NCOMP; f++) A.out[f][(size_t)F.base * A.npsDst + t] = __ddiv_rn(acc[f], vol);
    }
}

// dynamic shared memory: NCOMP*NP doubles (the parent) + 2*NCOMP doubles (the two integrals per component)
__global__ void split_kernel(const Params A) {
    extern __shared__ double sP[];
    const Family& F = A.fam[blockIdx.x];
    const int NP = A.NP, t = threadIdx.x;
    double* tot = sP + NCOMP * NP;
    if (t < NP) {
#pragma unroll
        for (int f = 0; f < NCOMP; f++) sP[f * NP + t] = A.in[f][(size_t)F.base * A.npsSrc + t];
    }
    __syncthreads();
    int io = 0, jo = 0, ko = 0;
    if (t < NP) node_ijk(t, A.NY, A.NZ, io, jo, ko);
    // interpolation onto the children, unscaled
    for (uint32_t j = 0; j < F.n; j++) {
        const double* fxT = A.psi + (0 * 2 + (F.half[j] & 1)) * 64;
        const double* fyT = A.psi + (1 * 2 + ((F.half[j] >> 1) & 1)) * 64;
        const double* fzT = A.psi + (2 * 2 + ((F.half[j] >> 2) & 1)) * 64;
        if (t < NP) {
            double acc[NCOMP];
#pragma unroll
            for (int f = 0; f < NCOMP; f++) acc[f] = 0.0;
            for (int q = 0; q < NP; q++) {
                int ii, jj, kk;
                node_ijk(q, A.NY, A.NZ, ii, jj, kk);
                const double fx = fxT[ii * A.NX + io], fy = fyT[jj * A.NY + jo], fz = fzT[kk * A.NZ + ko];
#pragma unroll
                for (int f = 0; f < NCOMP; f++)
                    acc[f] = __dadd_rn(acc[f], __dmul_rn(__dmul_rn(__dmul_rn(sP[f * NP + q], fx), fy), fz));
            }
#pragma unroll
            for (int f = 0; f < NCOMP; f++) A.out[f][(size_t)F.kid[j] * A.npsDst + t] = acc[f];
        }
    }
    // the two integrals, sequential per component in the reference's order (field.h:1966-1985)
    if (t < NCOMP) {
        const int f = t;
        double toto = 0.0, totn = 0.0;
        for (int q = 0; q < NP; q++) toto = __dadd_rn(toto, __dmul_rn(__dmul_rn(sP[f * NP + q], F.cvBase), A.wnode[q]));
        for (uint32_t j = 0; j < F.n; j++) {
            const double* fxT = A.psi + (0 * 2 + (F.half[j] & 1)) * 64;
            const double* fyT = A.psi + (1 * 2 + ((F.half[j] >> 1) & 1)) * 64;
            const double* fzT = A.psi + (2 * 2 + ((F.half[j] >> 2) & 1)) * 64;
            const double cv = F.cv[j];
            for (int q = 0; q < NP; q++) {
                int ii, jj, kk;
                node_ijk(q, A.NY, A.NZ, ii, jj, kk);
                const double P0 = sP[f * NP + q];
                int o = 0;
                for (int i1 = 0; i1 < A.NX; i1++) {
                    const double px = __dmul_rn(P0, fxT[ii * A.NX + i1]);
                    for (int j1 = 0; j1 < A.NY; j1++) {
                        const double pxy = __dmul_rn(px, fyT[jj * A.NY + j1]);
                        for (int k1 = 0; k1 < A.NZ; k1++, o++) {
                            const double P1 = __dmul_rn(pxy, fzT[kk * A.NZ + k1]);
                            totn = __dadd_rn(totn, __dmul_rn(__dmul_rn(P1, cv), A.wnode[o]));
                        }
                    }
                }
            }
        }
        tot[f] = toto;
        tot[NCOMP + f] = totn;
    }
    __syncthreads();
    if (t < NP) {
        // factor = sdiv(mag(toto), mag(totn)) per FIELD (tensor.h:71-76; mag of a Vector = sqrt(x*x + y*y + z*z))
        double fac[NCOMP];
        const int first[4] = {0, 1, 4, 5}, count[4] = {1, 3, 1, 1};
#pragma unroll
        for (int g = 0; g < 4; g++) {
            double mo, mn;
            if (count[g] == 1) {
                mo = fabs(tot[first[g]]);
                mn = fabs(tot[NCOMP + first[g]]);
            } else {
                const double* a = tot + first[g];
                const double* b = tot + NCOMP + first[g];
                mo = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a[0], a[0]), __dmul_rn(a[1], a[1])), __dmul_rn(a[2], a[2])));
                mn = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(b[0], b[0]), __dmul_rn(b[1], b[1])), __dmul_rn(b[2], b[2])));
            }
            const double fv = (mn != 0.0) ? __ddiv_rn(mo, mn) : 0.0;
            for (int k = 0; k < count[g]; k++) fac[first[g] + k] = fv;
        }
        for (uint32_t j = 0; j < F.n; j++) {
#pragma unroll
            for (int f = 0; f < NCOMP; f++) {
                double* o = A.out[f] + (size_t)F.kid[j] * A.npsDst + t;
                *o = __dmul_rn(*o, fac[f]);           // this thread wrote the value above
            }
        }
    }
}

// p = P0 (rho (T + T0) R / P0)^gamma - p_ref over the real nodes: the restart branch of the set-up (euler.cpp:150-162)
__global__ void pressure_from_density_kernel(uint64_t nReal, int NP, int NPS, double P0, double T0, double R, double gamma,
                                             const double* __restrict__ rho, const double* __restrict__ T, const double* __restrict__ p_ref,
                                             double* __restrict__ p) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nReal) return;
    const size_t i = (size_t)(g / NP) * NPS + (g % NP);
    p[i] = __dsub_rn(nsem::eos_pressure(P0, R, gamma, rho[i], __dadd_rn(T[i], T0)), p_ref[i]);
}

}  // namespace amr
