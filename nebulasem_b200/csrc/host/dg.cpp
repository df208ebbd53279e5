// dg.cpp -- LGL basis tables and per-node DG geometry on the host.
//
//   legendre / legendre_gauss_lobatto     src/field/dg.cpp:32-99   (Newton iteration from Chebyshev guesses)
//   lagrange_basis / _derivative          src/field/dg.cpp:103-143
//   init_poly                             src/field/dg.cpp:147-163 (NPF = product of the two larger extents)
//   init_geom                             src/field/dg.cpp:167-477 (transfinite node coordinates, FO/FN maps,
//                                                                   face weights, Jinv = ((J^T J)^-1 J^T)^T)
//   initGeomMeshFields                    src/field/field.cpp:171-280 (allFaces/faceIndices, fI rule)
// Geometry is NOT Jacobian-determinant based in the reference: cV = V_element * w_i w_j w_k / 8 and
// fN = A_face * w_a w_b / 4 (SURVEY finding 4); the same definitions are used here so that field files and
// results are interchangeable.
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <string>

#include "nsem_host.h"

namespace nsemh {

static void legendre(int p, double x, double& L0, double& L0_1, double& L0_2) {
    double L1 = 0, L1_1 = 0, L1_2 = 0, L2, L2_1, L2_2;
    L0 = 1; L0_1 = 0; L0_2 = 0;
    for (int i = 1; i <= p; i++) {
        L2 = L1; L2_1 = L1_1; L2_2 = L1_2;
        L1 = L0; L1_1 = L0_1; L1_2 = L0_2;
        const double a = (2 * i - 1.0) / i, b = (i - 1.0) / i;
        L0 = a * x * L1 - b * L2;
        L0_1 = a * (L1 + x * L1_1) - b * L2_1;
        L0_2 = a * (2 * L1_1 + x * L1_2) - b * L2_2;
    }
}

void legendre_gauss_lobatto(int N, double* xgl, double* wgl) {
    if (N == 1) { xgl[0] = 0; wgl[0] = 2; return; }
    const int p = N - 1, ph = N / 2;
    const double PI = 3.14159265358979323846264;
    double L0, L0_1, L0_2;
    for (int i = 0; i < ph; i++) {
        double x = std::cos((2 * i + 1) * PI / (2 * N));
        for (int k = 1; k <= 20; k++) {
            legendre(p, x, L0, L0_1, L0_2);
            const double dx = -(1 - x * x) * L0_1 / (-2 * x * L0_1 + (1 - x * x) * L0_2);
            x += dx;
            if (std::fabs(dx) < 1.0e-20) break;
        }
        xgl[p - i] = x;
        wgl[p - i] = 2 / (p * (p + 1) * L0 * L0);
    }
    if (N != 2 * ph) {
        legendre(p, 0.0, L0, L0_1, L0_2);
        xgl[ph] = 0;
        wgl[ph] = 2 / (p * (p + 1) * L0 * L0);
    }
    for (int i = 0; i < ph; i++) { xgl[i] = -xgl[p - i]; wgl[i] = wgl[p - i]; }
}

void lagrange_basis(int N, const double* xgl, int Ns, const double* xs, double* psi) {
    for (int s = 0; s < Ns; s++)
        for (int j = 0; j < N; j++) {
            double prod = 1;
            for (int k = 0; k < N; k++)
                if (k != j) prod *= ((xs[s] - xgl[k]) / (xgl[j] - xgl[k]));
            psi[j * Ns + s] = prod;
        }
}

void lagrange_basis_derivative(int N, const double* xgl, int Ns, const double* xs, double* dpsi) {
    for (int s = 0; s < Ns; s++)
        for (int i = 0; i < N; i++) {
            double acc = 0;
            for (int j = 0; j < N; j++) {
                if (i == j) continue;
                double prod = 1;
                for (int k = 0; k < N; k++)
                    if (k != i && k != j) prod *= ((xs[s] - xgl[k]) / (xgl[i] - xgl[k]));
                acc += prod / (xgl[i] - xgl[j]);
            }
            dpsi[s * N + i] = acc;
        }
}

// matinv / matmul of tensor.cpp:170-243: Gauss-Jordan with one bubble pass on column 0 and no further pivoting, final
// row scaling; sequential inner products
static void matinv(int N, std::vector<double> A, std::vector<double>& X) {
    X.assign((size_t)N * N, 0.0);
    for (int i = 0; i < N; i++) X[(size_t)i * N + i] = 1.0;
    for (int i = N - 1; i > 0; i--)
        if (A[(size_t)(i - 1) * N] < A[(size_t)i * N])
            for (int k = 0; k < N; k++) {
                std::swap(A[(size_t)i * N + k], A[(size_t)(i - 1) * N + k]);
                std::swap(X[(size_t)i * N + k], X[(size_t)(i - 1) * N + k]);
            }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            if (j != i) {
                const double temp = A[(size_t)j * N + i] / A[(size_t)i * N + i];
                for (int k = 0; k < N; k++) {
                    A[(size_t)j * N + k] -= A[(size_t)i * N + k] * temp;
                    X[(size_t)j * N + k] -= X[(size_t)i * N + k] * temp;
                }
            }
    for (int i = 0; i < N; i++) {
        const double temp = A[(size_t)i * N + i];
        for (int j = 0; j < N; j++) {
            A[(size_t)i * N + j] = A[(size_t)i * N + j] / temp;
            X[(size_t)i * N + j] = X[(size_t)i * N + j] / temp;
        }
    }
}
static void matmul(int N, const std::vector<double>& A, const std::vector<double>& B, std::vector<double>& X) {
    X.assign((size_t)N * N, 0.0);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            double acc = 0.0;
            for (int k = 0; k < N; k++) acc += A[(size_t)i * N + k] * B[(size_t)k * N + j];
            X[(size_t)i * N + j] = acc;
        }
}

Basis::Basis(const int nop[3]) {
    NPX = nop[0] + 1; NPY = nop[1] + 1; NPZ = nop[2] + 1;
    NP = NPX * NPY * NPZ;
    if (NPX <= NPY && NPX <= NPZ) NPF = NPY * NPZ;
    else if (NPY <= NPX && NPY <= NPZ) NPF = NPX * NPZ;
    else NPF = NPX * NPY;
    for (int d = 0; d < 3; d++) {
        const int m = n(d);
        xgl[d].resize(m); wgl[d].resize(m); psi[d].resize(m * m); dpsi[d].resize(m * m);
        legendre_gauss_lobatto(m, xgl[d].data(), wgl[d].data());
        lagrange_basis(m, xgl[d].data(), m, xgl[d].data(), psi[d].data());
        lagrange_basis_derivative(m, xgl[d].data(), m, xgl[d].data(), dpsi[d].data());
        // ---- 2:1 mortar projections (dg.cpp:497-590): L2 projection between a coarse face and its two halves, mass
        // matrices integrated with the (m+1)-point LGL rule ----
        const int me = m + 1;
        std::vector<double> xe(me), we(me), xre[2], psie((size_t)m * me), psire[2];
        legendre_gauss_lobatto(me, xe.data(), we.data());
        for (int c = 0; c < 2; c++) {
            xre[c].assign(me, 0.0);
            if (m != 1)
                for (int q = 0; q < me; q++) xre[c][q] = (c == 0 ? -0.5 : 0.5) + xe[q] / 2;
            psire[c].resize((size_t)m * me);
            lagrange_basis(m, xgl[d].data(), me, xre[c].data(), psire[c].data());
        }
        lagrange_basis(m, xgl[d].data(), me, xe.data(), psie.data());
        std::vector<double> Mcc((size_t)m * m, 0.0), Msc[2], Mga[2], iMcc, P;
        for (int c = 0; c < 2; c++) { Msc[c].assign((size_t)m * m, 0.0); Mga[c].assign((size_t)m * m, 0.0); }
        for (int j = 0; j < m; j++)
            for (int k = 0; k < m; k++)
                for (int q = 0; q < me; q++) {
                    Mcc[(size_t)j * m + k] += (we[q] / 2) * psie[(size_t)j * me + q] * psie[(size_t)k * me + q];
                    for (int c = 0; c < 2; c++) {
                        const double vs = (we[q] / 2) * psie[(size_t)j * me + q] * psire[c][(size_t)k * me + q];
                        Msc[c][(size_t)j * m + k] += vs;
                        Mga[c][(size_t)k * m + j] += vs;
                    }
                }
        matinv(m, Mcc, iMcc);
        for (int c = 0; c < 2; c++) {
            auto transposed = [&](const std::vector<double>& A) {
                std::vector<double> T((size_t)m * m);
                for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) T[(size_t)j * m + i] = A[(size_t)i * m + j];
                return T;
            };
            matmul(m, iMcc, Msc[c], P);
            psiRef[d * 2 + c] = transposed(P);
            matmul(m, iMcc, Mga[c], P);
            psiCor[d * 2 + c] = transposed(P);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
namespace {
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }   // AddScalarOperators: each component * s
inline V3 toV(const Vec3& a) { return {a[0], a[1], a[2]}; }
inline double mag3(V3 a) { return std::sqrt(a.x * a.x + (a.y * a.y + a.z * a.z)); }   // Unroll<3>::dot nests to the right

// tensor.h:522-571, same term order
inline V3 interp_face(double r, double s, V3 x00, V3 x01, V3 x10, V3 x11, V3 xr0, V3 xr1, V3 x0s, V3 x1s) {
    return -((1.0 - r) * (1.0 - s)) * x00 + (1.0 - r) * x0s - ((1.0 - r) * s) * x01 + (1.0 - s) * xr0 + s * xr1 -
           (r * (1.0 - s)) * x10 + r * x1s - (r * s) * x11;
}
inline V3 interp_cell(double r, double s, double t, V3 x000, V3 x001, V3 x010, V3 x011, V3 x100, V3 x101, V3 x110, V3 x111,
                      V3 xr00, V3 xr01, V3 xr10, V3 xr11, V3 x0s0, V3 x0s1, V3 x1s0, V3 x1s1, V3 x00t, V3 x01t, V3 x10t,
                      V3 x11t, V3 x0st, V3 x1st, V3 xr0t, V3 xr1t, V3 xrs0, V3 xrs1) {
    return ((1.0 - r) * (1.0 - s) * (1.0 - t)) * x000 - ((1.0 - r) * (1.0 - s)) * x00t + ((1.0 - r) * (1.0 - s) * t) * x001 -
           ((1.0 - r) * (1.0 - t)) * x0s0 + (1.0 - r) * x0st - ((1.0 - r) * t) * x0s1 + ((1.0 - r) * s * (1.0 - t)) * x010 -
           ((1.0 - r) * s) * x01t + ((1.0 - r) * s * t) * x011 - ((1.0 - s) * (1.0 - t)) * xr00 + (1.0 - s) * xr0t -
           ((1.0 - s) * t) * xr01 + (1.0 - t) * xrs0 + t * xrs1 - (s * (1.0 - t)) * xr10 + s * xr1t - (s * t) * xr11 +
           (r * (1.0 - s) * (1.0 - t)) * x100 - (r * (1.0 - s)) * x10t + (r * (1.0 - s) * t) * x101 - (r * (1.0 - t)) * x1s0 +
           r * x1st - (r * t) * x1s1 + (r * s * (1.0 - t)) * x110 - (r * s) * x11t + (r * s * t) * x111;
}

// 3x3 helpers in row-major [a][d], operation order of tensor.cpp:42-58 (mul) and :152-168 (inv)
inline void mul33(const double p[9], const double q[9], double r[9]) {
    for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++) r[a * 3 + c] = (p[a * 3 + 0] * q[0 * 3 + c] + p[a * 3 + 1] * q[1 * 3 + c]) + p[a * 3 + 2] * q[2 * 3 + c];
}
inline void inv33(const double p[9], double r[9]) {
    r[0] = p[4] * p[8] - p[5] * p[7];
    r[4] = p[0] * p[8] - p[2] * p[6];
    r[8] = p[0] * p[4] - p[1] * p[3];
    r[1] = p[2] * p[7] - p[1] * p[8];
    r[2] = p[1] * p[5] - p[2] * p[4];
    r[3] = p[5] * p[6] - p[3] * p[8];
    r[5] = p[2] * p[3] - p[0] * p[5];
    r[6] = p[3] * p[7] - p[4] * p[6];
    r[7] = p[1] * p[6] - p[0] * p[7];
    const double d = (p[0] * r[0] + p[1] * r[3]) + p[2] * r[6];
    if (d == 0) { for (int q = 0; q < 9; q++) r[q] = 0; }
    else for (int q = 0; q < 9; q++) r[q] /= d;
}
}  // namespace

void Geometry::build(const MeshTopo& t, const Basis& b) {
    const int NPX = b.NPX, NPY = b.NPY, NPZ = b.NPZ, NP = b.NP, NPF = b.NPF;
    nBCS = t.nBCS; nCells = t.nCells(); nFacets = t.nFacets();
    gBCSfield = (uint64_t)nBCS * NP;
    gALL = (uint64_t)nCells * NP;
    spherical = t.spherical; sphere_radius = t.sphere_radius;
    if (gALL + NP >= 0xffffffffull) throw Error("mesh exceeds the 32-bit node index of the reference layout");
    auto I4 = [&](uint64_t c, int i, int j, int k) { return (u32)(c * NP + (uint64_t)i * NPY * NPZ + j * NPZ + k); };

    const bool verbose = std::getenv("NSEM_VERBOSE") != nullptr;
    auto tl = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto t1 = std::chrono::steady_clock::now();
        if (verbose) std::printf("  geometry: %-34s %.3f s\n", what, std::chrono::duration<double>(t1 - tl).count());
        tl = t1;
    };
    faceBegin.resize(nCells); faceEnd.resize(nCells);
    allFaces = t.cellFaces; faceID = t.cellFaceID;
    for (u32 i = 0; i < nCells; i++) { faceBegin[i] = t.cellStart[i]; faceEnd[i] = t.cellStart[i + 1]; }
    faceOwner = t.FOC; faceNeigh = t.FNC; faceMortar = t.FMC;
    faceNormal.resize((size_t)nFacets * 3);
    faceCenter.resize((size_t)nFacets * 3);
    for (u32 f = 0; f < nFacets; f++)
        for (int d = 0; d < 3; d++) {
            faceNormal[(size_t)f * 3 + d] = t.FN[f][d];
            faceCenter[(size_t)f * 3 + d] = t.FC[f][d];
        }
    for (int q = 0; q < 6; q++) { psiRef[q] = b.psiRef[q]; psiCor[q] = b.psiCor[q]; }

    cC.assign(gALL * 3, 0.0); cV.assign(gALL, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)nCells; c++)
        for (int q = 0; q < NP; q++) {
            const uint64_t i = (uint64_t)c * NP + q;
            cV[i] = t.CV[c];
            for (int d = 0; d < 3; d++) cC[i * 3 + d] = t.CC[c][d];
        }
    const size_t nfn = (size_t)nFacets * NPF;
    fC.resize(nfn * 3); fN.resize(nfn * 3); fI.resize(nfn);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nFacets; f++)
        for (int q = 0; q < NPF; q++)
            for (int d = 0; d < 3; d++) {
                fC[((size_t)f * NPF + q) * 3 + d] = t.FC[f][d];
                fN[((size_t)f * NPF + q) * 3 + d] = t.FN[f][d];
            }
    FO.assign(nfn, (u32)gALL);
    FN.assign(nfn, (u32)gALL);

    lap("tables + array fills");
    const double* xg[3] = {b.xgl[0].data(), b.xgl[1].data(), b.xgl[2].data()};
    const double* wg[3] = {b.wgl[0].data(), b.wgl[1].data(), b.wgl[2].data()};

    // ---- node coordinates (dg.cpp:176-325) ----
    static const int sides[12][2] = {{0, 1}, {3, 2}, {7, 6}, {4, 5}, {0, 3}, {1, 2}, {5, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    std::string corner_error;      // an exception must not leave the parallel region
    const bool sph = t.spherical;
#pragma omp parallel for schedule(static)
    for (int64_t cs = 0; cs < (int64_t)nBCS; cs++) {
        const u32 ci = (u32)cs;
        const u32 c0 = t.cellStart[ci], c1 = t.cellStart[ci + 1];
        const u32 id0 = t.cellFaceID[c0], id1 = id0 ^ 1u;
        u32 vpi[8], vidx[8];
        // one quadrilateral on each of the two sides that define the corners: the short route
        int n0 = 0, n1 = 0;
        bool simple = true;
        for (u32 q = c0; q < c1; q++) {
            const bool quad = (t.facetStart[t.cellFaces[q] + 1] - t.facetStart[t.cellFaces[q]] == 4);
            if (t.cellFaceID[q] == id0) { n0++; simple = simple && quad; }
            else if (t.cellFaceID[q] == id1) { n1++; simple = simple && quad; }
        }
        simple = simple && n0 == 1 && n1 == 1;
        if (simple) {
            const u32 *f1 = nullptr, *f2 = nullptr;
            for (u32 q = c0; q < c1; q++) {
                if (t.cellFaceID[q] == id0 && !f1) f1 = &t.facetVerts[t.facetStart[t.cellFaces[q]]];
                else if (t.cellFaceID[q] == id1 && !f2) f2 = &t.facetVerts[t.facetStart[t.cellFaces[q]]];
            }
            if (!f1 || !f2) continue;   // cannot happen after fix_hex_cells
            t.hex_corners(f1, f2, vpi);
        } else {
            // non-conforming cell: the sub-facets of the two sides are merged first (dg.cpp:190-215)
            try {
                t.hex_corners_poly(t.merged_side(ci, id0), t.merged_side(ci, id1), vpi);
            } catch (const std::exception& e) {
#pragma omp critical
                corner_error = e.what();
                continue;
            }
        }
        static const int ord2[8] = {0, 1, 5, 4, 3, 2, 6, 7}, ord4[8] = {0, 3, 7, 4, 1, 2, 6, 5};
        for (int q = 0; q < 8; q++) vidx[id0 == 2 ? ord2[q] : (id0 == 4 ? ord4[q] : q)] = vpi[q];
        V3 vp[8];
        for (int q = 0; q < 8; q++) vp[q] = toV(t.V[vidx[q]]);
        for (int i = 0; i < NPX; i++)
            for (int j = 0; j < NPY; j++)
                for (int k = 0; k < NPZ; k++) {
                    const double rx = (xg[0][i] + 1) / 2, ry = (xg[1][j] + 1) / 2, rz = (xg[2][k] + 1) / 2;
                    V3 vd[12], vf[6];
                    for (int w = 0; w < 12; w++) {
                        const double m = (w < 4) ? rx : (w < 8 ? ry : rz);
                        vd[w] = (1 - m) * vp[sides[w][0]] + m * vp[sides[w][1]];
                        // on the sphere every blended point is pushed back to the blended radius (dg.cpp:257-285)
                        if (sph) vd[w] = (((1 - m) * mag3(vp[sides[w][0]]) + m * mag3(vp[sides[w][1]])) / mag3(vd[w])) * vd[w];
                    }
                    vf[0] = interp_face(rx, ry, vp[0], vp[3], vp[1], vp[2], vd[0], vd[1], vd[4], vd[5]);
                    vf[1] = interp_face(rx, ry, vp[4], vp[7], vp[5], vp[6], vd[3], vd[2], vd[7], vd[6]);
                    vf[2] = interp_face(rx, rz, vp[0], vp[4], vp[1], vp[5], vd[0], vd[3], vd[8], vd[9]);
                    vf[3] = interp_face(rx, rz, vp[3], vp[7], vp[2], vp[6], vd[1], vd[2], vd[11], vd[10]);
                    vf[4] = interp_face(ry, rz, vp[0], vp[4], vp[3], vp[7], vd[4], vd[7], vd[8], vd[11]);
                    vf[5] = interp_face(ry, rz, vp[1], vp[5], vp[2], vp[6], vd[5], vd[6], vd[9], vd[10]);
                    if (sph) {
                        static const int ir0[6] = {0, 3, 0, 1, 4, 5};
                        for (int w = 0; w < 6; w++) vf[w] = (mag3(vd[ir0[w]]) / mag3(vf[w])) * vf[w];
                    }
                    V3 v = interp_cell(rx, ry, rz, vp[0], vp[4], vp[3], vp[7], vp[1], vp[5], vp[2], vp[6], vd[0], vd[3],
                                             vd[1], vd[2], vd[4], vd[7], vd[5], vd[6], vd[8], vd[11], vd[9], vd[10], vf[4],
                                             vf[5], vf[2], vf[3], vf[0], vf[1]);
                    if (sph) v = (mag3(vd[8]) / mag3(v)) * v;
                    const u32 idx = I4(ci, i, j, k);
                    cC[(size_t)idx * 3 + 0] = v.x; cC[(size_t)idx * 3 + 1] = v.y; cC[(size_t)idx * 3 + 2] = v.z;
                    cV[idx] *= wg[0][i] * wg[1][j] * wg[2][k] / 8;
                }
    }

    if (!corner_error.empty()) throw Error(corner_error);
    lap("node coordinates");

    // ---- face node maps and weights (dg.cpp:328-410) ----
    const int face_map[6] = {0, NPZ - 1, 0, NPY - 1, 0, NPX - 1};
#pragma omp parallel for schedule(static)
    for (int64_t cs = 0; cs < (int64_t)nBCS; cs++) {
        const u32 ci = (u32)cs;
        for (u32 q = t.cellStart[ci]; q < t.cellStart[ci + 1]; q++) {
            const u32 fi = t.cellFaces[q];
            const int face_o = (int)t.cellFaceID[q];
            const u32 cj = t.FNC[fi];
            if (cj == ci) continue;
            int face_n = face_o ^ 1;
            if (cj < nBCS)
                for (u32 r = t.cellStart[cj]; r < t.cellStart[cj + 1]; r++)
                    if (t.cellFaces[r] == fi) { face_n = (int)t.cellFaceID[r]; break; }
            const int vo = face_map[face_o], vn = face_map[face_n];
            const int na = (face_o < 4) ? NPX : NPY;
            const int nb = (face_o < 2) ? NPY : NPZ;
            const double* wa = (face_o < 4) ? wg[0] : wg[1];
            const double* wb = (face_o < 2) ? wg[1] : wg[2];
            for (int a = 0; a < na; a++)
                for (int bb = 0; bb < nb; bb++) {
                    const double wgt = wa[a] * wb[bb] / 4;
                    const size_t indf = (size_t)fi * NPF + (size_t)a * nb + bb;
                    const u32 index0 = (face_o < 2) ? I4(ci, a, bb, vo) : (face_o < 4 ? I4(ci, a, vo, bb) : I4(ci, vo, a, bb));
                    const u32 index1 = (face_n < 2) ? I4(cj, a, bb, vn) : (face_n < 4 ? I4(cj, a, vn, bb) : I4(cj, vn, a, bb));
                    FO[indf] = index0;
                    FN[indf] = index1;
                    if (index1 >= gBCSfield) {
                        for (int d = 0; d < 3; d++) cC[(size_t)index1 * 3 + d] = cC[(size_t)index0 * 3 + d];
                        cV[index1] = cV[index0];
                    }
                    const u32 src = (t.FMC[fi] <= 1) ? index0 : index1;
                    for (int d = 0; d < 3; d++) {
                        fC[indf * 3 + d] = cC[(size_t)src * 3 + d];
                        fN[indf * 3 + d] *= wgt;
                    }
                }
        }
    }

    lap("face node maps");
    // ---- Jinv (dg.cpp:413-476), AoS XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX ----
    Jinv.assign(gBCSfield * 9, 0.0);
    const double* D[3] = {b.dpsi[0].data(), b.dpsi[1].data(), b.dpsi[2].data()};
    static const int rm[9] = {0, 4, 8, 1, 5, 2, 3, 7, 6};    // AoS component -> row-major a*3+d
#pragma omp parallel for schedule(static)
    for (int64_t cs = 0; cs < (int64_t)nBCS; cs++) {
        const u32 ci = (u32)cs;
        for (int ii = 0; ii < NPX; ii++)
            for (int jj = 0; jj < NPY; jj++)
                for (int kk = 0; kk < NPZ; kk++) {
                    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // J[a][d] += x_m[a] * dpsi_d
                    for (int i = 0; i < NPX; i++) {
                        const double* x = &cC[(size_t)I4(ci, i, jj, kk) * 3];
                        for (int a = 0; a < 3; a++) J[a * 3 + 0] += x[a] * D[0][ii * NPX + i];
                        if (i == ii)
                            for (int a = 0; a < 3; a++) {
                                J[a * 3 + 1] += x[a] * D[1][jj * NPY + jj];
                                J[a * 3 + 2] += x[a] * D[2][kk * NPZ + kk];
                            }
                    }
                    for (int j = 0; j < NPY; j++)
                        if (j != jj) {
                            const double* x = &cC[(size_t)I4(ci, ii, j, kk) * 3];
                            for (int a = 0; a < 3; a++) J[a * 3 + 1] += x[a] * D[1][jj * NPY + j];
                        }
                    for (int k = 0; k < NPZ; k++)
                        if (k != kk) {
                            const double* x = &cC[(size_t)I4(ci, ii, jj, k) * 3];
                            for (int a = 0; a < 3; a++) J[a * 3 + 2] += x[a] * D[2][kk * NPZ + k];
                        }
                    double JT[9], A[9], Ai[9], R[9];
                    for (int a = 0; a < 3; a++) for (int d = 0; d < 3; d++) JT[a * 3 + d] = J[d * 3 + a];
                    mul33(JT, J, A);
                    if (NPX == 1) A[0] = 1;
                    if (NPY == 1) A[4] = 1;
                    if (NPZ == 1) A[8] = 1;
                    inv33(A, Ai);
                    if (NPX == 1) Ai[0] = 0;
                    if (NPY == 1) Ai[4] = 0;
                    if (NPZ == 1) Ai[8] = 0;
                    mul33(Ai, JT, R);
                    double* out = &Jinv[(size_t)I4(ci, ii, jj, kk) * 9];
                    for (int c = 0; c < 9; c++) {
                        const int a = rm[c] / 3, d = rm[c] % 3;
                        out[c] = R[d * 3 + a];     // transpose
                    }
                }
    }

    lap("Jinv");
    // ---- fI (field.cpp:257-270): 0 on physical boundary faces, 0.5 elsewhere (ghost faces of other ranks too) ----
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)nfn; k++) fI[k] = (FN[k] >= gBCSfield) ? 0.0 : 0.5;
    for (const auto& kv : t.boundaries)
        if (kv.first.find("interMesh") != std::string::npos)
            for (u32 f : kv.second)
                for (int n = 0; n < NPF; n++)
                    if (FO[(size_t)f * NPF + n] < gALL) fI[(size_t)f * NPF + n] = 0.5;   // isGhostFace
}

nsem_mesh Geometry::as_c() const {
    nsem_mesh m;
    std::memset(&m, 0, sizeof m);
    m.n_cells_real = nBCS; m.n_cells_all = nCells; m.n_faces = nFacets;
    m.cV = cV.data(); m.Jinv = Jinv.data(); m.fN = fN.data(); m.fI = fI.data(); m.face_normal = faceNormal.data();
    m.FO = FO.data(); m.FN = FN.data(); m.face_begin = faceBegin.data(); m.face_end = faceEnd.data();
    m.all_faces = allFaces.data(); m.face_id = faceID.data(); m.face_owner = faceOwner.data();
    m.face_neigh = faceNeigh.data(); m.face_mortar = faceMortar.data();
    m.cC = cC.data(); m.face_center = faceCenter.data();
    for (int q = 0; q < 6; q++) { m.psi_ref[q] = psiRef[q].data(); m.psi_cor[q] = psiCor[q].data(); }
    return m;
}

}  // namespace nsemh
