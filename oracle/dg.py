"""CPU restatement of the reference's DG basis tables and node geometry -- TEST INFRASTRUCTURE ONLY.

Follows (all in /root/reference):
  legendre / legendre_gauss_lobatto    src/field/dg.cpp:32-99
  lagrange_basis / _derivative         src/field/dg.cpp:103-143
  init_poly                            src/field/dg.cpp:147-163
  init_geom                            src/field/dg.cpp:167-477
  init_basis (psiRef / psiCor)         src/field/dg.cpp:481-590
  Interpolate_face / Interpolate_cell  src/tensor/tensor.h:522-571
  initGeomMeshFields (fI rule)         src/field/field.cpp:171-280
  matinv/matmul/mattrn                 src/tensor/tensor.cpp:170-243
"""
from __future__ import annotations

import math

import numpy as np

from .mesh import MeshTopo

# tensor component order XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX (tensor.h:452-454) -> (row, col)
T9 = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2), (1, 0), (2, 1), (2, 0)]
T9_FLAT = np.array([r * 3 + c for r, c in T9])          # AoS9 index -> row-major 3x3 index
T9_INV = np.argsort(T9_FLAT)                             # row-major 3x3 index -> AoS9 index


def legendre(p: int, x: float):
    L1 = L1_1 = L1_2 = 0.0
    L0, L0_1, L0_2 = 1.0, 0.0, 0.0
    for i in range(1, p + 1):
        L2, L2_1, L2_2 = L1, L1_1, L1_2
        L1, L1_1, L1_2 = L0, L0_1, L0_2
        a = (2 * i - 1.0) / i
        b = (i - 1.0) / i
        L0 = a * x * L1 - b * L2
        L0_1 = a * (L1 + x * L1_1) - b * L2_1
        L0_2 = a * (2 * L1_1 + x * L1_2) - b * L2_2
    return L0, L0_1, L0_2


def lgl(N: int):
    xgl = np.zeros(N)
    wgl = np.zeros(N)
    if N == 1:
        xgl[0], wgl[0] = 0.0, 2.0
        return xgl, wgl
    p = N - 1
    ph = N // 2
    for i in range(ph):
        x = math.cos((2 * i + 1) * math.pi / (2 * N))
        for _ in range(20):
            L0, L0_1, L0_2 = legendre(p, x)
            dx = -(1 - x * x) * L0_1 / (-2 * x * L0_1 + (1 - x * x) * L0_2)
            x += dx
            if abs(dx) < 1.0e-20:
                break
        xgl[p - i] = x
        wgl[p - i] = 2 / (p * (p + 1) * L0 * L0)
    if N != 2 * ph:
        L0, _, _ = legendre(p, 0.0)
        xgl[ph] = 0.0
        wgl[ph] = 2 / (p * (p + 1) * L0 * L0)
    for i in range(ph):
        xgl[i] = -xgl[p - i]
        wgl[i] = wgl[p - i]
    return xgl, wgl


def lagrange_basis(xgl, xs):
    """psi[j*Ns+s] = l_j(xs[s])"""
    N, Ns = len(xgl), len(xs)
    psi = np.zeros(N * Ns)
    for s in range(Ns):
        x = xs[s]
        for j in range(N):
            prod = 1.0
            for k in range(N):
                if k != j:
                    prod *= (x - xgl[k]) / (xgl[j] - xgl[k])
            psi[j * Ns + s] = prod
    return psi


def lagrange_basis_derivative(xgl, xs):
    """dpsi[s*N+i] = l_i'(xs[s])"""
    N, Ns = len(xgl), len(xs)
    d = np.zeros(N * Ns)
    for s in range(Ns):
        x = xs[s]
        for i in range(N):
            acc = 0.0
            for j in range(N):
                if i != j:
                    prod = 1.0
                    for k in range(N):
                        if k != i and k != j:
                            prod *= (x - xgl[k]) / (xgl[i] - xgl[k])
                    acc += prod / (xgl[i] - xgl[j])
            d[s * N + i] = acc
    return d


def _matinv(A_):
    """Gauss-Jordan elimination exactly as the reference does it (tensor.cpp:170-215): one bubble pass on column 0,
    elimination without further pivoting, final row scaling."""
    N = A_.shape[0]
    A = np.array(A_, dtype=float).copy()
    X = np.eye(N)
    for i in range(N - 1, 0, -1):
        if A[i - 1, 0] < A[i, 0]:
            A[[i, i - 1]] = A[[i - 1, i]]
            X[[i, i - 1]] = X[[i - 1, i]]
    for i in range(N):
        for j in range(N):
            if j != i:
                temp = A[j, i] / A[i, i]
                for k in range(N):
                    A[j, k] -= A[i, k] * temp
                    X[j, k] -= X[i, k] * temp
    for i in range(N):
        temp = A[i, i]
        for j in range(N):
            A[i, j] = A[i, j] / temp
            X[i, j] = X[i, j] / temp
    return X


def _matmul(A, B):
    """matmul of tensor.cpp:218-228: sequential sums."""
    N = A.shape[0]
    X = np.zeros((N, N))
    for i in range(N):
        for j in range(N):
            acc = 0.0
            for k in range(N):
                acc += A[i, k] * B[k, j]
            X[i, j] = acc
    return X


class Basis:
    """DG::init_poly + DG::init_basis for orders (npx,npy,npz) = polynomial degree per direction."""

    def __init__(self, nop):
        self.NPX, self.NPY, self.NPZ = (int(nop[0]) + 1, int(nop[1]) + 1, int(nop[2]) + 1)
        NPX, NPY, NPZ = self.NPX, self.NPY, self.NPZ
        self.NP = NPX * NPY * NPZ
        if NPX <= NPY and NPX <= NPZ:
            self.NPF = NPY * NPZ
        elif NPY <= NPX and NPY <= NPZ:
            self.NPF = NPX * NPZ
        else:
            self.NPF = NPX * NPY
        self.n = [NPX, NPY, NPZ]
        self.xgl, self.wgl, self.psi, self.dpsi = [], [], [], []
        self.psiRef = [None] * 6
        self.psiCor = [None] * 6
        for d in range(3):
            ngl = self.n[d]
            ngle = ngl + 1
            x, w = lgl(ngl)
            xe, we = lgl(ngle)
            self.xgl.append(x)
            self.wgl.append(w)
            self.psi.append(lagrange_basis(x, x))
            self.dpsi.append(lagrange_basis_derivative(x, x))
            if ngl == 1:
                xre = [np.zeros(ngle), np.zeros(ngle)]
            else:
                xre = [-0.5 + xe / 2, 0.5 + xe / 2]
            psire = [lagrange_basis(x, xre[c]).reshape(ngl, ngle) for c in range(2)]
            psie = lagrange_basis(x, xe).reshape(ngl, ngle)
            # mass matrices with the (ngl+1)-point rule, accumulated in the reference's loop order (dg.cpp:558-573)
            Mcc = np.zeros((ngl, ngl))
            Msc = [np.zeros((ngl, ngl)), np.zeros((ngl, ngl))]
            Mga = [np.zeros((ngl, ngl)), np.zeros((ngl, ngl))]
            for j in range(ngl):
                for k in range(ngl):
                    for q in range(ngle):
                        Mcc[j, k] += (we[q] / 2) * psie[j, q] * psie[k, q]
                        for c in range(2):
                            vs = (we[q] / 2) * psie[j, q] * psire[c][k, q]
                            Msc[c][j, k] += vs
                            Mga[c][k, j] += vs
            iMcc = _matinv(Mcc)
            for c in range(2):
                self.psiRef[d * 2 + c] = _matmul(iMcc, Msc[c]).T.reshape(-1).copy()
                self.psiCor[d * 2 + c] = _matmul(iMcc, Mga[c]).T.reshape(-1).copy()

    def D(self, d):
        """D[s, i] = l_i'(x_s) as an (n,n) matrix."""
        n = self.n[d]
        return self.dpsi[d].reshape(n, n)


def interpolate_face(r, s, x00, x01, x10, x11, xr0, xr1, x0s, x1s):
    return (-(1.0 - r) * (1.0 - s) * x00 + (1.0 - r) * x0s - (1.0 - r) * s * x01 + (1.0 - s) * xr0
            + s * xr1 - r * (1.0 - s) * x10 + r * x1s - r * s * x11)


def interpolate_cell(r, s, t, x000, x001, x010, x011, x100, x101, x110, x111,
                     xr00, xr01, xr10, xr11, x0s0, x0s1, x1s0, x1s1,
                     x00t, x01t, x10t, x11t, x0st, x1st, xr0t, xr1t, xrs0, xrs1):
    return ((1.0 - r) * (1.0 - s) * (1.0 - t) * x000 - (1.0 - r) * (1.0 - s) * x00t + (1.0 - r) * (1.0 - s) * t * x001
            - (1.0 - r) * (1.0 - t) * x0s0 + (1.0 - r) * x0st - (1.0 - r) * t * x0s1
            + (1.0 - r) * s * (1.0 - t) * x010 - (1.0 - r) * s * x01t + (1.0 - r) * s * t * x011
            - (1.0 - s) * (1.0 - t) * xr00 + (1.0 - s) * xr0t - (1.0 - s) * t * xr01
            + (1.0 - t) * xrs0 + t * xrs1
            - s * (1.0 - t) * xr10 + s * xr1t - s * t * xr11
            + r * (1.0 - s) * (1.0 - t) * x100 - r * (1.0 - s) * x10t + r * (1.0 - s) * t * x101
            - r * (1.0 - t) * x1s0 + r * x1st - r * t * x1s1
            + r * s * (1.0 - t) * x110 - r * s * x11t + r * s * t * x111)


class Geometry:
    """Node-level geometry arrays of Mesh::initGeomMeshFields + DG::init_geom.

    Arrays (names as in the reference):
      cC (gALL,3), cV (gALL), Jinv (gBCSfield,9 in tensor.h component order), FO/FN (nF*NPF, sentinel = gALL),
      fN (nF*NPF,3), fC (nF*NPF,3), fI (nF*NPF), faceIndices (2,nCells), allFaces
    """

    def __init__(self, topo: MeshTopo, basis: Basis):
        self.topo, self.basis = topo, basis
        b = basis
        NPX, NPY, NPZ, NP, NPF = b.NPX, b.NPY, b.NPZ, b.NP, b.NPF
        nC, nF, nB = len(topo.cells), len(topo.facets), topo.nBCS
        self.nCells, self.nFacets, self.nBCS = nC, nF, nB
        self.gBCSfield, self.gALL = nB * NP, nC * NP
        gALL = self.gALL

        def I4(c, i, j, k):
            return c * NP + i * NPY * NPZ + j * NPZ + k

        # allFaces / faceIndices (field.cpp:178-193)
        fi0, fi1, allf = [], [], []
        for c in topo.cells:
            fi0.append(len(allf))
            allf.extend(c)
            fi1.append(len(allf))
        self.faceIndices = np.array([fi0, fi1], dtype=np.int64)
        self.allFaces = np.array(allf, dtype=np.int64)

        cC = np.repeat(topo.CC, NP, axis=0)
        cV = np.repeat(topo.CV, NP)
        fC = np.repeat(topo.FC, NPF, axis=0)
        fN = np.repeat(topo.FNv, NPF, axis=0)
        FO = np.full(nF * NPF, gALL, dtype=np.int64)
        FN = np.full(nF * NPF, gALL, dtype=np.int64)
        V = topo.V
        xgl, wgl = b.xgl, b.wgl
        sph = bool(getattr(topo, "spherical", False))
        self.spherical, self.sphere_radius = sph, float(getattr(topo, "sphere_radius", 0.0))

        def mag3(a):    # Unroll<3>::dot nests to the right
            return math.sqrt(float(a[0] * a[0] + (a[1] * a[1] + a[2] * a[2])))

        # ---- node coordinates by transfinite interpolation (dg.cpp:176-325) ----
        sides = [(0, 1), (3, 2), (7, 6), (4, 5), (0, 3), (1, 2), (5, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
        self.corners = np.zeros((nB, 8), dtype=np.int64)
        for ci in range(nB):
            c = topo.cells[ci]
            fids = topo.faceID[ci]
            id0 = fids[0]
            id1 = id0 ^ 1
            f1, f2 = None, None
            for i, fidx in enumerate(c):
                f = topo.facets[fidx]
                if fids[i] == id0:
                    f1 = list(f) if f1 is None else topo.merge_facets(f1, f)
                elif fids[i] == id1:
                    f2 = list(f) if f2 is None else topo.merge_facets(f2, f)
            vpi = topo.hex_corners(f1, f2)
            if id0 == 2:
                order = [0, 1, 5, 4, 3, 2, 6, 7]
            elif id0 == 4:
                order = [0, 3, 7, 4, 1, 2, 6, 5]
            else:
                order = list(range(8))
            vidx = [0] * 8
            for i in range(8):
                vidx[order[i]] = vpi[i]
            self.corners[ci] = vidx
            vp = [V[v] for v in vidx]
            ev = [(vp[a], vp[b_]) for a, b_ in sides]
            for i in range(NPX):
                rx = (xgl[0][i] + 1) / 2
                for j in range(NPY):
                    ry = (xgl[1][j] + 1) / 2
                    for k in range(NPZ):
                        rz = (xgl[2][k] + 1) / 2
                        m = [rx] * 4 + [ry] * 4 + [rz] * 4
                        vd = [(1 - m[w]) * ev[w][0] + m[w] * ev[w][1] for w in range(12)]
                        if sph:     # every blended point goes back to the blended radius (ADDV/ADDF/ADDC, dg.cpp:257-285)
                            vd = [vd[w] * (((1 - m[w]) * mag3(ev[w][0]) + m[w] * mag3(ev[w][1])) / mag3(vd[w])) for w in range(12)]
                        vf = [None] * 6

                        def addf(rr, rs, i00, i01, i10, i11, ir0, ir1, i0s, i1s):
                            x = interpolate_face(rr, rs, vp[i00], vp[i01], vp[i10], vp[i11],
                                                 vd[ir0], vd[ir1], vd[i0s], vd[i1s])
                            return (mag3(vd[ir0]) / mag3(x)) * x if sph else x
                        vf[0] = addf(rx, ry, 0, 3, 1, 2, 0, 1, 4, 5)
                        vf[1] = addf(rx, ry, 4, 7, 5, 6, 3, 2, 7, 6)
                        vf[2] = addf(rx, rz, 0, 4, 1, 5, 0, 3, 8, 9)
                        vf[3] = addf(rx, rz, 3, 7, 2, 6, 1, 2, 11, 10)
                        vf[4] = addf(ry, rz, 0, 4, 3, 7, 4, 7, 8, 11)
                        vf[5] = addf(ry, rz, 1, 5, 2, 6, 5, 6, 9, 10)
                        v = interpolate_cell(rx, ry, rz,
                                             vp[0], vp[4], vp[3], vp[7], vp[1], vp[5], vp[2], vp[6],
                                             vd[0], vd[3], vd[1], vd[2], vd[4], vd[7], vd[5], vd[6],
                                             vd[8], vd[11], vd[9], vd[10],
                                             vf[4], vf[5], vf[2], vf[3], vf[0], vf[1])
                        if sph:
                            v = (mag3(vd[8]) / mag3(v)) * v
                        idx = I4(ci, i, j, k)
                        cC[idx] = v
                        cV[idx] *= wgl[0][i] * wgl[1][j] * wgl[2][k] / 8

        # ---- face node maps and weights (dg.cpp:328-410) ----
        face_map = [0, NPZ - 1, 0, NPY - 1, 0, NPX - 1]
        for ci in range(nB):
            c = topo.cells[ci]
            ids = topo.faceID[ci]
            for mm, fi in enumerate(c):
                face_o = ids[mm]
                cj = topo.FNC[fi]
                fm = topo.FMC[fi]
                if cj == ci:
                    continue
                face_n = face_o ^ 1
                if cj < nB:
                    cn = topo.cells[cj]
                    for i2, f2i in enumerate(cn):
                        if f2i == fi:
                            face_n = topo.faceID[cj][i2]
                            break
                vo, vn = face_map[face_o], face_map[face_n]
                if face_o in (0, 1):
                    loops = [(a, b_, wgl[0][a] * wgl[1][b_] / 4, a * NPY + b_, I4(ci, a, b_, vo))
                             for a in range(NPX) for b_ in range(NPY)]
                elif face_o in (2, 3):
                    loops = [(a, b_, wgl[0][a] * wgl[2][b_] / 4, a * NPZ + b_, I4(ci, a, vo, b_))
                             for a in range(NPX) for b_ in range(NPZ)]
                else:
                    loops = [(a, b_, wgl[1][a] * wgl[2][b_] / 4, a * NPZ + b_, I4(ci, vo, a, b_))
                             for a in range(NPY) for b_ in range(NPZ)]
                for a, b_, wgt, off, index0 in loops:
                    indf = fi * NPF + off
                    if face_n in (0, 1):
                        index1 = I4(cj, a, b_, vn)
                    elif face_n in (2, 3):
                        index1 = I4(cj, a, vn, b_)
                    else:
                        index1 = I4(cj, vn, a, b_)
                    FO[indf] = index0
                    FN[indf] = index1
                    if index1 >= self.gBCSfield:
                        cC[index1] = cC[index0]
                        cV[index1] = cV[index0]
                    fC[indf] = cC[index0] if fm <= 1 else cC[index1]
                    fN[indf] = fN[indf] * wgt

        # ---- Jinv (dg.cpp:413-476), same accumulation and product order as the reference ----
        D = [b.D(0), b.D(1), b.D(2)]
        X = cC[: self.gBCSfield].reshape(nB, NPX, NPY, NPZ, 3)
        J = np.zeros((nB, NPX, NPY, NPZ, 3, 3))          # J[a,d] = sum_m x_m[a]*dpsi_d ; mul(cC[index1], dpsi_ij)
        for ii in range(NPX):
            for jj in range(NPY):
                for kk in range(NPZ):
                    Jq = J[:, ii, jj, kk]
                    for i in range(NPX):
                        Jq[:, :, 0] += X[:, i, jj, kk] * D[0][ii, i]
                        if i == ii:
                            Jq[:, :, 1] += X[:, i, jj, kk] * D[1][jj, jj]
                            Jq[:, :, 2] += X[:, i, jj, kk] * D[2][kk, kk]
                    for j in range(NPY):
                        if j != jj:
                            Jq[:, :, 1] += X[:, ii, j, kk] * D[1][jj, j]
                    for k in range(NPZ):
                        if k != kk:
                            Jq[:, :, 2] += X[:, ii, jj, k] * D[2][kk, k]
        J = J.reshape(-1, 3, 3)

        def mul33(p, q):       # mul(Tensor,Tensor), tensor.cpp:42-58
            r = np.empty_like(p)
            for a in range(3):
                for c in range(3):
                    r[:, a, c] = (p[:, a, 0] * q[:, 0, c] + p[:, a, 1] * q[:, 1, c]) + p[:, a, 2] * q[:, 2, c]
            return r

        def inv33(p):          # inv(Tensor), tensor.cpp:152-168
            r = np.empty_like(p)
            r[:, 0, 0] = p[:, 1, 1] * p[:, 2, 2] - p[:, 1, 2] * p[:, 2, 1]
            r[:, 1, 1] = p[:, 0, 0] * p[:, 2, 2] - p[:, 0, 2] * p[:, 2, 0]
            r[:, 2, 2] = p[:, 0, 0] * p[:, 1, 1] - p[:, 0, 1] * p[:, 1, 0]
            r[:, 0, 1] = p[:, 0, 2] * p[:, 2, 1] - p[:, 0, 1] * p[:, 2, 2]
            r[:, 0, 2] = p[:, 0, 1] * p[:, 1, 2] - p[:, 0, 2] * p[:, 1, 1]
            r[:, 1, 0] = p[:, 1, 2] * p[:, 2, 0] - p[:, 1, 0] * p[:, 2, 2]
            r[:, 1, 2] = p[:, 0, 2] * p[:, 1, 0] - p[:, 0, 0] * p[:, 1, 2]
            r[:, 2, 0] = p[:, 1, 0] * p[:, 2, 1] - p[:, 1, 1] * p[:, 2, 0]
            r[:, 2, 1] = p[:, 0, 1] * p[:, 2, 0] - p[:, 0, 0] * p[:, 2, 1]
            d = (p[:, 0, 0] * r[:, 0, 0] + p[:, 0, 1] * r[:, 1, 0]) + p[:, 0, 2] * r[:, 2, 0]
            return r / d[:, None, None]

        JT = np.swapaxes(J, 1, 2).copy()
        A = mul33(JT, J)
        if NPX == 1:
            A[:, 0, 0] = 1
        if NPY == 1:
            A[:, 1, 1] = 1
        if NPZ == 1:
            A[:, 2, 2] = 1
        A = inv33(A)
        if NPX == 1:
            A[:, 0, 0] = 0
        if NPY == 1:
            A[:, 1, 1] = 0
        if NPZ == 1:
            A[:, 2, 2] = 0
        Ji = np.swapaxes(mul33(A, JT), 1, 2).copy()
        self.Jinv33 = Ji                                           # (gBCSfield,3,3) row-major [a,d]
        self.Jinv = Ji.reshape(-1, 9)[:, T9_FLAT]                  # reference AoS component order

        # ---- fI (field.cpp:257-270); single rank: no ghost faces ----
        fI = np.where(FN >= self.gBCSfield, 0.0, 0.5)
        self.cC, self.cV, self.fC, self.fN, self.FO, self.FN, self.fI = cC, cV, fC, fN, FO, FN, fI


def load_geometry(grid, nop):
    topo = MeshTopo(grid).load()
    basis = Basis(nop)
    return Geometry(topo, basis)
