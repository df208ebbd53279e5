"""CPU restatement (numpy) of the reference's explicit dGSEM Euler step -- TEST INFRASTRUCTURE ONLY.

This is the parity ORACLE for the CUDA path.  It is never imported by the product package
(nebulasem_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.

Pinned: tests/test_oracle_vs_reference.py checks this file against dumps produced by the UNMODIFIED
reference (oracle/_ref/parity/euler, built by oracle/build_ref.sh) and against the committed golden
fixtures in tests/golden/ generated from the same binary (tests/golden/make_golden.py).

Follows (all in /root/reference):
  time loop                 apps/euler/euler.cpp:179-287
  set-up of rho/p/refs      apps/euler/euler.cpp:58-176
  cds / rusanov             src/field/field.h:2881-2943
  scatter/gather (conf.)    src/field/field.h:2019-2033, 2135-2149
  div_flux / grad_flux      src/field/field.h:3051-3117
  gradf<strong>             src/field/field.h:3328-3362
  divf<weak>                src/field/field.h:3417-3478
  fillBCs                   src/field/field.h:2731-2776
  applyExplicitBCs          src/field/field.h:2586-2727
  src / ddt / addTemporal   src/field/field.h:3712-3723, 3741-3841, 3875-3920
  SolveTexplicit            src/solvers/solve.cpp:563-581
  sym() (SYMMETRY BC)       src/tensor/tensor.h:486-494
  tensor dot/mul ordering   src/tensor/tensor.cpp:14-29,72-78 ; tensor.h:124-127 (Unroll::dot is right-nested)

`exact_order=True` reproduces the reference's floating-point operation ORDER (per-cell face-ID order in
div_flux/grad_flux, q-lexicographic scatter order in the volume loops, expression-template evaluation
order); with libm from oracle/libm.py the result is then bit-identical to the -O2 -ffp-contract=off build
of the reference on the same machine.  `exact_order=False` uses einsum contractions (same maths, rounding
differs at 1e-16) and is ~50x faster.
"""
from __future__ import annotations

from dataclasses import dataclass, field as dfield

import numpy as np

from . import libm
from .dg import Geometry
from .mesh import equal


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------
@dataclass
class Params:
    """general{} + euler{} control values that reach the hot path (properties.cpp:14-34, euler.cpp:19-48)."""
    viscosity: float = 1.568e-5
    Pr: float = 0.9
    T0: float = 300.0
    P0: float = 101325.0
    cp: float = 1004.67
    cv: float = 715.5
    dt: float = 0.1
    gravity: tuple = (0.0, 0.0, -9.860616)     # Controls::gravity default, field.cpp:73
    buoyancy: bool = True
    diffusion: bool = True
    time_scheme: str = "BDF1"                  # BDF1 | AB1 | RK1..RK4 (all one forward-Euler stage on this path)
    problem_init: str = "NONE"
    convection_scheme: str = "RUSANOV"         # Controls::convection_scheme / blend_factor (field.cpp:520-527): read by the convection app's divf
    blend_factor: float = 0.2
    is_spherical: bool = False                 # Mesh::is_spherical / sphere_radius / sphere_height, mesh.cpp:31-33
    sphere_radius: float = 6371220.0
    sphere_height: float = 10000.0

    @staticmethod
    def from_controls(blocks: dict) -> "Params":
        g = blocks.get("general", {})
        e = blocks.get("euler", {})
        p = Params()

        def f(blk, key, cur):
            return float(blk[key][0]) if key in blk else cur
        p.viscosity = f(g, "viscosity", p.viscosity)
        p.Pr = f(g, "Pr", p.Pr)
        p.T0 = f(g, "T0", p.T0)
        p.P0 = f(g, "P0", p.P0)
        p.cp = f(g, "cp", p.cp)
        p.cv = f(g, "cv", p.cv)
        p.dt = f(g, "dt", p.dt)
        if "gravity" in g:
            p.gravity = tuple(float(x) for x in g["gravity"][:3])
        if "time_scheme" in g:
            p.time_scheme = g["time_scheme"][0]
        yes = lambda s: s.upper() in ("YES", "1", "TRUE")  # noqa: E731
        if "buoyancy" in e:
            p.buoyancy = yes(e["buoyancy"][0])
        if "diffusion" in e:
            p.diffusion = yes(e["diffusion"][0])
        if "problem_init" in e:
            p.problem_init = e["problem_init"][0]
        if "convection_scheme" in g:
            p.convection_scheme = g["convection_scheme"][0]
        p.blend_factor = f(g, "blend_factor", p.blend_factor)
        if "is_spherical" in g:
            p.is_spherical = yes(g["is_spherical"][0])
        p.sphere_radius = f(g, "sphere_radius", p.sphere_radius)
        p.sphere_height = f(g, "sphere_height", p.sphere_height)
        return p


@dataclass
class BCSpec:
    """A boundary condition bound to face lists (BCondition<T>, field.h:144-173)."""
    kind: str
    faces: np.ndarray
    value: np.ndarray
    neighbor_faces: np.ndarray | None = None
    shape: float = 0.0
    tvalue: np.ndarray | None = None
    tshape: float = 0.0
    zMin: float = 0.0
    fixed: np.ndarray | None = None            # frozen CALC_DIRICHLET values (nfaces*NPF, comps)


# ------------------------------------------------------------------------------------------------
# small tensor helpers with the reference's association order
# ------------------------------------------------------------------------------------------------
def vdot(a, b):
    """dot(Vector,Vector): a0*b0 + (a1*b1 + a2*b2)   (Unroll<3>::dot, tensor.h:124-127)"""
    return a[..., 0] * b[..., 0] + (a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2])


def vmag(a):
    return libm.sqrt_(vdot(a, a))


def tdotv(T, v):
    """dot(Tensor,Vector)_a = (T[a,0]*v0 + T[a,1]*v1) + T[a,2]*v2   (tensor.cpp:72-78); T is (...,3,3) row-major"""
    return (T[..., :, 0] * v[..., None, 0] + T[..., :, 1] * v[..., None, 1]) + T[..., :, 2] * v[..., None, 2]


def sym_vec(p, n):
    """sym(Vector p, Vector n), tensor.h:486-494 (A is an STensor: I - en en)."""
    en = n / vmag(n)[..., None]
    Axx = 1.0 - en[..., 0] * en[..., 0]
    Ayy = 1.0 - en[..., 1] * en[..., 1]
    Azz = 1.0 - en[..., 2] * en[..., 2]
    Axy = 0.0 - en[..., 0] * en[..., 1]
    Ayz = 0.0 - en[..., 1] * en[..., 2]
    Axz = 0.0 - en[..., 0] * en[..., 2]
    r = np.empty_like(p)
    # dot(STensor,Vector), tensor.cpp:80-86
    r[..., 0] = Axx * p[..., 0] + Axy * p[..., 1] + Axz * p[..., 2]
    r[..., 1] = Axy * p[..., 0] + Ayy * p[..., 1] + Ayz * p[..., 2]
    r[..., 2] = Axz * p[..., 0] + Ayz * p[..., 1] + Azz * p[..., 2]
    magR = vmag(r)
    magP = vmag(p)
    out = r.copy()
    for i in range(len(magR)):
        if not equal(float(magR[i]), 0.0):
            out[i] = r[i] * (magP[i] / magR[i])
    return out


# ------------------------------------------------------------------------------------------------
# the solver state + step
# ------------------------------------------------------------------------------------------------
class EulerOracle:
    """State and one-step update on a Geometry (single rank)."""

    def __init__(self, geo: Geometry, params: Params, exact_order: bool = True):
        self.g = geo
        self.p = params
        self.exact = exact_order
        b = geo.basis
        self.NP, self.NPF = b.NP, b.NPF
        self.shape = (b.NPX, b.NPY, b.NPZ)
        self.nB = geo.nBCS
        self.gB, self.gA = geo.gBCSfield, geo.gALL
        self.valid = geo.FO < self.gA                      # used face-node slots
        self.kv = np.nonzero(self.valid)[0]
        self.FOv = geo.FO[self.kv]
        self.FNv = geo.FN[self.kv]
        self.fNv = geo.fN[self.kv]
        self.fIv = geo.fI[self.kv]
        self.unit_fN = self.fNv / vmag(self.fNv)[:, None]
        # Jin = Jinv * cV (field.h:3341,3448), row-major [a,d]
        self.Jin = geo.Jinv33 * geo.cV[: self.gB, None, None]
        self.D = [b.D(0), b.D(1), b.D(2)]
        self._build_face_tables()
        self._build_mortar_tables()
        self.bcs: dict[str, list[BCSpec]] = {"rho": [], "U": [], "T": [], "p": [], "p_ref": [], "rho_ref": [], "g": []}
        self.step_count = 0

    # ---- per-cell face accumulation tables (order of div_flux/grad_flux, field.h:3058-3114) ----
    def _build_face_tables(self):
        g = self.g
        NP, NPF, gA = self.NP, self.NPF, self.gA
        nfmax = int((g.faceIndices[1][: self.nB] - g.faceIndices[0][: self.nB]).max())
        T = nfmax * NPF
        tgt = np.full((self.nB, T), -1, dtype=np.int64)
        kid = np.zeros((self.nB, T), dtype=np.int64)
        own = np.zeros((self.nB, T), dtype=bool)
        for i in range(self.nB):
            t = 0
            for f in range(g.faceIndices[0][i], g.faceIndices[1][i]):
                face = g.allFaces[f]
                ks = face * NPF + np.arange(NPF)
                c1 = g.FO[ks]
                c2 = g.FN[ks]
                is_own = (c1 >= i * NP) & (c1 < (i + 1) * NP)
                is_nb = (~is_own) & (c2 < gA)
                sl = slice(t, t + NPF)
                tgt[i, sl] = np.where(is_own, c1, np.where(is_nb, c2, -1))
                kid[i, sl] = ks
                own[i, sl] = is_own
                t += NPF
        self.ft_tgt, self.ft_k, self.ft_own = tgt, kid, own
        # boundary cells -> (ghost node, owner node) pairs for fillBCs (field.h:2735-2743)
        gh, ow = [], []
        for i in range(self.nB, g.nCells):
            face = g.allFaces[g.faceIndices[0][i]]
            ks = face * NPF + np.arange(NPF)
            m = g.FN[ks] < gA
            gh.append(g.FN[ks][m])
            ow.append(g.FO[ks][m])
        self.fill_ghost = np.concatenate(gh) if gh else np.zeros(0, dtype=np.int64)
        self.fill_owner = np.concatenate(ow) if ow else np.zeros(0, dtype=np.int64)


    # ---- mortar (non-conforming, 2:1) faces: scatter/gather_non_conforming (field.h:2019-2248) -------------------
    def _build_mortar_tables(self):
        """Per mortar sub-face (gFMC >= 1): the coarse cell's face nodes, the half chosen along each face axis and the
        tensor-product weights psiRef (coarse -> fine trace, field.h:2198-2208) / psiCor (fine -> coarse flux, :2082-2092)."""
        g, t, b = self.g, self.g.topo, self.g.basis
        NPX, NPY, NPZ = self.shape
        NP, NPF = self.NP, self.NPF
        FMC = np.asarray(t.FMC)
        self.mortar_faces = np.nonzero(FMC >= 1)[0]
        self.has_mortar = len(self.mortar_faces) > 0
        if not self.has_mortar:
            return
        pos = np.full(len(g.FO), -1, dtype=np.int64)            # (face, slot) -> position in the used-slot arrays
        pos[self.kv] = np.arange(len(self.kv))
        n = [NPX, NPY, NPZ]

        def I4(c, i, j, k):
            return c * NP + i * NPY * NPZ + j * NPZ + k

        def dot3(a, bb):
            return (a[0] * bb[0] + a[1] * bb[1]) + a[2] * bb[2]

        M = []
        for fi in self.mortar_faces:
            fm = FMC[fi]
            co = t.FOC[fi] if fm == 1 else t.FNC[fi]
            cn = t.FNC[fi] if fm == 1 else t.FOC[fi]
            cell = t.cells[cn]
            ids = t.faceID[cn]
            fid = next(ids[j] for j in range(len(cell)) if cell[j] == fi)
            cco = np.asarray(t.FC[fi], dtype=float)
            ccn = np.zeros(3)
            nch = 0
            for j in range(len(cell)):
                if ids[j] == fid:
                    ccn = ccn + np.asarray(t.FC[cell[j]], dtype=float)
                    nch += 1
            ccn = ccn / float(nch)
            if fid in (0, 1):
                d1, d2 = 0, 1
                kf = 0 if fid == 0 else NPZ - 1
                node = lambda a, c2: I4(cn, a, c2, kf)
                v0, v1, v2 = g.cC[I4(cn, 0, 0, kf)], g.cC[I4(cn, NPX - 1, 0, kf)], g.cC[I4(cn, 0, NPY - 1, kf)]
            elif fid in (2, 3):
                d1, d2 = 0, 2
                jf = 0 if fid == 2 else NPY - 1
                node = lambda a, c2: I4(cn, a, jf, c2)
                v0, v1, v2 = g.cC[I4(cn, 0, jf, 0)], g.cC[I4(cn, NPX - 1, jf, 0)], g.cC[I4(cn, 0, jf, NPZ - 1)]
            else:
                d1, d2 = 1, 2
                i_f = 0 if fid == 4 else NPX - 1
                node = lambda a, c2: I4(cn, i_f, a, c2)
                v0, v1, v2 = g.cC[I4(cn, i_f, 0, 0)], g.cC[I4(cn, i_f, NPY - 1, 0)], g.cC[I4(cn, i_f, 0, NPZ - 1)]
            off1 = 0 if dot3(ccn - v0, v1 - v0) >= dot3(cco - v0, v1 - v0) else 1
            off2 = 0 if dot3(ccn - v0, v2 - v0) >= dot3(cco - v0, v2 - v0) else 1
            n1, n2 = n[d1], n[d2]
            slots = np.array([fi * NPF + a * n2 + c2 for a in range(n1) for c2 in range(n2)])      # face slots (a,b), b fastest
            nodes = np.array([node(a, c2) for a in range(n1) for c2 in range(n2)])                   # coarse face nodes, same order
            R1, R2 = b.psiRef[d1 * 2 + off1], b.psiRef[d2 * 2 + off2]
            C1, C2 = b.psiCor[d1 * 2 + off1], b.psiCor[d2 * 2 + off2]
            ns = n1 * n2
            # scatter: slot (ao,bo) <- sum over coarse (an,bn):  fx = R1[an*n1+ao], fy = R2[bn*n2+bo]
            sfx = np.array([[R1[an * n1 + ao] for an in range(n1) for bn in range(n2)] for ao in range(n1) for bo in range(n2)])
            sfy = np.array([[R2[bn * n2 + bo] for an in range(n1) for bn in range(n2)] for ao in range(n1) for bo in range(n2)])
            # gather: slot (an,bn) <- sum over fine (ao,bo):  fx = C1[ao*n1+an], fy = C2[bo*n2+bn]
            gfx = np.array([[C1[ao * n1 + an] for ao in range(n1) for bo in range(n2)] for an in range(n1) for bn in range(n2)])
            gfy = np.array([[C2[bo * n2 + bn] for ao in range(n1) for bo in range(n2)] for an in range(n1) for bn in range(n2)])
            assert (pos[slots] >= 0).all()
            M.append(dict(pos=pos[slots], nodes=nodes, sfx=sfx, sfy=sfy, gfx=gfx, gfy=gfy, to_N=(fm == 1), ns=ns))
        self.mortar = M

    def scatter(self, c):
        """(fFO, fFN) on the used slots: fFO = c[FO], fFN = c[FN]; on a mortar sub-face the coarse side's trace is the
        projection of the coarse cell's face values onto the sub-face (scatter_non_conforming, field.h:2135-2248)."""
        fO, fN = c[self.FOv], c[self.FNv]
        if self.has_mortar:
            fO, fN = fO.copy(), fN.copy()
            tail = (1,) * (c.ndim - 1)
            for m in self.mortar:
                ns = m["ns"]
                acc = np.zeros((ns,) + c.shape[1:])
                cv = c[m["nodes"]]                                    # coarse face values, term order (an,bn)
                for q in range(ns):
                    acc = acc + (cv[q] * m["sfx"][:, q].reshape((ns,) + tail)) * m["sfy"][:, q].reshape((ns,) + tail)
                (fN if m["to_N"] else fO)[m["pos"]] = acc
        return fO, fN

    def gather(self, fF):
        """(fFO, fFN) copies of a facet field on the used slots; on a mortar sub-face the coarse side receives the
        projection of the sub-face values onto the coarse face nodes (gather_non_conforming, field.h:2019-2132)."""
        if not self.has_mortar:
            return fF, fF
        fO, fN = fF.copy(), fF.copy()
        tail = (1,) * (fF.ndim - 1)
        for m in self.mortar:
            ns = m["ns"]
            acc = np.zeros((ns,) + fF.shape[1:])
            fv = fF[m["pos"]]                                         # sub-face values, term order (ao,bo)
            for q in range(ns):
                acc = acc + (fv[q] * m["gfx"][:, q].reshape((ns,) + tail)) * m["gfy"][:, q].reshape((ns,) + tail)
            (fN if m["to_N"] else fO)[m["pos"]] = acc
        return fO, fN

    # ---- face operators ----------------------------------------------------------------------
    def face_full(self, vals_valid, comps_shape=()):
        """expand values on used slots to the full (nF*NPF) facet array"""
        out = np.zeros((len(self.g.FO),) + comps_shape)
        out[self.kv] = vals_valid
        return out

    def cds(self, c):
        fi = self.fIv.reshape((-1,) + (1,) * (c.ndim - 1))
        fO, fN = self.scatter(c)
        return fO * fi + fN * (1 - fi)

    def rusanov(self, F, q, lam):
        """fF = cds(F) - mul(unit(fN), lam*(q_N - q_O))   (field.h:2928-2943)"""
        fF = self.cds(F)
        qO, qN = self.scatter(q)
        dq = qN - qO
        if q.ndim == 1:
            return fF - self.unit_fN * (lam * dq)[:, None]
        ldq = lam[:, None] * dq
        return fF - self.unit_fN[:, :, None] * ldq[:, None, :]

    def _accumulate_faces(self, r, contrib_o, contrib_n):
        """r[c1] += contrib_o[k] (cell owns the face node) else r[c2] -= contrib_n[k], in the reference's order."""
        if self.exact:
            tgt, kid, own = self.ft_tgt, self.ft_k, self.ft_own
            for t in range(tgt.shape[1]):
                sel = tgt[:, t] >= 0
                if not sel.any():
                    continue
                tg = tgt[sel, t]
                k = kid[sel, t]
                o = own[sel, t]
                o_b = o.reshape((-1,) + (1,) * (contrib_o.ndim - 1))
                r[tg] += np.where(o_b, contrib_o[k], -contrib_n[k])
        else:
            k = self.kv
            np.add.at(r, self.FOv, contrib_o[k])
            m = self.FNv < self.gB          # ghost targets are overwritten by fillBCs anyway
            np.subtract.at(r, self.FNv[m], contrib_n[k][m])
        return r

    def fill_bcs(self, r):
        r[self.fill_ghost] = r[self.fill_owner]

    # ---- volume operators ----------------------------------------------------------------------
    def _dpsi_vec(self, Jin_q, d, coef):
        """dot(Jin, e_d*coef) with exact zeros elsewhere -> Jin[:, :, d]*coef"""
        return Jin_q[:, :, d] * coef

    def div_volume(self, r, F):
        """weak form: r[m] -= dot(F[q], Jin_q . dpsi(q->m))   (field.h:3443-3464). F is (gA,3) or (gA,3,3)."""
        NPX, NPY, NPZ = self.shape
        nB, NP = self.nB, self.NP
        tensor = F.ndim == 3
        Fe = F[: self.gB].reshape((nB, NPX, NPY, NPZ) + F.shape[1:])
        Je = self.Jin.reshape(nB, NPX, NPY, NPZ, 3, 3)
        re = r[: self.gB].reshape((nB, NPX, NPY, NPZ) + r.shape[1:])
        D0, D1, D2 = self.D
        if not self.exact:
            if tensor:
                G = np.einsum("cijkab,cijkbd->cijkad", Fe, Je)      # G[a,d] = sum_b F[a,b] Jin[b,d]
                re -= np.einsum("cqjka,qi->cijka", G[..., 0], D0)
                re -= np.einsum("ciqka,qj->cijka", G[..., 1], D1)
                re -= np.einsum("cijqa,qk->cijka", G[..., 2], D2)
            else:
                G = np.einsum("cijkb,cijkbd->cijkd", Fe, Je)
                re -= np.einsum("cqjk,qi->cijk", G[..., 0], D0)
                re -= np.einsum("ciqk,qj->cijk", G[..., 1], D1)
                re -= np.einsum("cijq,qk->cijk", G[..., 2], D2)
            return r

        def dotF(Fq, v):
            if tensor:
                return (Fq[:, :, 0] * v[:, None, 0] + Fq[:, :, 1] * v[:, None, 1]) + Fq[:, :, 2] * v[:, None, 2]
            return Fq[:, 0] * v[:, 0] + (Fq[:, 1] * v[:, 1] + Fq[:, 2] * v[:, 2])

        for ii in range(NPX):
            for jj in range(NPY):
                for kk in range(NPZ):
                    Jq = Je[:, ii, jj, kk]
                    Fq = Fe[:, ii, jj, kk]
                    for i in range(NPX):
                        if i == ii:
                            d0, d1, d2 = D0[ii, ii], D1[jj, jj], D2[kk, kk]
                            v = (Jq[:, :, 0] * d0 + Jq[:, :, 1] * d1) + Jq[:, :, 2] * d2
                        else:
                            v = Jq[:, :, 0] * D0[ii, i]
                        re[:, i, jj, kk] -= dotF(Fq, v)
                    for j in range(NPY):
                        if j != jj:
                            re[:, ii, j, kk] -= dotF(Fq, Jq[:, :, 1] * D1[jj, j])
                    for k in range(NPZ):
                        if k != kk:
                            re[:, ii, jj, k] -= dotF(Fq, Jq[:, :, 2] * D2[kk, k])
        return r

    def grad_volume(self, r, P):
        """strong form: r[q] += mul(Jin_q . dpsi(q<-m), P[m])   (field.h:3336-3357). P is (gA,) or (gA,3)."""
        NPX, NPY, NPZ = self.shape
        nB = self.nB
        vec = P.ndim == 2
        Pe = P[: self.gB].reshape((nB, NPX, NPY, NPZ) + P.shape[1:])
        Je = self.Jin.reshape(nB, NPX, NPY, NPZ, 3, 3)
        re = r[: self.gB].reshape((nB, NPX, NPY, NPZ) + r.shape[1:])
        D0, D1, D2 = self.D
        if not self.exact:
            if vec:
                dP = np.stack([np.einsum("si,cijkb->csjkb", D0, Pe), np.einsum("sj,cijkb->ciskb", D1, Pe),
                               np.einsum("sk,cijkb->cijsb", D2, Pe)], axis=-2)          # [d,b]
                re += np.einsum("cijkad,cijkdb->cijkab", Je, dP)
            else:
                dP = np.stack([np.einsum("si,cijk->csjk", D0, Pe), np.einsum("sj,cijk->cisk", D1, Pe),
                               np.einsum("sk,cijk->cijs", D2, Pe)], axis=-1)
                re += np.einsum("cijkad,cijkd->cijka", Je, dP)
            return r

        def outer(v, Pm):
            if vec:
                return v[:, :, None] * Pm[:, None, :]
            return v * Pm[:, None]

        for ii in range(NPX):
            for jj in range(NPY):
                for kk in range(NPZ):
                    Jq = Je[:, ii, jj, kk]
                    acc = re[:, ii, jj, kk]
                    for i in range(NPX):
                        if i == ii:
                            d0, d1, d2 = D0[ii, ii], D1[jj, jj], D2[kk, kk]
                            v = (Jq[:, :, 0] * d0 + Jq[:, :, 1] * d1) + Jq[:, :, 2] * d2
                        else:
                            v = Jq[:, :, 0] * D0[ii, i]
                        acc += outer(v, Pe[:, i, jj, kk])
                    for j in range(NPY):
                        if j != jj:
                            acc += outer(Jq[:, :, 1] * D1[jj, j], Pe[:, ii, j, kk])
                    for k in range(NPZ):
                        if k != kk:
                            acc += outer(Jq[:, :, 2] * D2[kk, k], Pe[:, ii, jj, k])
        return r

    # ---- operators of the reference API -----------------------------------------------------------
    def uds(self, c, flux):
        """upwind value by the sign of the facet flux (field.h:2904-2917)"""
        fO, fN = self.scatter(c)
        sel = (flux >= 0).reshape((-1,) + (1,) * (c.ndim - 1))
        return np.where(sel, fO, fN)

    def divf(self, F, q, lam, scheme="RUSANOV", flux=None, blend=0.0):
        """divf<weak>(F,false,&flux,&q,&lambdaMax) (field.h:3417-3478): RUSANOV (what the euler app runs), or CDS / UDS / BLENDED of the
        convection app's examples (flux = flx(U) on the used slots)."""
        if scheme == "RUSANOV":
            fF = self.rusanov(F, q, lam)                           # on used slots
        elif scheme == "CDS":
            fF = self.cds(F)
        elif scheme == "BLENDED":
            fF = blend * self.cds(F) + (1.0 - blend) * self.uds(F, flux)
        elif scheme == "UDS":
            fF = self.uds(F, flux)
        else:
            raise NotImplementedError(scheme)
        fFO, fFN = self.gather(fF)                                 # div_flux<weak>: gather_non_conforming first (field.h:3091)
        dotN = vdot if F.ndim == 2 else tdotv
        r = np.zeros(self.gA) if F.ndim == 2 else np.zeros((self.gA, 3))
        flux_o = dotN(fFO, self.fNv)
        flux_n = flux_o if fFN is fFO else dotN(fFN, self.fNv)
        full_o = self.face_full(flux_o, flux_o.shape[1:])
        full_n = full_o if flux_n is flux_o else self.face_full(flux_n, flux_n.shape[1:])
        self._accumulate_faces(r, full_o, full_n)
        self.div_volume(r, F)
        self.fill_bcs(r)
        return r

    def gradf(self, P, field_name):
        """gradf<strong>(P, per-unit-volume=true) + fillBCs(r, P.fIndex) (field.h:3328-3362, 2745-2769)."""
        fF = self.cds(P)
        fFO, fFN = self.gather(fF)                                 # grad_flux<strong>: gather_non_conforming first (field.h:3056)
        dO = fFO - P[self.FOv]
        dN = fFN - P[self.FNv]
        if P.ndim == 1:
            co = self.fNv * dO[:, None]
            cn = self.fNv * dN[:, None]
            r = np.zeros((self.gA, 3))
        else:
            co = self.fNv[:, :, None] * dO[:, None, :]
            cn = self.fNv[:, :, None] * dN[:, None, :]
            r = np.zeros((self.gA, 3, 3))
        self._accumulate_faces(r, self.face_full(co, co.shape[1:]), self.face_full(cn, cn.shape[1:]))
        self.grad_volume(r, P)
        r = r / self.g.cV.reshape((-1,) + (1,) * (r.ndim - 1))
        self.fill_bcs(r)
        for bc in self.bcs[field_name]:
            gh, _, _ = self._bc_nodes(bc.faces)
            if bc.kind == "NEUMANN":
                r[gh] = self._punned_neumann(bc, r.shape[1:])
            elif bc.kind == "SYMMETRY":
                r[gh] = 0.0
        return r

    @staticmethod
    def _punned_neumann(bc: BCSpec, shape):
        """fillBCs reads BCondition<T>::value through a pointer of the GRADIENT type (field.h:2750,2761): the bytes
        that follow `value` in the object are shape, tvalue, tshape, zMin (field.h:146-150)."""
        val = np.atleast_1d(bc.value).astype(float)
        tv = np.zeros_like(val) if bc.tvalue is None else np.atleast_1d(bc.tvalue).astype(float)
        raw = np.concatenate([val, [bc.shape], tv, [bc.tshape, bc.zMin, 0.0, 0.0, 0.0]])
        n = int(np.prod(shape))
        raw = raw[:n]
        if len(shape) == 2:
            # Tensor AoS order XX,YY,ZZ,XY,YZ,XZ,YX,ZY,ZX -> row-major 3x3
            from .dg import T9
            out = np.zeros((3, 3))
            for idx, (a, b) in enumerate(T9):
                out[a, b] = raw[idx]
            return out
        return raw.reshape(shape)

    # ---- boundary conditions ------------------------------------------------------------------
    def _bc_nodes(self, faces):
        NPF = self.NPF
        ks = (np.asarray(faces, dtype=np.int64)[:, None] * NPF + np.arange(NPF)[None, :]).reshape(-1)
        m = self.g.FN[ks] < self.gA
        return self.g.FN[ks][m], self.g.FO[ks][m], ks[m]

    def apply_bcs(self, name, F):
        """applyExplicitBCs (field.h:2586-2727), single rank."""
        g = self.g
        for bc in self.bcs[name]:
            if bc.kind == "GHOST" or len(bc.faces) == 0:
                continue
            gh, ow, ks = self._bc_nodes(bc.faces)
            if bc.kind == "NEUMANN":
                dv = vmag(g.cC[gh] - g.cC[ow])
                if F.ndim == 1:
                    F[gh] = F[ow] + float(np.atleast_1d(bc.value)[0]) * dv
                else:
                    F[gh] = F[ow] + np.asarray(bc.value)[None, :] * dv[:, None]
            elif bc.kind == "SYMMETRY":
                F[gh] = F[ow] if F.ndim == 1 else sym_vec(F[ow], g.fN[ks])
            elif bc.kind == "CYCLIC":
                NPF = self.NPF
                k1 = (np.asarray(bc.neighbor_faces, dtype=np.int64)[:, None] * NPF + np.arange(NPF)[None, :]).reshape(-1)
                own_ks = (np.asarray(bc.faces, dtype=np.int64)[:, None] * NPF + np.arange(NPF)[None, :]).reshape(-1)
                k1 = k1[g.FN[own_ks] < self.gA]
                F[gh] = F[g.FO[k1]]
            elif bc.kind == "DIRICHLET":
                F[gh] = float(np.atleast_1d(bc.value)[0]) if F.ndim == 1 else np.asarray(bc.value)[None, :]
            elif bc.kind == "CALC_DIRICHLET":
                if bc.fixed is None:
                    bc.fixed = F[ow].copy()
                F[gh] = bc.fixed
            else:
                raise NotImplementedError(f"BC kind {bc.kind}")

    # ---- set-up (euler.cpp:58-176) --------------------------------------------------------------
    def setup(self, rho, U, T, p, bcs_in: dict):
        """rho,U,T,p: arrays over gALL as read from the field files (before BCs). bcs_in: name -> [BCSpec]."""
        P = self.p
        g = self.g
        import copy
        for k in ("rho", "U", "T", "p"):
            self.bcs[k] = bcs_in.get(k, [])
        self.rho, self.U, self.T, self.pp = rho.copy(), U.copy(), T.copy(), p.copy()
        # MeshField::read_ applies the BCs right after reading (field.h:1562-1565)
        self.apply_bcs("p", self.pp)
        self.apply_bcs("U", self.U)
        self.apply_bcs("T", self.T)
        self.apply_bcs("rho", self.rho)
        R = P.cp - P.cv
        gamma = P.cp / P.cv
        self.R, self.gamma, self.iPr = R, gamma, 1 / P.Pr
        if P.problem_init == "ISENTROPIC_VORTEX":
            beta = 5.0
            PI = 3.14159265358979323846264
            r = vmag(g.cC)
            self.T = (-((gamma - 1) * beta * beta) / (8 * gamma * PI * PI)) * libm.exp_(1 - r * r)
            e = libm.exp_((1 - r * r) / 2.0)
            nb = self.gB
            self.U[:nb, 0] += (beta / (2 * PI)) * e[:nb] * -g.cC[:nb, 1]
            self.U[:nb, 1] += (beta / (2 * PI)) * e[:nb] * g.cC[:nb, 0]
            self.pp = libm.pow_(self.T + P.T0, gamma / (gamma - 1)) - P.P0
            self.rho = (P.P0 / (R * (self.T + P.T0))) * libm.pow_((self.pp + P.P0) / P.P0, 1 / gamma) - (P.P0 / (R * P.T0))
        if P.buoyancy:
            grav = np.array(P.gravity, dtype=float)
            if P.is_spherical:      # gravity towards the centre of the sphere, euler.cpp:109-111
                r = vmag(g.cC)
                mg = float(vmag(grav[None, :])[0])
                self.gvec = -(g.cC / r[:, None]) * mg
                self.gh = -(r - P.sphere_radius) * mg
            else:
                self.gvec = np.tile(grav, (self.gA, 1))
                self.gh = vdot(self.gvec, g.cC)
            # fixedBCs<Vector>(U,g): every BC of U becomes CALC_DIRICHLET for g (field.h:2779-2795)
            self.bcs["g"] = [BCSpec("CALC_DIRICHLET", bc.faces, np.zeros(3)) for bc in self.bcs["U"]]
            self.apply_bcs("g", self.gvec)
            self.p_ref = P.P0 * libm.pow_(1.0 + self.gh / (P.cp * P.T0), P.cp / R)
            self.rho_ref = (P.P0 / (R * P.T0)) * libm.pow_(self.p_ref / P.P0, 1 / gamma)
        else:
            self.gvec = np.zeros((self.gA, 3))
            self.gh = np.zeros(self.gA)
            self.p_ref = np.full(self.gA, P.P0)
            self.rho_ref = np.full(self.gA, P.P0 / (R * P.T0))

        def scale_bcs(src):
            out = []
            for bc in self.bcs[src]:
                nb_ = copy.deepcopy(bc)
                if nb_.kind not in ("NEUMANN", "ROBIN", "SYMMETRY", "CYCLIC"):
                    nb_.kind = "CALC_DIRICHLET"
                    nb_.fixed = None
                out.append(nb_)
            return out

        # ait.start() branch (euler.cpp:133-146)
        self.pp = self.pp + self.p_ref
        self.bcs["p_ref"] = scale_bcs("p")
        self.apply_bcs("p_ref", self.p_ref)
        self.apply_bcs("p", self.pp)
        self.rho = (P.P0 / (R * (self.T + P.T0))) * libm.pow_(self.pp / P.P0, 1 / gamma)
        self.apply_bcs("rho", self.rho)
        self.bcs["rho_ref"] = scale_bcs("rho")
        self.apply_bcs("rho_ref", self.rho_ref)
        self.pp = self.pp - self.p_ref
        self.mass0, self.energy0, self.volume0 = self.diagnostics_sums(self.T + P.T0)

    def diagnostics_sums(self, theta):
        P = self.p
        cV = self.g.cV
        nb = self.gB
        sf = self.rho * cV
        mass = float(np.sum(sf[:nb]))
        e = self.gh + 0.5 * vdot(self.U, self.U) + libm.pow_((self.pp + self.p_ref) / P.P0, self.R / P.cp) * theta * P.cv
        sf = self.rho * cV * e
        return mass, float(np.sum(sf[:nb])), float(np.sum(cV[:nb]))

    # ---- one time step (euler.cpp:179-287) ---------------------------------------------------------
    def step(self):
        P = self.p
        g = self.g
        cV = g.cV
        R, gamma = self.R, self.gamma
        rho, U = self.rho, self.U
        T = self.T + P.T0                                            # :181
        Fc = rho[:, None] * U                                        # :184
        lam = (self.cds(vmag(U)) + self.cds(libm.sqrt_(gamma * R * T))) / 2      # :186
        mu = rho * P.viscosity if P.diffusion else np.zeros_like(rho)       # :189-190
        rhof = rho.copy()                                            # :193
        ap0 = (-1.0 / P.dt) * cV
        bdf = P.time_scheme.startswith("BDF")

        def temporal(r, prev):
            return (r + prev * ap0.reshape((-1,) + (1,) * (r.ndim - 1))) if bdf else \
                   (prev * ap0.reshape((-1,) + (1,) * (r.ndim - 1)) + r)

        # rho-equation :195-208
        fq = U * rho[:, None]
        r = self.divf(fq, rho, lam)
        Su = temporal(r, rho)
        rho = Su / ap0
        self.apply_bcs("rho", rho)
        # pressure :211-213
        pp = P.P0 * libm.pow_((rho * T * R) / P.P0, gamma)
        self.apply_bcs("p", pp)
        pp = pp - self.p_ref
        # U-equation :216-240
        Sc = np.zeros_like(U)
        if P.buoyancy:
            Sc = Sc + (rho - self.rho_ref)[:, None] * self.gvec
        G = self.gradf(U, "U")
        eye = np.eye(3)
        fqT = (Fc[:, :, None] * U[:, None, :] + eye[None] * pp[:, None, None]) - mu[:, None, None] * G
        q = rho[:, None] * U
        r = self.divf(fqT, q, lam)
        Su = r - Sc * cV[:, None]
        Su = temporal(Su, U * rhof[:, None])
        ap = ap0 * rho
        Unew = Su / ap[:, None]
        self.apply_bcs("U", Unew)
        # T-equation :243-258
        imu = mu * self.iPr
        fq = Fc * T[:, None] - imu[:, None] * self.gradf(T, "T")
        q = rho * T
        r = self.divf(fq, q, lam)
        Su = temporal(r, T * rhof)
        Tnew = Su / ap
        self.apply_bcs("T", Tnew)
        self.rho, self.U, self.pp = rho, Unew, pp
        self.theta = Tnew                                           # T + T0 (what the conserved rho*theta uses)
        self.T = Tnew - P.T0                                        # :286
        self.step_count += 1

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self
