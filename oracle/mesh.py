"""CPU restatement of the reference's mesh topology + element geometry -- TEST INFRASTRUCTURE ONLY.

Follows (all in /root/reference):
  addBoundaryCells     src/mesh/mesh.cpp:55-109
  getHexCorners        src/mesh/mesh.cpp:113-157
  fixHexCells          src/mesh/mesh.cpp:161-446   (non USE_HEX_REFINEMENT branch)
  calcGeometry         src/mesh/mesh.cpp:450-577   (with the cubed-sphere corrections :520-570)
  ExtrudeMesh          src/mesh/mesh.cpp:723-750   (is_spherical)
  cart_to_sphere/geodesic_distance/spherical_triangle_area   src/tensor/tensor.h:598-635
  removeBoundary       src/mesh/mesh.cpp:581-669
  pointInLine/calcUnitNormal/coplanarFaces/mergeFacets/mergeFacetsGroup  mesh.cpp:791-1020
  LoadMesh             src/field/field.cpp:95-167

Pure-Python loops: meant for the parity meshes (<= a few 10^4 cells), not for throughput.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

from .refio import Grid

MAX_INT = 1 << 31
EQ_EPS = 1e-7  # Constants::EqualEpsilon, tensor.h:462


def equal(p: float, q: float, tol: float = EQ_EPS) -> bool:
    """tensor.h:469-474"""
    d = abs(p - q)
    return d <= tol or d <= tol * abs(p) or d <= tol * abs(q)


def _cross(p, q):
    return np.array([p[1] * q[2] - p[2] * q[1], p[2] * q[0] - p[0] * q[2], p[0] * q[1] - p[1] * q[0]])


def _unit(v):
    return v / math.sqrt(float(v @ v))


def _dot(a, b) -> float:
    """Unroll<3>::dot nests to the right (tensor.h:124-127)"""
    return float(a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]))


def _mag(a) -> float:
    return math.sqrt(_dot(a, a))


def cart_to_sphere(c):
    """(radius, latitude, longitude), tensor.h:598-605; math.* is the platform libm the reference binary links"""
    return np.array([_mag(c), math.atan2(c[2], math.sqrt(float(c[0] * c[0] + c[1] * c[1]))), math.atan2(c[1], c[0])])


def geodesic_distance(s1, s2) -> float:
    """tensor.h:608-612"""
    d = (s1[0] + s2[0]) / 2
    d *= math.acos(math.sin(s1[1]) * math.sin(s2[1]) + math.cos(s1[1]) * math.cos(s2[1]) * math.cos(s1[2] - s2[2]))
    return float(d)


def spherical_triangle_area(radius: float, v0, v1, v2) -> float:
    """tensor.h:624-635"""
    a, b, c = v0 / _mag(v0), v1 / _mag(v1), v2 / _mag(v2)
    t = abs(_dot(a, _cross(b, c)))
    t /= (1 + _dot(a, b) + _dot(b, c) + _dot(a, c))
    return 2 * math.atan(t) * radius * radius


class MeshTopo:
    """State of Mesh::MeshObject after LoadMesh (before the DG node expansion)."""

    # Mesh::is_spherical / sphere_radius / sphere_height (mesh.cpp:31-33); set before load()
    spherical = False
    sphere_radius = 6371220.0
    sphere_height = 10000.0

    def __init__(self, grid: Grid):
        self.V = np.array(grid.vertices, dtype=np.float64)
        self.facets = [list(f) for f in grid.facets]
        self.cells = [list(c) for c in grid.cells]
        # std::map iterates in key order
        self.boundaries = {k: list(v) for k, v in sorted(grid.boundaries.items())}

    # ---- small geometric predicates -------------------------------------------------------
    def point_in_line(self, v, v1, v2) -> bool:
        p = v - v1
        q = v - v2
        if not equal(p[1] * q[2] - p[2] * q[1], 0.0):
            return False
        if not equal(p[2] * q[0] - p[0] * q[2], 0.0):
            return False
        if not equal(p[0] * q[1] - p[1] * q[0], 0.0):
            return False
        e = float((v - v2) @ (v1 - v2))
        if e > 0:
            e1 = float((v1 - v2) @ (v1 - v2))
            if e < e1:
                return True
        return False

    def unit_normal(self, f):
        V = self.V
        v1, v2 = V[f[0]], V[f[1]]
        for j in range(1, len(f)):
            v3 = V[f[len(f) - j]]
            if not self.point_in_line(v2, v1, v3):
                return _unit(_cross(v2 - v1, v3 - v1))
        return None

    def coplanar(self, f1, f2) -> bool:
        n1 = self.unit_normal(f1)
        n2 = self.unit_normal(f2)
        c = _cross(n1, n2)
        if equal(float(c @ c), 0.0):
            v = self.V[f2[1]] - self.V[f1[0]]
            if equal(float(n1 @ v), 0.0):
                return True
        return False

    def merge_facets(self, f1_, f2_):
        """Union of two edge-sharing coplanar polygons (mesh.cpp:890-980); returns None when they share no edge."""
        f1, f2 = list(f1_), list(f2_)
        if float(self.unit_normal(f1_) @ self.unit_normal(f2_)) < 0:
            f2[1:] = f2[1:][::-1]
        v1 = next((i for i, x in enumerate(f1) if x not in f2), -1)
        v2 = next((i for i, x in enumerate(f2) if x not in f1), -1)
        contained = False
        f1 = f1[v1:] + f1[:v1]
        if v2 != -1:
            f2 = f2[v2:] + f2[:v2]
        else:
            contained = True
        a = [0, 0]
        b = [0, 0]
        count = 0
        for i, x in enumerate(f1):
            for j, y in enumerate(f2):
                if x == y:
                    if count == 0:
                        a[0], b[0] = i, j
                    else:
                        a[1], b[1] = i, j
                    count += 1
        if count < 2:
            return None
        f = f1[:a[0] + 1]
        if contained:
            f += f2[b[0] + 1:b[1]]
        else:
            f += f2[b[0] + 1:] + f2[:b[1]]
        f += f1[a[1]:]
        while self.point_in_line(self.V[f[0]], self.V[f[-1]], self.V[f[1]]):
            f = f[1:] + f[:1]
        return f

    def merge_group(self, faces):
        fn = list(self.facets[faces[0]])
        rest = list(faces[1:])
        while rest:
            merged = []
            repeat = False
            for j, fi in enumerate(rest):
                fm = self.merge_facets(fn, self.facets[fi])
                if fm is not None:
                    fn = fm
                    merged.append(j)
                else:
                    repeat = True
            rest = [x for j, x in enumerate(rest) if j not in merged]
            if not repeat:
                break
            if not merged:
                raise RuntimeError("mergeFacetsGroup cannot make progress")
        return fn

    # ---- addBoundaryCells -------------------------------------------------------------------
    def add_boundary_cells(self):
        nf = len(self.facets)
        self.nBCS = len(self.cells)
        FOC = [MAX_INT] * nf
        FNC = [MAX_INT] * nf
        for i, c in enumerate(self.cells):
            for fi in c:
                if FOC[fi] == MAX_INT:
                    FOC[fi] = i
                else:
                    FNC[fi] = i
        in_b = [0] * nf
        for name, fs in self.boundaries.items():
            if name == "delete":
                continue
            for f in fs:
                in_b[f] = 1
        self.boundaries["delete"] = [i for i in range(nf) if FNC[i] == MAX_INT and not in_b[i]]
        self.boundaries = dict(sorted(self.boundaries.items()))
        for name, fs in self.boundaries.items():
            for fi in fs:
                if FNC[fi] == MAX_INT:
                    self.cells.append([fi])
                    FNC[fi] = len(self.cells) - 1
        self.FOC, self.FNC = FOC, FNC

    # ---- getHexCorners ----------------------------------------------------------------------
    def hex_corners(self, f1, f2):
        V = self.V
        fm = [[], []]
        tol = math.pi / 16
        for w, fk in enumerate((f1, f2)):
            i0 = len(fk) - 1
            for i in range(len(fk)):
                if len(fm[w]) >= 4:
                    break
                i1 = 0 if i == len(fk) - 1 else i + 1
                v0 = _unit(V[fk[i]] - V[fk[i0]])
                v1 = _unit(V[fk[i1]] - V[fk[i]])
                dt = max(-1.0, min(1.0, float(v0[0] * v1[0] + (v0[1] * v1[1] + v0[2] * v1[2]))))
                ang = math.acos(dt)
                if not (ang < tol or ang >= math.pi - tol):
                    fm[w].append(fk[i])
                    i0 = i
        vp = list(fm[0][:4])
        mind, best = 1e20, None
        for order in itertools.permutations(range(4)):     # lexicographic == std::next_permutation from sorted
            dist = 0.0
            for i in range(4):
                d = V[fm[1][order[i]]] - V[fm[0][i]]
                dist += math.sqrt(float(d[0] * d[0] + (d[1] * d[1] + d[2] * d[2])))
            if dist < mind:
                mind, best = dist, order
        vp += [fm[1][best[i]] for i in range(4)]
        return vp

    # ---- fixHexCells ------------------------------------------------------------------------
    def fix_hex_cells(self):
        V = self.V
        self.FMC = [0] * len(self.facets)
        self.faceID = []
        for cidx, c in enumerate(self.cells):
            if len(c) == 1:
                self.faceID.append([0])
                continue
            # group coplanar faces, always seeding with the first remaining face
            rem = list(c)
            cng, fng = [], []
            for _ in range(6):
                grp = [rem[0]]
                keep = []
                for fj in rem[1:]:
                    if self.coplanar(self.facets[rem[0]], self.facets[fj]):
                        grp.append(fj)
                    else:
                        keep.append(fj)
                rem = keep
                cng.append(grp)
                fng.append(self.merge_group(grp))
            gid = [-1] * 6
            f0 = fng[0]
            for j in range(6):
                if gid[j] >= 0:
                    continue
                fj = fng[j]
                ident = j
                if j >= 1:
                    local = -1
                    if sum(1 for x in fj if x == f0[0] or x == f0[1]) >= 2:
                        local = 2
                    elif sum(1 for x in fj if x == f0[0] or x == f0[-1]) >= 2:
                        local = 4
                    if local < 0:
                        continue
                    ident = local
                gid[j] = ident
                for k in range(6):
                    if gid[k] >= 0:
                        continue
                    if not any(x in fng[k] for x in fj):
                        gid[k] = gid[j] ^ 1
                        break
            N = _cross(V[f0[1]] - V[f0[0]], V[f0[-1]] - V[f0[0]])
            e = V[fng[1][0]] - V[f0[0]]
            if float(N @ e) < 0:
                gid = [1 if g == 0 else 0 if g == 1 else g for g in gid]
            i0 = gid.index(0)
            i1 = gid.index(1)
            fa, fb = cng[i0][0], cng[i1][0]
            i0n = self.FNC[fa] if self.FNC[fa] != cidx else self.FOC[fa]
            i1n = self.FNC[fb] if self.FNC[fb] != cidx else self.FOC[fb]
            flip = (i0n > i1n) and (cidx >= i0n or cidx >= i1n)
            if not flip:
                vp = self.hex_corners(fng[i0], fng[i1])
            else:
                vp = self.hex_corners(fng[i1], fng[i0])
                vp = vp[4:] + vp[:4]
            rots = [vp[0], vp[4], vp[0], vp[3], vp[0], vp[1]]
            rote = [vp[1], vp[5], vp[1], vp[2], vp[3], vp[2]]
            for i in range(6):
                fn = fng[i]
                g = gid[i]
                rs, re = rots[g], rote[g]
                p = fn.index(rs)
                fn[:] = fn[p:] + fn[:p]
                d = float(_unit(V[fn[1]] - V[fn[0]]) @ _unit(V[re] - V[rs]))
                if d < 0.99:
                    fn[1:] = fn[1:][::-1]
                for j, fid in enumerate(cng[i]):
                    f = self.facets[fid]
                    it1 = next(k for k, x in enumerate(fn) if x in f)
                    p2 = f.index(fn[it1])
                    f[:] = f[p2:] + f[:p2]
                    if j >= 2:
                        # std::rotate(f.rbegin(), f.rbegin() + (j-1), f.rend()) == rotate right by (j-1)
                        s = (j - 1) % len(f)
                        if s:
                            f[:] = f[-s:] + f[:-s]
                    for k in range(1, len(f)):
                        if f[k] not in fn:
                            continue
                        it3 = fn.index(f[k])
                        if it1 > it3:
                            f[1:] = f[1:][::-1]
                            break
                        it1 = it3
            newc, ids = [], []
            for want in range(6):
                for i in range(6):
                    if gid[i] == want:
                        for fid in cng[i]:
                            newc.append(fid)
                            ids.append(want)
            c[:] = newc
            self.faceID.append(ids)
            if len(c) > 6:
                for j, f in enumerate(c):
                    if ids.count(ids[j]) > 1:
                        self.FMC[f] = 2 if self.FOC[f] == cidx else 1

    # ---- calcGeometry -----------------------------------------------------------------------
    def calc_geometry(self):
        V = self.V
        nf, nc = len(self.facets), len(self.cells)
        FC = np.zeros((nf, 3))
        FNv = np.zeros((nf, 3))
        CC = np.zeros((nc, 3))
        CV = np.zeros(nc)
        for i, f in enumerate(self.facets):
            C = np.zeros(3)
            for v in f:
                C = C + V[v]
            FC[i] = C / float(len(f))
        for i, c in enumerate(self.cells):
            C = np.zeros(3)
            for fi in c:
                C = C + FC[fi]
            CC[i] = C / float(len(c))
        for i, f in enumerate(self.facets):
            N = np.zeros(3)
            C = np.zeros(3)
            Ntot = 0.0
            v1 = FC[i].copy()
            n = len(f)
            for j in range(n):
                v2 = V[f[j]]
                v3 = V[f[0 if j + 1 == n else j + 1]]
                Ni = _cross(v2 - v1, v3 - v1)
                magN = math.sqrt(float(Ni[0] * Ni[0] + (Ni[1] * Ni[1] + Ni[2] * Ni[2])))
                C = C + magN * ((v1 + v2 + v3) / 3)
                Ntot += magN
                N = N + Ni
            FC[i] = C / Ntot
            v = FC[i] - CC[self.FOC[i]]
            if float(v[0] * N[0] + (v[1] * N[1] + v[2] * N[2])) < 0:
                N = -N
            FNv[i] = N / 2.0
        for i in range(self.nBCS):
            c = self.cells[i]
            Vt = 0.0
            C = np.zeros(3)
            for fi in c:
                v = CC[i] - FC[fi]
                Vi = abs(float(v[0] * FNv[fi][0] + (v[1] * FNv[fi][1] + v[2] * FNv[fi][2])))
                C = C + Vi * (3 * FC[fi] + CC[i]) / 4
                Vt += Vi
            CC[i] = C / Vt
            CV[i] = Vt / 3.0
        if self.spherical:
            self._sphere_geometry(FC, FNv, CC, CV)
        for i in range(self.nBCS, nc):
            fi = self.cells[i][0]
            CV[i] = CV[self.FOC[fi]]
            CC[i] = FC[fi]
        self.FC, self.FNv, self.CC, self.CV = FC, FNv, CC, CV

    # ---- cubed-sphere shells ------------------------------------------------------------------
    def extrude(self):
        """ExtrudeMesh (mesh.cpp:723-750): the grid file holds a shell between two concentric cubes; the cube a vertex lies on (its largest
        |coordinate|) picks the radius it is projected to -- the SMALLER cube goes to the OUTER radius, as in the reference."""
        V = self.V
        h = [max(max(abs(float(v[0])), abs(float(v[1]))), abs(float(v[2]))) for v in V]
        minh, maxh = min(h + [1e30]), max(h + [0.0])
        ri, ro = self.sphere_radius, self.sphere_radius + self.sphere_height
        for i in range(len(V)):
            f = (h[i] - minh) / (maxh - minh)
            V[i] = (V[i] / _mag(V[i])) * (f * ri + (1 - f) * ro)

    def _sphere_geometry(self, FC, FNv, CC, CV):
        """mesh.cpp:520-570: centres pushed to the shell radii, vertical face areas from the geodesic length of their outline, radial face
        areas from spherical triangles, volumes = thickness x mean radial area.  Sides 0 and 1 of a cell are its two radial faces."""
        V = self.V
        for i in range(self.nBCS):
            c = self.cells[i]
            rb = _mag(V[self.facets[c[0]][0]])
            rt = _mag(V[self.facets[c[1]][0]])
            CC[i] = ((rb + rt) / (2 * _mag(CC[i]))) * CC[i]
            FC[c[0]] = (rb / _mag(FC[c[0]])) * FC[c[0]]
            FC[c[1]] = (rt / _mag(FC[c[1]])) * FC[c[1]]
            for j in range(2, 6):
                FC[c[j]] = ((rb + rt) / (2 * _mag(FC[c[j]]))) * FC[c[j]]
                d = 0.0
                f = self.facets[c[j]]
                for k in range(len(f)):
                    v0 = V[f[k]]
                    v1 = V[f[0 if k == len(f) - 1 else k + 1]]
                    r0, r1 = v0 / _mag(v0), v1 / _mag(v1)
                    if equal(r0[0], r1[0]) and equal(r0[1], r1[1]) and equal(r0[2], r1[2]):
                        continue
                    d += geodesic_distance(cart_to_sphere(v0), cart_to_sphere(v1))
                d /= 2
                area = abs(rt - rb) * d
                FNv[c[j]] = area * (FNv[c[j]] / _mag(FNv[c[j]]))
        for i in range(self.nBCS):
            c = self.cells[i]
            area = 0.0
            for k in range(2):
                f = self.facets[c[k]]
                radius = _mag(V[f[0]])
                a = 0.0
                for j in range(len(f)):
                    a += spherical_triangle_area(radius, V[f[j]], V[f[0 if j == len(f) - 1 else j + 1]], FC[c[k]])
                FNv[c[k]] = a * (FNv[c[k]] / _mag(FNv[c[k]]))
                area += a
            area /= 2
            rb = _mag(V[self.facets[c[0]][0]])
            rt = _mag(V[self.facets[c[1]][0]])
            CV[i] = abs(rt - rb) * area

    # ---- removeBoundary ("delete" patch of 2-D meshes) ----------------------------------------
    def remove_boundary(self, fs):
        fs_set = set(fs)
        for f in fs:
            for cell_i in (self.FOC[f], self.FNC[f]):
                c = self.cells[cell_i]
                ids = self.faceID[cell_i]
                if f in c:
                    j = c.index(f)
                    del c[j]
                    del ids[j]
        nf = len(self.facets)
        idf = [MAX_INT] * nf
        cnt = 0
        for i in range(nf):
            if i not in fs_set:
                idf[i] = cnt
                cnt += 1
        keepf = [i for i in range(nf) if i not in fs_set]
        nc = len(self.cells)
        idc = [MAX_INT] * nc
        cnt = 0
        for i in range(nc):
            if len(self.cells[i]) != 0:
                idc[i] = cnt
                cnt += 1
        keepc = [i for i in range(nc) if len(self.cells[i]) != 0]
        self.facets = [self.facets[i] for i in keepf]
        self.FOC = [idc[self.FOC[i]] for i in keepf]
        self.FNC = [idc[self.FNC[i]] for i in keepf]
        self.FMC = [self.FMC[i] for i in keepf]
        self.FC = self.FC[keepf]
        self.FNv = self.FNv[keepf]
        self.cells = [[idf[f] for f in self.cells[i]] for i in keepc]
        self.faceID = [self.faceID[i] for i in keepc]
        self.CC = self.CC[keepc]
        self.CV = self.CV[keepc]
        for name in self.boundaries:
            self.boundaries[name] = [idf[f] for f in self.boundaries[name]]

    # ---- LoadMesh ---------------------------------------------------------------------------
    def load(self):
        self.add_boundary_cells()
        self.fix_hex_cells()
        if self.spherical:
            self.extrude()
        self.calc_geometry()
        fs = self.boundaries.pop("delete")
        self.remove_boundary(fs)
        self.boundaries = {k: v for k, v in self.boundaries.items() if len(v) > 0 and "interior" not in k}
        return self
