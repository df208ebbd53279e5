// mesh.cpp -- mesh topology and element geometry for conforming hexahedral grids.
//
// Produces the state Mesh::LoadMesh leaves behind (src/field/field.cpp:95-167):
//   boundary ("ghost") cells appended per patch in name order      mesh.cpp:55-109
//   local face ids 0/1 = zeta-/+, 2/3 = eta-/+, 4/5 = xi-/+ and a canonical vertex order per facet
//                                                                    mesh.cpp:161-446 (+ getHexCorners :113-157)
//   facet centroids/area vectors, cell centroids/volumes            mesh.cpp:450-577
//   removal of the "delete" patch of 2-D meshes                     mesh.cpp:581-669
// The element's local axes -- and therefore the node order of every field file -- depend on the facet vertex
// order the reference leaves after visiting the cells in index order, so the cells are visited in the same
// order here and each visit re-orients the six facets of the cell exactly as the reference does.
// Non-conforming (AMR) cells (more than six faces, facets with hanging vertices) are rejected.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "nsem_host.h"

namespace nsemh {

namespace {
inline Vec3 sub(const Vec3& a, const Vec3& b) { return Vec3{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline Vec3 add(const Vec3& a, const Vec3& b) { return Vec3{a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
inline Vec3 mul(const Vec3& a, double s) { return Vec3{a[0] * s, a[1] * s, a[2] * s}; }
inline Vec3 divs(const Vec3& a, double s) { return Vec3{a[0] / s, a[1] / s, a[2] / s}; }
// Unroll<3>::dot nests to the right (tensor.h:124-127)
inline double dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline Vec3 cross(const Vec3& p, const Vec3& q) {
    return Vec3{p[1] * q[2] - p[2] * q[1], p[2] * q[0] - p[0] * q[2], p[0] * q[1] - p[1] * q[0]};
}
inline double mag(const Vec3& a) { return std::sqrt(dot(a, a)); }
inline Vec3 unit(const Vec3& a) { return divs(a, mag(a)); }
}  // namespace

// ---------------------------------------------------------------------------------------------------------
void MeshTopo::add_boundary_cells() {
    const u32 nf = nFacets();
    nBCS = nCells();
    FOC.assign(nf, MAX_INT);
    FNC.assign(nf, MAX_INT);
    for (u32 i = 0; i < nBCS; i++)
        for (u32 q = cellStart[i]; q < cellStart[i + 1]; q++) {
            const u32 fi = cellFaces[q];
            if (FOC[fi] == MAX_INT) FOC[fi] = i;
            else FNC[fi] = i;
        }
    std::vector<char> inB(nf, 0);
    for (const auto& kv : boundaries) {
        if (kv.first == "delete") continue;
        for (u32 f : kv.second) inB[f] = 1;
    }
    auto& del = boundaries["delete"];
    del.clear();
    for (u32 i = 0; i < nf; i++)
        if (FNC[i] == MAX_INT && !inB[i]) del.push_back(i);
    for (const auto& kv : boundaries)
        for (u32 fi : kv.second)
            if (FNC[fi] == MAX_INT) {
                cellFaces.push_back(fi);
                cellStart.push_back((u32)cellFaces.size());
                FNC[fi] = nCells() - 1;
            }
}

// first four vertices of f1, then the vertices of f2 paired to them by minimum total distance (mesh.cpp:113-157)
void MeshTopo::hex_corners(const u32* f1, const u32* f2, u32 out[8]) const {
    for (int i = 0; i < 4; i++) out[i] = f1[i];
    int order[4] = {0, 1, 2, 3}, best[4] = {0, 1, 2, 3};
    double mind = 1e20;
    do {
        double dist = 0;
        for (int i = 0; i < 4; i++) dist += mag(sub(V[f2[order[i]]], V[f1[i]]));
        if (dist < mind) { mind = dist; std::memcpy(best, order, sizeof best); }
    } while (std::next_permutation(order, order + 4));
    for (int i = 0; i < 4; i++) out[4 + i] = f2[best[i]];
}

void MeshTopo::fix_hex_cells() {
    FMC.assign(nFacets(), 0);
    cellFaceID.assign(cellFaces.size(), 0);
    for (u32 ci = 0; ci < nBCS; ci++) {
        const u32 c0 = cellStart[ci];
        if (cellStart[ci + 1] - c0 != 6)
            throw Error("cell " + std::to_string(ci) + " does not have 6 faces: non-conforming (AMR) grids are not supported by this build");
        u32 fc[6];
        u32* fv[6];
        for (int q = 0; q < 6; q++) {
            fc[q] = cellFaces[c0 + q];
            if (facetStart[fc[q] + 1] - facetStart[fc[q]] != 4)
                throw Error("facet " + std::to_string(fc[q]) + " is not a quadrilateral: hanging nodes are not supported by this build");
            fv[q] = &facetVerts[facetStart[fc[q]]];
        }
        auto has = [&](int q, u32 v) { return fv[q][0] == v || fv[q][1] == v || fv[q][2] == v || fv[q][3] == v; };
        auto shares = [&](int a, int b) { for (int k = 0; k < 4; k++) if (has(b, fv[a][k])) return true; return false; };
        // ids: face 0 is the first face of the cell; 2 shares its first edge, 4 its last edge; opposites get id^1
        int gid[6] = {-1, -1, -1, -1, -1, -1};
        const u32* f0 = fv[0];
        for (int j = 0; j < 6; j++) {
            if (gid[j] >= 0) continue;
            int id = j;
            if (j >= 1) {
                if (has(j, f0[0]) && has(j, f0[1])) id = 2;
                else if (has(j, f0[0]) && has(j, f0[3])) id = 4;
                else continue;
            }
            gid[j] = id;
            for (int k = 0; k < 6; k++) {
                if (gid[k] >= 0) continue;
                if (!shares(j, k)) { gid[k] = id ^ 1; break; }
            }
        }
        for (int q = 0; q < 6; q++)
            if (gid[q] < 0) throw Error("cell " + std::to_string(ci) + " is not a hexahedron with 3 pairs of opposite faces");
        const Vec3 N = cross(sub(V[f0[1]], V[f0[0]]), sub(V[f0[3]], V[f0[0]]));
        const Vec3 e = sub(V[fv[1][0]], V[f0[0]]);
        if (dot(N, e) < 0)
            for (int q = 0; q < 6; q++) gid[q] = (gid[q] == 0) ? 1 : (gid[q] == 1 ? 0 : gid[q]);
        int i0 = 0, i1 = 0;
        for (int q = 0; q < 6; q++) { if (gid[q] == 0) i0 = q; else if (gid[q] == 1) i1 = q; }
        const u32 i0n = (FNC[fc[i0]] != ci) ? FNC[fc[i0]] : FOC[fc[i0]];
        const u32 i1n = (FNC[fc[i1]] != ci) ? FNC[fc[i1]] : FOC[fc[i1]];
        const bool flip = (i0n > i1n) && (ci >= i0n || ci >= i1n);
        u32 vp[8];
        if (!flip) hex_corners(fv[i0], fv[i1], vp);
        else {
            u32 t[8];
            hex_corners(fv[i1], fv[i0], t);
            for (int q = 0; q < 4; q++) { vp[q] = t[q + 4]; vp[q + 4] = t[q]; }
        }
        const u32 rots[6] = {vp[0], vp[4], vp[0], vp[3], vp[0], vp[1]};
        const u32 rote[6] = {vp[1], vp[5], vp[1], vp[2], vp[3], vp[2]};
        for (int q = 0; q < 6; q++) {
            u32* f = fv[q];
            const u32 rs = rots[gid[q]], re = rote[gid[q]];
            int p = -1;
            for (int k = 0; k < 4; k++) if (f[k] == rs) p = k;
            if (p < 0) throw Error("cell " + std::to_string(ci) + ": facet does not contain its reference corner");
            std::rotate(f, f + p, f + 4);
            const double d = dot(unit(sub(V[f[1]], V[f[0]])), unit(sub(V[re], V[rs])));
            if (d < 0.99) std::reverse(f + 1, f + 4);
        }
        // rewrite the face list in id order
        for (int want = 0, w = 0; want < 6; want++)
            for (int q = 0; q < 6; q++)
                if (gid[q] == want) { cellFaces[c0 + w] = fc[q]; cellFaceID[c0 + w] = (u32)want; w++; }
    }
}

void MeshTopo::calc_geometry() {
    const u32 nf = nFacets(), nc = nCells();
    FC.assign(nf, Vec3{0, 0, 0});
    FN.assign(nf, Vec3{0, 0, 0});
    CC.assign(nc, Vec3{0, 0, 0});
    CV.assign(nc, 0.0);
    for (u32 i = 0; i < nf; i++) {
        Vec3 C{0, 0, 0};
        const u32 n = facetStart[i + 1] - facetStart[i];
        for (u32 q = facetStart[i]; q < facetStart[i + 1]; q++) C = add(C, V[facetVerts[q]]);
        FC[i] = divs(C, (double)n);
    }
    for (u32 i = 0; i < nc; i++) {
        Vec3 C{0, 0, 0};
        for (u32 q = cellStart[i]; q < cellStart[i + 1]; q++) C = add(C, FC[cellFaces[q]]);
        CC[i] = divs(C, (double)(cellStart[i + 1] - cellStart[i]));
    }
    for (u32 i = 0; i < nf; i++) {
        Vec3 N{0, 0, 0}, C{0, 0, 0};
        double Ntot = 0;
        const Vec3 v1 = FC[i];
        const u32 s = facetStart[i], n = facetStart[i + 1] - s;
        for (u32 j = 0; j < n; j++) {
            const Vec3& v2 = V[facetVerts[s + j]];
            const Vec3& v3 = V[facetVerts[s + (j + 1 == n ? 0 : j + 1)]];
            const Vec3 Ni = cross(sub(v2, v1), sub(v3, v1));
            const double magN = mag(Ni);
            C = add(C, mul(divs(add(add(v1, v2), v3), 3.0), magN));
            Ntot += magN;
            N = add(N, Ni);
        }
        FC[i] = divs(C, Ntot);
        const Vec3 v = sub(FC[i], CC[FOC[i]]);
        if (dot(v, N) < 0) N = Vec3{-N[0], -N[1], -N[2]};
        FN[i] = divs(N, 2.0);
    }
    for (u32 i = 0; i < nBCS; i++) {
        double Vt = 0;
        Vec3 C{0, 0, 0};
        for (u32 q = cellStart[i]; q < cellStart[i + 1]; q++) {
            const u32 fi = cellFaces[q];
            const Vec3 v = sub(CC[i], FC[fi]);
            const double Vi = std::fabs(dot(v, FN[fi]));
            C = add(C, divs(mul(add(mul(FC[fi], 3.0), CC[i]), Vi), 4.0));
            Vt += Vi;
        }
        CC[i] = divs(C, Vt);
        CV[i] = Vt / 3.0;
    }
    for (u32 i = nBCS; i < nc; i++) {
        const u32 fi = cellFaces[cellStart[i]];
        CV[i] = CV[FOC[fi]];
        CC[i] = FC[fi];
    }
}

void MeshTopo::remove_boundary(const std::vector<u32>& fs) {
    if (fs.empty()) return;
    const u32 nf = nFacets(), nc = nCells();
    std::vector<char> dead(nf, 0);
    for (u32 f : fs) dead[f] = 1;
    std::vector<u32> idf(nf, MAX_INT), idc(nc, MAX_INT);
    u32 cnt = 0;
    for (u32 i = 0; i < nf; i++) if (!dead[i]) idf[i] = cnt++;
    // cells: drop dead faces; a cell left without faces (the ghost cell of a deleted face) disappears
    std::vector<u32> ncs{0}, ncf, nid;
    cnt = 0;
    for (u32 i = 0; i < nc; i++) {
        const size_t before = ncf.size();
        for (u32 q = cellStart[i]; q < cellStart[i + 1]; q++)
            if (!dead[cellFaces[q]]) { ncf.push_back(idf[cellFaces[q]]); nid.push_back(cellFaceID[q]); }
        if (ncf.size() == before) continue;
        idc[i] = cnt++;
        ncs.push_back((u32)ncf.size());
    }
    std::vector<u32> nfs{0}, nfv, nFOC, nFNC, nFMC;
    std::vector<Vec3> nFC, nFN, nCC;
    std::vector<double> nCV;
    for (u32 i = 0; i < nf; i++) {
        if (dead[i]) continue;
        for (u32 q = facetStart[i]; q < facetStart[i + 1]; q++) nfv.push_back(facetVerts[q]);
        nfs.push_back((u32)nfv.size());
        nFOC.push_back(idc[FOC[i]]);
        nFNC.push_back(idc[FNC[i]]);
        nFMC.push_back(FMC[i]);
        nFC.push_back(FC[i]);
        nFN.push_back(FN[i]);
    }
    for (u32 i = 0; i < nc; i++)
        if (idc[i] != MAX_INT) { nCC.push_back(CC[i]); nCV.push_back(CV[i]); }
    facetStart.swap(nfs); facetVerts.swap(nfv);
    cellStart.swap(ncs); cellFaces.swap(ncf); cellFaceID.swap(nid);
    FOC.swap(nFOC); FNC.swap(nFNC); FMC.swap(nFMC);
    FC.swap(nFC); FN.swap(nFN); CC.swap(nCC); CV.swap(nCV);
    for (auto& kv : boundaries)
        for (auto& f : kv.second) f = idf[f];
}

void MeshTopo::load(const Grid& g) {
    V = g.V;
    facetStart = g.facetStart; facetVerts = g.facetVerts;
    cellStart = g.cellStart; cellFaces = g.cellFaces;
    boundaries = g.boundaries;
    add_boundary_cells();
    fix_hex_cells();
    calc_geometry();
    std::vector<u32> del = boundaries["delete"];
    boundaries.erase("delete");
    remove_boundary(del);
    for (auto it = boundaries.begin(); it != boundaries.end();) {
        if (it->second.empty() || it->first.find("interior") != std::string::npos) it = boundaries.erase(it);
        else ++it;
    }
}

// ---------------------------------------------------------------------------------------------------------
// structured box grid in the reference block mesher's cell/face order (hexMesh.cpp:227-342)
// ---------------------------------------------------------------------------------------------------------
Grid box_grid(const int n[3], const double lo[3], const double hi[3], const std::array<std::string, 6>& patches,
              void (*vertex_map)(Vec3&, const void*), const void* map_arg) {
    const u32 nx = n[0], ny = n[1], nz = n[2], vx = nx + 1, vy = ny + 1, vz = nz + 1;
    Grid g;
    g.V.resize((size_t)vx * vy * vz);
    auto lin = [](double a, double b, u32 i, u32 m) { return a + (b - a) * ((double)i / (double)m); };
    for (u32 i = 0; i < vx; i++)
        for (u32 j = 0; j < vy; j++)
            for (u32 k = 0; k < vz; k++) {
                Vec3 v{lin(lo[0], hi[0], i, nx), lin(lo[1], hi[1], j, ny), lin(lo[2], hi[2], k, nz)};
                if (vertex_map) vertex_map(v, map_arg);
                g.V[((size_t)i * vy + j) * vz + k] = v;
            }
    auto vid = [&](u32 i, u32 j, u32 k) { return (u32)(((size_t)i * vy + j) * vz + k); };
    auto quad = [&](u32 a, u32 b, u32 c, u32 d) {
        g.facetVerts.push_back(a); g.facetVerts.push_back(b); g.facetVerts.push_back(c); g.facetVerts.push_back(d);
        g.facetStart.push_back((u32)g.facetVerts.size());
        return g.nFacets() - 1;
    };
    const size_t nfz = (size_t)nx * ny * vz, nfy = (size_t)nx * vy * nz;
    g.facetVerts.reserve(4 * (nfz + nfy + (size_t)vx * ny * nz));
    auto fz = [&](u32 i, u32 j, u32 k) { return (u32)(((size_t)i * ny + j) * vz + k); };
    auto fy = [&](u32 i, u32 j, u32 k) { return (u32)(nfz + ((size_t)i * vy + j) * nz + k); };
    auto fx = [&](u32 i, u32 j, u32 k) { return (u32)(nfz + nfy + ((size_t)i * ny + j) * nz + k); };
    for (u32 i = 0; i < nx; i++) for (u32 j = 0; j < ny; j++) for (u32 k = 0; k < vz; k++)
        quad(vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k));
    for (u32 i = 0; i < nx; i++) for (u32 j = 0; j < vy; j++) for (u32 k = 0; k < nz; k++)
        quad(vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j, k + 1), vid(i, j, k + 1));
    for (u32 i = 0; i < vx; i++) for (u32 j = 0; j < ny; j++) for (u32 k = 0; k < nz; k++)
        quad(vid(i, j, k), vid(i, j + 1, k), vid(i, j + 1, k + 1), vid(i, j, k + 1));
    g.cellFaces.reserve((size_t)nx * ny * nz * 6);
    for (u32 i = 0; i < nx; i++) for (u32 j = 0; j < ny; j++) for (u32 k = 0; k < nz; k++) {
        const u32 f[6] = {fz(i, j, k), fz(i, j, k + 1), fy(i, j, k), fy(i, j + 1, k), fx(i, j, k), fx(i + 1, j, k)};
        for (u32 q : f) g.cellFaces.push_back(q);
        g.cellStart.push_back((u32)g.cellFaces.size());
    }
    // sides in the order x-,x+,y-,y+,z-,z+
    auto addp = [&](const std::string& name, u32 f) { if (!name.empty()) g.boundaries[name].push_back(f); };
    for (u32 j = 0; j < ny; j++) for (u32 k = 0; k < nz; k++) addp(patches[0], fx(0, j, k));
    for (u32 j = 0; j < ny; j++) for (u32 k = 0; k < nz; k++) addp(patches[1], fx(nx, j, k));
    for (u32 i = 0; i < nx; i++) for (u32 k = 0; k < nz; k++) addp(patches[2], fy(i, 0, k));
    for (u32 i = 0; i < nx; i++) for (u32 k = 0; k < nz; k++) addp(patches[3], fy(i, ny, k));
    for (u32 i = 0; i < nx; i++) for (u32 j = 0; j < ny; j++) addp(patches[4], fz(i, j, 0));
    for (u32 i = 0; i < nx; i++) for (u32 j = 0; j < ny; j++) addp(patches[5], fz(i, j, nz));
    return g;
}

}  // namespace nsemh
