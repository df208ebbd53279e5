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
// Non-conforming (2:1 AMR) cells -- more than six faces, several coplanar sub-facets per side -- take the general
// route of fixHexCells: coplanar facets are grouped and merged into one polygon per side (coplanarFaces /
// mergeFacets / mergeFacetsGroup, mesh.cpp:791-1020), the side ids and corner order come from the merged polygons,
// every sub-facet is re-oriented against its side's polygon, and facets that share a local id are flagged in gFMC
// (1: the owner cell is the fine one, 2: the neighbour is; mesh.cpp:431-444).
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


// ---------------------------------------------------------------------------------------------------------
// predicates and polygon merging of the non-conforming route (mesh.cpp:791-1020)
// ---------------------------------------------------------------------------------------------------------
namespace {
constexpr double EQ_EPS = 1e-7;      // Constants::EqualEpsilon (tensor.h:462)
inline bool equal_eps(double p, double q) {                          // tensor.h:469-474
    const double d = std::fabs(p - q);
    return d <= EQ_EPS || d <= EQ_EPS * std::fabs(p) || d <= EQ_EPS * std::fabs(q);
}
inline double dot_lr(const Vec3& a, const Vec3& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline bool contains(const std::vector<u32>& f, u32 v) { return std::find(f.begin(), f.end(), v) != f.end(); }
inline void rotate_left(std::vector<u32>& f, long k) {
    const long n = (long)f.size();
    if (n == 0) return;
    k = ((k % n) + n) % n;
    std::rotate(f.begin(), f.begin() + k, f.end());
}
}  // namespace

struct PolyOps {
    const std::vector<Vec3>& V;
    bool point_in_line(const Vec3& v, const Vec3& v1, const Vec3& v2) const {
        const Vec3 p = sub(v, v1), q = sub(v, v2);
        if (!equal_eps(p[1] * q[2] - p[2] * q[1], 0.0)) return false;
        if (!equal_eps(p[2] * q[0] - p[0] * q[2], 0.0)) return false;
        if (!equal_eps(p[0] * q[1] - p[1] * q[0], 0.0)) return false;
        const double e = dot_lr(sub(v, v2), sub(v1, v2));
        if (e > 0) {
            const double e1 = dot_lr(sub(v1, v2), sub(v1, v2));
            if (e < e1) return true;
        }
        return false;
    }
    bool unit_normal(const std::vector<u32>& f, Vec3& n) const {
        const Vec3 &v1 = V[f[0]], &v2 = V[f[1]];
        for (size_t j = 1; j < f.size(); j++) {
            const Vec3& v3 = V[f[f.size() - j]];
            if (!point_in_line(v2, v1, v3)) {
                const Vec3 c = cross(sub(v2, v1), sub(v3, v1));
                n = divs(c, std::sqrt(dot_lr(c, c)));
                return true;
            }
        }
        return false;
    }
    bool coplanar(const std::vector<u32>& f1, const std::vector<u32>& f2) const {
        Vec3 n1, n2;
        if (!unit_normal(f1, n1) || !unit_normal(f2, n2)) throw Error("degenerate facet (all vertices on one line)");
        const Vec3 c = cross(n1, n2);
        if (equal_eps(dot_lr(c, c), 0.0)) {
            const Vec3 v = sub(V[f2[1]], V[f1[0]]);
            if (equal_eps(dot_lr(n1, v), 0.0)) return true;
        }
        return false;
    }
    // union of two edge-sharing coplanar polygons; false when they share no edge
    bool merge(const std::vector<u32>& f1_, const std::vector<u32>& f2_, std::vector<u32>& out) const {
        std::vector<u32> f1 = f1_, f2 = f2_;
        Vec3 n1, n2;
        if (!unit_normal(f1_, n1) || !unit_normal(f2_, n2)) throw Error("degenerate facet (all vertices on one line)");
        if (dot_lr(n1, n2) < 0) std::reverse(f2.begin() + 1, f2.end());
        long v1 = -1, v2 = -1;
        for (size_t i = 0; i < f1.size(); i++) if (!contains(f2, f1[i])) { v1 = (long)i; break; }
        for (size_t i = 0; i < f2.size(); i++) if (!contains(f1, f2[i])) { v2 = (long)i; break; }
        bool contained = false;
        rotate_left(f1, v1);
        if (v2 != -1) rotate_left(f2, v2);
        else contained = true;
        size_t a[2] = {0, 0}, b[2] = {0, 0};
        int count = 0;
        for (size_t i = 0; i < f1.size(); i++)
            for (size_t j = 0; j < f2.size(); j++)
                if (f1[i] == f2[j]) {
                    if (count == 0) { a[0] = i; b[0] = j; }
                    else { a[1] = i; b[1] = j; }
                    count++;
                }
        if (count < 2) return false;
        std::vector<u32> f(f1.begin(), f1.begin() + a[0] + 1);
        if (contained) {
            for (size_t j = b[0] + 1; j < b[1]; j++) f.push_back(f2[j]);
        } else {
            for (size_t j = b[0] + 1; j < f2.size(); j++) f.push_back(f2[j]);
            for (size_t j = 0; j < b[1] && j < f2.size(); j++) f.push_back(f2[j]);
        }
        for (size_t i = a[1]; i < f1.size(); i++) f.push_back(f1[i]);
        while (point_in_line(V[f[0]], V[f.back()], V[f[1]])) rotate_left(f, 1);
        out.swap(f);
        return true;
    }
};

// corner vertices of the two (possibly merged, > 4 vertices) polygons: vertices where the boundary turns by more than
// pi/16, then the pairing of getHexCorners (mesh.cpp:113-157)
void MeshTopo::hex_corners_poly(const std::vector<u32>& f1, const std::vector<u32>& f2, u32 out[8]) const {
    const double PI = 3.14159265358979323846264, tol = PI / 16;
    std::vector<u32> fm[2];
    const std::vector<u32>* fk2[2] = {&f1, &f2};
    for (int w = 0; w < 2; w++) {
        const std::vector<u32>& fk = *fk2[w];
        size_t i0 = fk.size() - 1;
        for (size_t i = 0; i < fk.size(); i++) {
            if (fm[w].size() >= 4) break;
            const size_t i1 = (i == fk.size() - 1) ? 0 : i + 1;
            const Vec3 v0 = unit(sub(V[fk[i]], V[fk[i0]])), v1 = unit(sub(V[fk[i1]], V[fk[i]]));
            const double dt = std::max(-1.0, std::min(1.0, dot(v0, v1)));
            const double ang = std::acos(dt);
            if (!(ang < tol || ang >= PI - tol)) { fm[w].push_back(fk[i]); i0 = i; }
        }
        if (fm[w].size() != 4) throw Error("a cell side does not have four corners");
    }
    hex_corners(fm[0].data(), fm[1].data(), out);
}

// merged polygon of the facets of `cell` that carry local id `id` (DG::init_geom merges them the same way, dg.cpp:176-215)
std::vector<u32> MeshTopo::merged_side(u32 cell, u32 id) const {
    PolyOps ops{V};
    std::vector<u32> f;
    bool first = true;
    for (u32 q = cellStart[cell]; q < cellStart[cell + 1]; q++) {
        if (cellFaceID[q] != id) continue;
        const u32 fi = cellFaces[q];
        std::vector<u32> g(facetVerts.begin() + facetStart[fi], facetVerts.begin() + facetStart[fi + 1]);
        if (first) { f.swap(g); first = false; }
        else {
            std::vector<u32> m;
            if (!ops.merge(f, g, m)) throw Error("cell " + std::to_string(cell) + ": sub-facets of one side do not share an edge in face order");
            f.swap(m);
        }
    }
    return f;
}

void MeshTopo::fix_general_cell(u32 ci) {
    PolyOps ops{V};
    const u32 c0 = cellStart[ci], nfc = cellStart[ci + 1] - c0;
    auto facet = [&](u32 fi) { return std::vector<u32>(facetVerts.begin() + facetStart[fi], facetVerts.begin() + facetStart[fi + 1]); };
    // group coplanar faces, always seeding with the first remaining face
    std::vector<u32> rem(cellFaces.begin() + c0, cellFaces.begin() + c0 + nfc);
    std::vector<std::vector<u32>> cng, fng;
    for (int grp = 0; grp < 6; grp++) {
        if (rem.empty()) throw Error("cell " + std::to_string(ci) + " has fewer than six sides");
        std::vector<u32> g{rem[0]}, keep;
        const std::vector<u32> seed = facet(rem[0]);
        for (size_t j = 1; j < rem.size(); j++) {
            if (ops.coplanar(seed, facet(rem[j]))) g.push_back(rem[j]);
            else keep.push_back(rem[j]);
        }
        rem.swap(keep);
        // mergeFacetsGroup
        std::vector<u32> fn = facet(g[0]);
        std::vector<u32> rest(g.begin() + 1, g.end());
        while (!rest.empty()) {
            std::vector<u32> left;
            bool repeat = false, any = false;
            for (u32 fi : rest) {
                std::vector<u32> m;
                if (ops.merge(fn, facet(fi), m)) { fn.swap(m); any = true; }
                else { repeat = true; left.push_back(fi); }
            }
            rest.swap(left);
            if (!repeat) break;
            if (!any) throw Error("cell " + std::to_string(ci) + ": coplanar facets cannot be merged into one side");
        }
        cng.push_back(g);
        fng.push_back(fn);
    }
    if (!rem.empty()) throw Error("cell " + std::to_string(ci) + " has more than six sides");
    int gid[6] = {-1, -1, -1, -1, -1, -1};
    const std::vector<u32>& f0 = fng[0];
    for (int j = 0; j < 6; j++) {
        if (gid[j] >= 0) continue;
        const std::vector<u32>& fj = fng[j];
        int id = j;
        if (j >= 1) {
            int n01 = 0, n0l = 0;
            for (u32 x : fj) { if (x == f0[0] || x == f0[1]) n01++; if (x == f0[0] || x == f0.back()) n0l++; }
            if (n01 >= 2) id = 2;
            else if (n0l >= 2) id = 4;
            else continue;
        }
        gid[j] = id;
        for (int k = 0; k < 6; k++) {
            if (gid[k] >= 0) continue;
            bool shares = false;
            for (u32 x : fj) if (contains(fng[k], x)) { shares = true; break; }
            if (!shares) { gid[k] = id ^ 1; break; }
        }
    }
    for (int q = 0; q < 6; q++)
        if (gid[q] < 0) throw Error("cell " + std::to_string(ci) + " is not a hexahedron with 3 pairs of opposite sides");
    const Vec3 N = cross(sub(V[f0[1]], V[f0[0]]), sub(V[f0.back()], V[f0[0]]));
    const Vec3 e = sub(V[fng[1][0]], V[f0[0]]);
    if (dot_lr(N, e) < 0)
        for (int q = 0; q < 6; q++) gid[q] = (gid[q] == 0) ? 1 : (gid[q] == 1 ? 0 : gid[q]);
    int i0 = 0, i1 = 0;
    for (int q = 0; q < 6; q++) { if (gid[q] == 0) i0 = q; else if (gid[q] == 1) i1 = q; }
    const u32 fa = cng[i0][0], fb = cng[i1][0];
    const u32 i0n = (FNC[fa] != ci) ? FNC[fa] : FOC[fa];
    const u32 i1n = (FNC[fb] != ci) ? FNC[fb] : FOC[fb];
    const bool flip = (i0n > i1n) && (ci >= i0n || ci >= i1n);
    u32 vp[8];
    if (!flip) hex_corners_poly(fng[i0], fng[i1], vp);
    else {
        u32 t[8];
        hex_corners_poly(fng[i1], fng[i0], t);
        for (int q = 0; q < 4; q++) { vp[q] = t[q + 4]; vp[q + 4] = t[q]; }
    }
    const u32 rots[6] = {vp[0], vp[4], vp[0], vp[3], vp[0], vp[1]};
    const u32 rote[6] = {vp[1], vp[5], vp[1], vp[2], vp[3], vp[2]};
    for (int i = 0; i < 6; i++) {
        std::vector<u32>& fn = fng[i];
        const u32 rs = rots[gid[i]], re = rote[gid[i]];
        const auto itp = std::find(fn.begin(), fn.end(), rs);
        if (itp == fn.end()) throw Error("cell " + std::to_string(ci) + ": side does not contain its reference corner");
        rotate_left(fn, (long)(itp - fn.begin()));
        const double d = dot_lr(unit(sub(V[fn[1]], V[fn[0]])), unit(sub(V[re], V[rs])));
        if (d < 0.99) std::reverse(fn.begin() + 1, fn.end());
        // orient every sub-facet of the side against the side's polygon
        for (size_t j = 0; j < cng[i].size(); j++) {
            const u32 fid = cng[i][j];
            u32* fbeg = &facetVerts[facetStart[fid]];
            const size_t fl = facetStart[fid + 1] - facetStart[fid];
            auto in_f = [&](u32 v) { return std::find(fbeg, fbeg + fl, v) != fbeg + fl; };
            size_t it1 = 0;
            while (it1 < fn.size() && !in_f(fn[it1])) it1++;
            if (it1 == fn.size()) throw Error("cell " + std::to_string(ci) + ": sub-facet shares no vertex with its side");
            const size_t p2 = (size_t)(std::find(fbeg, fbeg + fl, fn[it1]) - fbeg);
            std::rotate(fbeg, fbeg + p2, fbeg + fl);
            if (j >= 2) {
                const size_t sft = (j - 1) % fl;       // std::rotate(f.rbegin(), f.rbegin() + (j-1), f.rend()): right by j-1
                if (sft) std::rotate(fbeg, fbeg + (fl - sft), fbeg + fl);
            }
            for (size_t k = 1; k < fl; k++) {
                const auto it = std::find(fn.begin(), fn.end(), fbeg[k]);
                if (it == fn.end()) continue;
                const size_t it3 = (size_t)(it - fn.begin());
                if (it1 > it3) { std::reverse(fbeg + 1, fbeg + fl); break; }
                it1 = it3;
            }
        }
    }
    // rewrite the face list in id order (sub-facets of a side stay in their group order)
    size_t w = 0;
    std::vector<u32> ids;
    for (int want = 0; want < 6; want++)
        for (int i = 0; i < 6; i++)
            if (gid[i] == want)
                for (u32 fid : cng[i]) { cellFaces[c0 + w] = fid; cellFaceID[c0 + w] = (u32)want; ids.push_back((u32)want); w++; }
    if (nfc > 6)
        for (u32 j = 0; j < nfc; j++)
            if (std::count(ids.begin(), ids.end(), ids[j]) > 1) {
                const u32 f = cellFaces[c0 + j];
                FMC[f] = (FOC[f] == ci) ? 2u : 1u;
            }
}

void MeshTopo::fix_hex_cells() {
    FMC.assign(nFacets(), 0);
    cellFaceID.assign(cellFaces.size(), 0);
    for (u32 ci = 0; ci < nBCS; ci++) {
        const u32 c0 = cellStart[ci];
        // conforming hexahedron (six quadrilaterals): the short route below; anything else: the general route
        bool simple = (cellStart[ci + 1] - c0 == 6);
        for (u32 q = c0; simple && q < cellStart[ci + 1]; q++)
            simple = (facetStart[cellFaces[q] + 1] - facetStart[cellFaces[q]] == 4);
        if (!simple) { fix_general_cell(ci); continue; }
        u32 fc[6];
        u32* fv[6];
        for (int q = 0; q < 6; q++) {
            fc[q] = cellFaces[c0 + q];
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
    if (spherical) sphere_geometry();
    for (u32 i = nBCS; i < nc; i++) {
        const u32 fi = cellFaces[cellStart[i]];
        CV[i] = CV[FOC[fi]];
        CC[i] = FC[fi];
    }
}

// ---------------------------------------------------------------------------------------------------------
// cubed-sphere shells (Mesh::is_spherical): vertices projected onto two radii, then centres, face areas and cell
// volumes corrected to the curved elements
// ---------------------------------------------------------------------------------------------------------
namespace {
// (radius, latitude, longitude), tensor.h:598-605
inline Vec3 cart_to_sphere(const Vec3& c) {
    return Vec3{mag(c), std::atan2(c[2], std::sqrt(c[0] * c[0] + c[1] * c[1])), std::atan2(c[1], c[0])};
}
// tensor.h:608-612
inline double geodesic_distance(const Vec3& s1, const Vec3& s2) {
    double d = (s1[0] + s2[0]) / 2;
    d *= std::acos(std::sin(s1[1]) * std::sin(s2[1]) + std::cos(s1[1]) * std::cos(s2[1]) * std::cos(s1[2] - s2[2]));
    return d;
}
// tensor.h:624-635
inline double spherical_triangle_area(double radius, const Vec3& v0, const Vec3& v1, const Vec3& v2) {
    const Vec3 a = unit(v0), b = unit(v1), c = unit(v2);
    double t = std::fabs(dot(a, cross(b, c)));
    t /= (1 + dot(a, b) + dot(b, c) + dot(a, c));
    return 2 * std::atan(t) * radius * radius;
}
// equal(Scalar, Scalar) with Constants::EqualEpsilon = 1e-7, tensor.h:460-480
inline bool near(double p, double q) {
    const double tol = 1e-7, delta = std::fabs(p - q);
    return delta <= tol || delta <= tol * std::fabs(p) || delta <= tol * std::fabs(q);
}
}  // namespace

// Mesh::MeshObject::ExtrudeMesh, mesh.cpp:723-750: the grid file holds a shell between two concentric cubes; the cube a vertex lies on
// (its largest |coordinate|) picks the radius it is projected to.  The SMALLER cube goes to the OUTER radius, as in the reference.
void MeshTopo::extrude() {
    double minh = 1e30, maxh = 0;
    auto height = [](const Vec3& v) { return std::max(std::max(std::fabs(v[0]), std::fabs(v[1])), std::fabs(v[2])); };
    for (const Vec3& v : V) {
        const double h = height(v);
        if (h > maxh) maxh = h;
        if (h < minh) minh = h;
    }
    if (shell_h[1] > 0) { minh = shell_h[0]; maxh = shell_h[1]; }
    const double radiusi = sphere_radius, radiuso = sphere_radius + sphere_height;
    for (Vec3& v : V) {
        const double f = (height(v) - minh) / (maxh - minh);
        v = mul(unit(v), f * radiusi + (1 - f) * radiuso);
    }
}

// mesh.cpp:520-570.  Sides 0 and 1 of every cell are its two radial faces (the block mesher's third direction).
void MeshTopo::sphere_geometry() {
    for (u32 i = 0; i < nBCS; i++) {
        const u32* c = &cellFaces[cellStart[i]];
        // a cell next to refined ones lists more than six facets (its split sides); like the reference the loop below visits entries 2..5
        // only -- every sub-facet is one of the entries 2..5 of the fine cell on its other side
        if (cellStart[i + 1] - cellStart[i] < 6) throw Error("spherical meshes need hexahedral cells (cell " + std::to_string(i) + ")");
        const double radiusb = mag(V[facetVerts[facetStart[c[0]]]]);
        const double radiust = mag(V[facetVerts[facetStart[c[1]]]]);
        CC[i] = mul(CC[i], (radiusb + radiust) / (2 * mag(CC[i])));
        FC[c[0]] = mul(FC[c[0]], radiusb / mag(FC[c[0]]));
        FC[c[1]] = mul(FC[c[1]], radiust / mag(FC[c[1]]));
        const int nfac = (int)(cellStart[i + 1] - cellStart[i]);
        for (int j = 2; j < nfac; j++) {
            // The reference stops at entry 5 (mesh.cpp:529): a further sub-facet of a split side is corrected from the cell on its other
            // side.  In a partition that cell can be on another rank; then, and only then, it is done from here (same radii: one layer)
            if (j >= 6 && std::min(FOC[c[j]], FNC[c[j]]) < nBCS && std::max(FOC[c[j]], FNC[c[j]]) < nBCS) continue;
            FC[c[j]] = mul(FC[c[j]], (radiusb + radiust) / (2 * mag(FC[c[j]])));
            // area of a vertical face: half the geodesic length of its outline times the shell thickness
            double d = 0;
            const u32 s = facetStart[c[j]], n = facetStart[c[j] + 1] - s;
            for (u32 k = 0; k < n; k++) {
                const Vec3& v0 = V[facetVerts[s + k]];
                const Vec3& v1 = V[facetVerts[s + (k == n - 1 ? 0 : k + 1)]];
                const Vec3 r0 = unit(v0), r1 = unit(v1);
                if (near(r0[0], r1[0]) && near(r0[1], r1[1]) && near(r0[2], r1[2])) continue;
                d += geodesic_distance(cart_to_sphere(v0), cart_to_sphere(v1));
            }
            d /= 2;
            const double area = std::fabs(radiust - radiusb) * d;
            FN[c[j]] = mul(unit(FN[c[j]]), area);
        }
    }
    for (u32 i = 0; i < nBCS; i++) {
        const u32* c = &cellFaces[cellStart[i]];
        double area = 0;
        for (int k = 0; k < 2; k++) {
            const u32 s = facetStart[c[k]], n = facetStart[c[k] + 1] - s;
            const double radius = mag(V[facetVerts[s]]);
            const Vec3& fc = FC[c[k]];
            double a = 0;
            for (u32 j = 0; j < n; j++)
                a += spherical_triangle_area(radius, V[facetVerts[s + j]], V[facetVerts[s + (j == n - 1 ? 0 : j + 1)]], fc);
            FN[c[k]] = mul(unit(FN[c[k]]), a);
            area += a;
        }
        area /= 2;
        const double radiusb = mag(V[facetVerts[facetStart[c[0]]]]);
        const double radiust = mag(V[facetVerts[facetStart[c[1]]]]);
        CV[i] = std::fabs(radiust - radiusb) * area;
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

void MeshTopo::load_flags_only(const Grid& g) {
    V = g.V;
    facetStart = g.facetStart; facetVerts = g.facetVerts;
    cellStart = g.cellStart; cellFaces = g.cellFaces;
    boundaries = g.boundaries;
    add_boundary_cells();
    fix_hex_cells();
}

std::vector<u32> mortar_flags(const Grid& g) {
    bool conforming = true;
    for (u32 c = 0; c < g.nCells() && conforming; c++) conforming = (g.cellStart[c + 1] - g.cellStart[c] == 6);
    if (conforming) return std::vector<u32>(g.nFacets(), 0);
    MeshTopo t;
    t.load_flags_only(g);
    return t.FMC;
}

void MeshTopo::load(const Grid& g) {
    V = g.V;
    facetStart = g.facetStart; facetVerts = g.facetVerts;
    cellStart = g.cellStart; cellFaces = g.cellFaces;
    boundaries = g.boundaries;
    add_boundary_cells();
    fix_hex_cells();
    if (spherical && !no_extrude) extrude();
    calc_geometry();
    if (!keep_empty) {
        std::vector<u32> del = boundaries["delete"];
        boundaries.erase("delete");
        remove_boundary(del);
    }
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
