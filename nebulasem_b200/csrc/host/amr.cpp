// amr.cpp -- adaptive regrid of hexahedral grids in memory: what Prepare::refineMesh (tagging, src/field/field.cpp:625-858) and
// MeshObject::refineMesh (src/mesh/mesh.cpp:2216-2748) do through grid_<k>/amrTree_<k> files in the reference.
//
// Not a restatement of the reference's facet-splitting code: a forest of octrees over the cells of the loaded grid.  A split cell gets its
// children from the 27-point (9-point in 2-D) lattice of corner, edge-midpoint, face-centre and cell-centre vertices; vertices are shared
// through maps keyed by the parent vertices, facets through a map keyed by their four corners.  After every regrid the whole grid is
// re-emitted in the reference's grid format (vertices, facets as polygons with the hanging vertices inserted into their edges like
// MeshObject::breakEdges does, cells as facet lists in which a coarse side next to finer cells is its 2 or 4 sub-facets, boundary patches),
// so it loads through MeshTopo::load like a grid the reference wrote, and the regrid is reported as the refineMap / coarseMap / cellMap
// triple of MeshObject::refineMesh that nsem_refine_state consumes.  Cell order and local frames differ from the reference's; the grid
// is the same set of cells (tests/test_amr_regrid.py compares centroids, volumes and mortar faces with the reference's own regrid).
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "nsem_host.h"

namespace nsemh {

namespace {
using Key2 = std::array<u32, 2>;
using Key4 = std::array<u32, 4>;
Key2 key2(u32 a, u32 b) { return a < b ? Key2{a, b} : Key2{b, a}; }
Key4 key4(u32 a, u32 b, u32 c, u32 d) { Key4 k{a, b, c, d}; std::sort(k.begin(), k.end()); return k; }
// corners c0..c7 <-> (xi, eta, zeta) bits: c0 000, c1 100, c2 110, c3 010, c4 001, c5 101, c6 111, c7 011
const int kCornerBits[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
int corner_of(int a, int b, int c) {
    for (int q = 0; q < 8; q++)
        if (kCornerBits[q][0] == a && kCornerBits[q][1] == b && kCornerBits[q][2] == c) return q;
    return -1;
}
// sides in the facet order of a cell: zeta-, zeta+, eta-, eta+, xi-, xi+ (the block mesher's order, hexMesh.cpp:322-342)
const int kSide[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {3, 2, 6, 7}, {0, 3, 7, 4}, {1, 2, 6, 5}};
const int kSideAxis[6] = {2, 2, 1, 1, 0, 0};
const int kSideHigh[6] = {0, 1, 0, 1, 0, 1};
}  // namespace

void AmrForest::init(const Grid& g, const Vec3& direction) {
    V = g.V;
    nodes.clear(); leaves.clear(); edgeMid.clear(); faceMid.clear(); patchOf.clear();
    auto facet = [&](u32 f) { return std::vector<u32>(g.facetVerts.begin() + g.facetStart[f], g.facetVerts.begin() + g.facetStart[f + 1]); };
    for (u32 c = 0; c < g.nCells(); c++) {
        const u32 c0 = g.cellStart[c], nf = g.cellStart[c + 1] - c0;
        if (nf != 6) throw Error("AmrForest: the base grid must be conforming hexahedra (cell " + std::to_string(c) + " has " + std::to_string(nf) + " facets)");
        std::vector<std::vector<u32>> fv;
        for (u32 q = 0; q < 6; q++) {
            fv.push_back(facet(g.cellFaces[c0 + q]));
            if (fv.back().size() != 4) throw Error("AmrForest: the base grid must be conforming hexahedra (non-quadrilateral facet)");
        }
        auto has = [](const std::vector<u32>& f, u32 v) { return std::find(f.begin(), f.end(), v) != f.end(); };
        int opp = -1;
        for (int q = 1; q < 6; q++) {
            bool shares = false;
            for (u32 v : fv[0]) shares |= has(fv[q], v);
            if (!shares) opp = q;
        }
        if (opp < 0) throw Error("AmrForest: cell " + std::to_string(c) + " has no facet opposite its first one");
        Node n;
        for (int k = 0; k < 4; k++) {
            const u32 a = fv[0][k];
            u32 b = MAX_INT;
            for (int q = 1; q < 6 && b == MAX_INT; q++) {
                if (q == opp || !has(fv[q], a)) continue;
                const std::vector<u32>& f = fv[q];
                const size_t p = std::find(f.begin(), f.end(), a) - f.begin();
                for (u32 cand : {f[(p + 1) % 4], f[(p + 3) % 4]})
                    if (has(fv[opp], cand)) b = cand;
            }
            if (b == MAX_INT) throw Error("AmrForest: cell " + std::to_string(c) + " is not a hexahedron");
            n.v[k] = a; n.v[k + 4] = b;
        }
        // The frame stays as the grid file has it, left-handed cells included (three of the six panels of the reference's cubed-sphere block
        // files are): cells a regrid does not touch must be re-emitted with the local axes they had, because fields are copied node by node
        // (refineField, field.h:1877-1883) and, on the sphere, even the positions of an element's interior nodes depend on which of its
        // axes is which (the radial rescale of the node placement, dg.cpp:257-285, is not symmetric in them)
        nodes.push_back(n);
        leaves.push_back((int)nodes.size() - 1);
    }
    for (const auto& kv : g.boundaries)
        for (u32 f : kv.second) {
            const std::vector<u32> q = facet(f);
            if (q.size() == 4) patchOf[key4(q[0], q[1], q[2], q[3])] = kv.first;
        }
    // 2-D refinement: the local axis along `direction` is never split (Mesh::amr_direction, field.cpp:633)
    dir = direction;
}

int AmrForest::split_mask(const Node& n) const {
    double dl = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    if (dl == 0) return 7;
    Vec3 dir = this->dir;
    if (spherical) {
        // on a cubed sphere the axis that is never split is the radial one: uDir = unit(cell centre) (mesh.cpp:1299)
        dir = Vec3{0, 0, 0};
        for (int k = 0; k < 8; k++) for (int d = 0; d < 3; d++) dir[d] += V[n.v[k]][d];
        dl = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        if (dl == 0) return 7;
    }
    const int far[3] = {1, 3, 4};
    int best = 0;
    double bestv = -1;
    for (int a = 0; a < 3; a++) {
        const Vec3& p = V[n.v[far[a]]];
        const Vec3& o = V[n.v[0]];
        const double e[3] = {p[0] - o[0], p[1] - o[1], p[2] - o[2]};
        const double el = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        const double c = std::fabs(e[0] * dir[0] + e[1] * dir[1] + e[2] * dir[2]) / (el * dl);
        if (c > bestv) { bestv = c; best = a; }
    }
    return 7 & ~(1 << best);
}

// MeshObject::calcFaceCenter (mesh.cpp:1166-1188): fan of triangles around the vertex average, centroids weighted by the triangle areas.
// On a parallelogram this is the vertex average; on a general quadrilateral (terrain-following grids) it is not.  `area2` = sum of the
// fan's cross products (twice the area vector).
static Vec3 corrected_face_centre(const std::vector<Vec3>& V, const u32 q[4], Vec3* area2 = nullptr) {
    Vec3 c0{0, 0, 0};
    for (int j = 0; j < 4; j++) for (int d = 0; d < 3; d++) c0[d] += V[q[j]][d];
    for (int d = 0; d < 3; d++) c0[d] /= 4.0;
    Vec3 ct{0, 0, 0}, nsum{0, 0, 0};
    double ntot = 0;
    for (int j = 0; j < 4; j++) {
        const Vec3 &v2 = V[q[j]], &v3 = V[q[(j + 1) % 4]];
        const double a[3] = {v2[0] - c0[0], v2[1] - c0[1], v2[2] - c0[2]}, b[3] = {v3[0] - c0[0], v3[1] - c0[1], v3[2] - c0[2]};
        const double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        const double m = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int d = 0; d < 3; d++) { ct[d] += m * ((c0[d] + v2[d] + v3[d]) / 3.0); nsum[d] += n[d]; }
        ntot += m;
    }
    if (area2) *area2 = nsum;
    if (ntot == 0) return c0;
    return Vec3{ct[0] / ntot, ct[1] / ntot, ct[2] / ntot};
}

u32 AmrForest::mid_vertex(const std::vector<u32>& of) {
    // the vertex in the middle of 2 (edge), 4 (face) or 8 (cell) parent corners, placed where MeshObject::refineMesh places it: edge
    // midpoint (mesh.cpp:1296), corrected face centre (refineFacet, mesh.cpp:1251-1253), cell centroid (calcCellCenter, mesh.cpp:1192-1243,
    // used at mesh.cpp:1647).  `of` lists the corners in lexicographic (x, y, z) order; edges and faces are shared through the maps
    auto make = [&]() {
        Vec3 s{0, 0, 0};
        if (of.size() == 4) {
            const u32 q[4] = {of[0], of[1], of[3], of[2]};                        // lexicographic -> around the face
            s = corrected_face_centre(V, q);
        } else if (of.size() == 8) {
            static const int kFace[6][4] = {{0, 1, 3, 2}, {4, 5, 7, 6}, {0, 1, 5, 4}, {2, 3, 7, 6}, {0, 2, 6, 4}, {1, 3, 7, 5}};
            Vec3 c0{0, 0, 0};
            for (u32 v : of) for (int d = 0; d < 3; d++) c0[d] += V[v][d];
            for (int d = 0; d < 3; d++) c0[d] /= 8.0;
            Vec3 ct{0, 0, 0};
            double vt = 0;
            for (int f = 0; f < 6; f++) {
                const u32 q[4] = {of[kFace[f][0]], of[kFace[f][1]], of[kFace[f][2]], of[kFace[f][3]]};
                Vec3 a2;
                const Vec3 fc = corrected_face_centre(V, q, &a2);
                const double vi = std::fabs((c0[0] - fc[0]) * a2[0] + (c0[1] - fc[1]) * a2[1] + (c0[2] - fc[2]) * a2[2]) / 2.0;   // 3 x pyramid volume
                for (int d = 0; d < 3; d++) ct[d] += vi * (3 * fc[d] + c0[d]) / 4.0;
                vt += vi;
            }
            s = vt == 0 ? c0 : Vec3{ct[0] / vt, ct[1] / vt, ct[2] / vt};
        } else {
            for (u32 v : of) for (int d = 0; d < 3; d++) s[d] += V[v][d];
            for (int d = 0; d < 3; d++) s[d] /= (double)of.size();
        }
        V.push_back(s);
        return (u32)V.size() - 1;
    };
    if (of.size() == 2) {
        auto it = edgeMid.find(key2(of[0], of[1]));
        if (it != edgeMid.end()) return it->second;
        return edgeMid[key2(of[0], of[1])] = make();
    }
    if (of.size() == 4) {
        auto it = faceMid.find(key4(of[0], of[1], of[2], of[3]));
        if (it != faceMid.end()) return it->second;
        return faceMid[key4(of[0], of[1], of[2], of[3])] = make();
    }
    return make();
}

void AmrForest::split(int ni) {
    const int mask = split_mask(nodes[ni]);
    const Node parent = nodes[ni];
    const int np[3] = {(mask & 1) ? 3 : 2, (mask & 2) ? 3 : 2, (mask & 4) ? 3 : 2};     // lattice points per axis
    // lattice coordinate -> parent coordinate in {0, 1, 2} (2 = far corner, 1 = middle)
    auto pc = [&](int axis, int i) { return np[axis] == 3 ? i : 2 * i; };
    u32 L[3][3][3];
    for (int a = 0; a < np[0]; a++)
        for (int b = 0; b < np[1]; b++)
            for (int c = 0; c < np[2]; c++) {
                const int P[3] = {pc(0, a), pc(1, b), pc(2, c)};
                std::vector<u32> of;
                // the parent corners this lattice point lies between
                for (int x = 0; x < 2; x++) for (int y = 0; y < 2; y++) for (int z = 0; z < 2; z++) {
                    const int q[3] = {x, y, z};
                    bool ok = true;
                    for (int d = 0; d < 3; d++) ok &= (P[d] == 1) || (P[d] == 2 * q[d]);
                    if (ok) of.push_back(parent.v[corner_of(x, y, z)]);
                }
                L[a][b][c] = (of.size() == 1) ? of[0] : mid_vertex(of);
            }
    const int first = (int)nodes.size();
    int count = 0;
    for (int a = 0; a + 1 < np[0]; a++)
        for (int b = 0; b + 1 < np[1]; b++)
            for (int c = 0; c + 1 < np[2]; c++) {
                Node ch;
                ch.level = parent.level + 1;
                ch.parent = ni;
                for (int q = 0; q < 8; q++) ch.v[q] = L[a + kCornerBits[q][0]][b + kCornerBits[q][1]][c + kCornerBits[q][2]];
                // boundary sides inherit the parent's patch
                for (int s = 0; s < 6; s++) {
                    const int idx[3] = {a, b, c};
                    const int ax = kSideAxis[s];
                    const bool onSide = kSideHigh[s] ? (idx[ax] + 2 == np[ax]) : (idx[ax] == 0);
                    if (!onSide) continue;
                    auto it = patchOf.find(key4(parent.v[kSide[s][0]], parent.v[kSide[s][1]], parent.v[kSide[s][2]], parent.v[kSide[s][3]]));
                    if (it == patchOf.end()) continue;
                    const std::string name = it->second;
                    patchOf[key4(ch.v[kSide[s][0]], ch.v[kSide[s][1]], ch.v[kSide[s][2]], ch.v[kSide[s][3]])] = name;
                }
                nodes.push_back(ch);
                count++;
            }
    nodes[ni].child0 = first;
    nodes[ni].nchild = count;
}

std::vector<int> AmrForest::levels() const {
    std::vector<int> l(leaves.size());
    for (size_t i = 0; i < leaves.size(); i++) l[i] = nodes[leaves[i]].level;
    return l;
}

// Families (all children of one parent are leaves) as lists of cells
std::vector<std::vector<u32>> AmrForest::families() const {
    std::vector<int> cellOf(nodes.size(), -1);
    for (size_t i = 0; i < leaves.size(); i++) cellOf[leaves[i]] = (int)i;
    std::vector<std::vector<u32>> out;
    for (size_t n = 0; n < nodes.size(); n++) {
        if (nodes[n].child0 < 0) continue;
        std::vector<u32> f;
        for (int k = 0; k < nodes[n].nchild; k++) {
            const int c = cellOf[nodes[n].child0 + k];
            if (c < 0) { f.clear(); break; }
            f.push_back((u32)c);
        }
        if (!f.empty()) out.push_back(f);
    }
    return out;
}

AmrForest::Maps AmrForest::regrid(const std::vector<uint8_t>& refine, const std::vector<uint8_t>& coarsen) {
    const u32 nOld = (u32)leaves.size();
    if (refine.size() != nOld || coarsen.size() != nOld) throw Error("AmrForest::regrid: one flag per cell expected");
    Maps m;
    std::vector<int> cellOf(nodes.size(), -1);
    for (u32 i = 0; i < nOld; i++) cellOf[leaves[i]] = (int)i;
    std::vector<uint8_t> gone(nOld, 0);
    std::vector<int> merged;                          // parent nodes that become leaves
    std::vector<std::vector<u32>> mergedKids;
    for (size_t n = 0; n < nodes.size(); n++) {
        if (nodes[n].child0 < 0) continue;
        std::vector<u32> kids;
        bool all = true;
        for (int k = 0; k < nodes[n].nchild && all; k++) {
            const int c = cellOf[nodes[n].child0 + k];
            all = (c >= 0) && coarsen[c] && !refine[c];
            if (all) kids.push_back((u32)c);
        }
        if (!all) continue;
        merged.push_back((int)n);
        mergedKids.push_back(kids);
        for (u32 c : kids) gone[c] = 1;
    }
    std::vector<int> splitCells;
    for (u32 i = 0; i < nOld; i++)
        if (refine[i] && !gone[i]) { splitCells.push_back((int)i); gone[i] = 1; }
    // new cell order: surviving cells in their order, then the merged parents, then the children family by family
    std::vector<int> newLeaves;
    m.cellMap.assign(nOld, MAX_INT);
    for (u32 i = 0; i < nOld; i++)
        if (!gone[i]) { m.cellMap[i] = (u32)newLeaves.size(); newLeaves.push_back(leaves[i]); }
    for (size_t f = 0; f < merged.size(); f++) {
        // children nodes stay in the tree only as history: cut them off
        nodes[merged[f]].child0 = -1;
        nodes[merged[f]].nchild = 0;
        m.coarseMap.push_back((u32)mergedKids[f].size());
        m.coarseMap.push_back((u32)m.cellMap.size());
        for (u32 c : mergedKids[f]) m.coarseMap.push_back(c);
        m.cellMap.push_back((u32)newLeaves.size());
        newLeaves.push_back(merged[f]);
    }
    for (int c : splitCells) {
        const int ni = leaves[c];
        split(ni);
        m.refineMap.push_back((u32)nodes[ni].nchild);
        m.refineMap.push_back((u32)c);
        for (int k = 0; k < nodes[ni].nchild; k++) {
            m.refineMap.push_back((u32)m.cellMap.size());
            m.cellMap.push_back((u32)newLeaves.size());
            newLeaves.push_back(nodes[ni].child0 + k);
        }
    }
    leaves.swap(newLeaves);
    return m;
}

Grid AmrForest::grid() const {
    Grid g;
    // live vertices, renumbered in forest order
    std::vector<u32> vmap(V.size(), MAX_INT);
    for (int ni : leaves) for (u32 v : nodes[ni].v) vmap[v] = 0;
    for (size_t v = 0; v < V.size(); v++)
        if (vmap[v] == 0) { vmap[v] = (u32)g.V.size(); g.V.push_back(V[v]); }
    auto live = [&](u32 v) { return vmap[v] != MAX_INT; };
    // every side of every leaf, keyed by its corners
    struct Use { int cell, side; };
    std::map<Key4, std::vector<Use>> users;
    for (size_t c = 0; c < leaves.size(); c++) {
        const Node& n = nodes[leaves[c]];
        for (int s = 0; s < 6; s++) users[key4(n.v[kSide[s][0]], n.v[kSide[s][1]], n.v[kSide[s][2]], n.v[kSide[s][3]])].push_back({(int)c, s});
    }
    auto single = [&](const Key4& k) { auto it = users.find(k); return it != users.end() && it->second.size() == 1; };
    auto emid = [&](u32 a, u32 b) -> u32 { auto it = edgeMid.find(key2(a, b)); return (it != edgeMid.end() && live(it->second)) ? it->second : MAX_INT; };
    // the finer quads that tile quad q (2:1): four around the face centre, or two across a pair of opposite edges; empty = not subdivided
    auto subquads = [&](const u32 q[4]) {
        std::vector<std::array<u32, 4>> out;
        auto fm = faceMid.find(key4(q[0], q[1], q[2], q[3]));
        const u32 m01 = emid(q[0], q[1]), m12 = emid(q[1], q[2]), m23 = emid(q[2], q[3]), m30 = emid(q[3], q[0]);
        if (fm != faceMid.end() && live(fm->second) && m01 != MAX_INT && m12 != MAX_INT && m23 != MAX_INT && m30 != MAX_INT) {
            const u32 c = fm->second;
            out = {{q[0], m01, c, m30}, {m01, q[1], m12, c}, {c, m12, q[2], m23}, {m30, c, m23, q[3]}};
        } else if (m01 != MAX_INT && m23 != MAX_INT) {
            out = {{q[0], m01, m23, q[3]}, {m01, q[1], q[2], m23}};
        } else if (m12 != MAX_INT && m30 != MAX_INT) {
            out = {{q[0], q[1], m12, m30}, {m30, m12, q[2], q[3]}};
        }
        for (const auto& s : out)
            if (!single(key4(s[0], s[1], s[2], s[3]))) return std::vector<std::array<u32, 4>>();
        return out;
    };
    std::map<Key4, u32> facetOf;
    std::function<void(u32, u32)> expand = [&](u32 a, u32 b) {      // hanging vertices strictly inside edge a-b, in order
        const u32 mm = emid(a, b);
        if (mm == MAX_INT) return;
        expand(a, mm);
        g.facetVerts.push_back(vmap[mm]);
        expand(mm, b);
    };
    auto facet_id = [&](const u32 q[4]) {
        const Key4 k = key4(q[0], q[1], q[2], q[3]);
        auto it = facetOf.find(k);
        if (it != facetOf.end()) return it->second;
        for (int e = 0; e < 4; e++) {
            g.facetVerts.push_back(vmap[q[e]]);
            expand(q[e], q[(e + 1) % 4]);
        }
        g.facetStart.push_back((u32)g.facetVerts.size());
        const u32 id = g.nFacets() - 1;
        facetOf[k] = id;
        return id;
    };
    std::map<u32, std::string> facetPatch;
    for (size_t c = 0; c < leaves.size(); c++) {
        const Node& n = nodes[leaves[c]];
        for (int s = 0; s < 6; s++) {
            const u32 q[4] = {n.v[kSide[s][0]], n.v[kSide[s][1]], n.v[kSide[s][2]], n.v[kSide[s][3]]};
            const Key4 k = key4(q[0], q[1], q[2], q[3]);
            std::vector<std::array<u32, 4>> subs;
            if (users[k].size() == 1) subs = subquads(q);
            if (!subs.empty()) {
                for (const auto& sq : subs) g.cellFaces.push_back(facet_id(sq.data()));
                continue;
            }
            const u32 f = facet_id(q);
            g.cellFaces.push_back(f);
            if (users[k].size() == 1) {
                // not shared and not subdivided: either the fine side of a 2:1 face (the coarse cell lists it too) or a boundary facet
                auto it = patchOf.find(k);
                if (it != patchOf.end()) facetPatch[f] = it->second;
            }
        }
        g.cellStart.push_back((u32)g.cellFaces.size());
    }
    // a fine facet listed by a coarse neighbour is interior even if a stale patch entry exists for its key
    std::vector<u32> uses(g.nFacets(), 0);
    for (u32 f : g.cellFaces) uses[f]++;
    for (const auto& kv : patchOf) g.boundaries[kv.second];          // keep every patch name, even when it ends up empty
    for (const auto& kv : facetPatch)
        if (uses[kv.first] == 1) g.boundaries[kv.second].push_back(kv.first);
    for (u32 f = 0; f < g.nFacets(); f++)
        if (uses[f] == 1 && !facetPatch.count(f)) throw Error("AmrForest::grid: facet " + std::to_string(f) + " has one cell and no boundary patch");
    return g;
}

// ---------------------------------------------------------------------------------------------------------
// persistence
// ---------------------------------------------------------------------------------------------------------
namespace {
const char kForestMagic[8] = {'N', 'S', 'E', 'M', 'F', 'R', 'S', '1'};
struct Writer {
    FILE* f;
    template <class T> void pod(const T& v) { if (std::fwrite(&v, sizeof(T), 1, f) != 1) throw Error("AmrForest::save: write failed"); }
    template <class T> void arr(const T* p, size_t n) { if (n && std::fwrite(p, sizeof(T), n, f) != n) throw Error("AmrForest::save: write failed"); }
};
struct Reader {
    FILE* f;
    template <class T> T pod() { T v; if (std::fread(&v, sizeof(T), 1, f) != 1) throw Error("AmrForest::load: truncated file"); return v; }
    template <class T> void arr(T* p, size_t n) { if (n && std::fread(p, sizeof(T), n, f) != n) throw Error("AmrForest::load: truncated file"); }
};
}  // namespace

void AmrForest::save(const std::string& path) const {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw Error("AmrForest::save: cannot open " + path);
    try {
        Writer w{f};
        w.arr(kForestMagic, 8);
        w.arr(dir.data(), 3);
        w.pod<uint64_t>(V.size());
        for (const Vec3& v : V) w.arr(v.data(), 3);
        w.pod<uint64_t>(nodes.size());
        for (const Node& n : nodes) { w.arr(n.v, 8); const int32_t q[4] = {n.level, n.parent, n.child0, n.nchild}; w.arr(q, 4); }
        w.pod<uint64_t>(leaves.size());
        for (int l : leaves) w.pod<int32_t>(l);
        w.pod<uint64_t>(edgeMid.size());
        for (const auto& kv : edgeMid) { w.arr(kv.first.data(), 2); w.pod<u32>(kv.second); }
        w.pod<uint64_t>(faceMid.size());
        for (const auto& kv : faceMid) { w.arr(kv.first.data(), 4); w.pod<u32>(kv.second); }
        w.pod<uint64_t>(patchOf.size());
        for (const auto& kv : patchOf) { w.arr(kv.first.data(), 4); w.pod<u32>((u32)kv.second.size()); w.arr(kv.second.data(), kv.second.size()); }
    } catch (...) { std::fclose(f); throw; }
    std::fclose(f);
}

void AmrForest::load(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error("AmrForest::load: cannot open " + path);
    try {
        Reader r{f};
        char magic[8];
        r.arr(magic, 8);
        if (std::memcmp(magic, kForestMagic, 8) != 0) throw Error("AmrForest::load: " + path + " is not a forest file");
        r.arr(dir.data(), 3);
        V.resize(r.pod<uint64_t>());
        for (Vec3& v : V) r.arr(v.data(), 3);
        nodes.resize(r.pod<uint64_t>());
        for (Node& n : nodes) { r.arr(n.v, 8); int32_t q[4]; r.arr(q, 4); n.level = q[0]; n.parent = q[1]; n.child0 = q[2]; n.nchild = q[3]; }
        leaves.resize(r.pod<uint64_t>());
        for (int& l : leaves) { l = r.pod<int32_t>(); if (l < 0 || (size_t)l >= nodes.size()) throw Error("AmrForest::load: leaf out of range"); }
        edgeMid.clear(); faceMid.clear(); patchOf.clear();
        for (uint64_t n = r.pod<uint64_t>(); n > 0; n--) { Key2 k; r.arr(k.data(), 2); edgeMid[k] = r.pod<u32>(); }
        for (uint64_t n = r.pod<uint64_t>(); n > 0; n--) { Key4 k; r.arr(k.data(), 4); faceMid[k] = r.pod<u32>(); }
        for (uint64_t n = r.pod<uint64_t>(); n > 0; n--) {
            Key4 k; r.arr(k.data(), 4);
            std::string name(r.pod<u32>(), ' ');
            r.arr(&name[0], name.size());
            patchOf[k] = name;
        }
    } catch (...) { std::fclose(f); throw; }
    std::fclose(f);
}

// ---------------------------------------------------------------------------------------------------------
// tagging (Prepare::calcQOI + the first half of Prepare::refineMesh, field.cpp:606-620, 696-824)
// ---------------------------------------------------------------------------------------------------------
static bool pair_cyclic_owners(const EulerSolver& s, std::vector<uint8_t>& refine, std::vector<uint8_t>& coarsen, bool check_only);

void amr_tag_cells(const EulerSolver& s, const RefineParams& rp, const std::vector<int>& levels, const std::vector<std::vector<u32>>& families,
                   std::vector<uint8_t>& refine, std::vector<uint8_t>& coarsen, const std::vector<double>* node_volumes) {
    const u32 nB = s.geo.nBCS;
    const int NP = Basis(s.nop).NP;
    const uint64_t n = (uint64_t)nB * NP;
    if (levels.size() != nB) throw Error("amr_tag_cells: one level per cell expected");
    const std::vector<double>* f = nullptr;
    int comps = 1;
    if (rp.field == "T") f = s.convection ? &s.rho : &s.T;      // the convection app's scalar T lives in the rho slot
    else if (rp.field == "p") f = &s.p; else if (rp.field == "rho") f = &s.rho;
    else if (rp.field == "U") { f = &s.U; comps = 3; }
    else throw Error("refinement{field " + rp.field + "}: rho, U, T or p expected");
    // qoi = sqrt(|f|) * (cV^(1/8) / max), normalised to max 1 (calcQOI, field.cpp:606-620)
    const std::vector<double>& cV = node_volumes ? *node_volumes : s.geo.cV;
    if (cV.size() < n) throw Error("amr_tag_cells: node volumes do not cover the mesh");
    std::vector<double> qoi(n);
    double maxdx = -1e30;
    for (uint64_t i = 0; i < n; i++) maxdx = std::max(maxdx, std::fabs(std::pow(cV[i], 0.125)));
    double maxq = -1e30;
    for (uint64_t i = 0; i < n; i++) {
        double m;
        if (comps == 1) m = std::fabs((*f)[i]);
        else { const double* u = f->data() + i * 3; m = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]); }
        qoi[i] = std::pow(m, 0.5) * (std::pow(cV[i], 0.125) / maxdx);
        maxq = std::max(maxq, std::fabs(qoi[i]));
    }
    std::vector<int> want(nB, 0);
    refine.assign(nB, 0);
    coarsen.assign(nB, 0);
    for (u32 c = 0; c < nB; c++) {
        double q = 0, vol = 0;
        for (int j = 0; j < NP; j++) { const uint64_t i = (uint64_t)c * NP + j; q += (qoi[i] / maxq) * cV[i]; vol += cV[i]; }
        q /= vol;
        if (q >= rp.field_max) {
            int level = 1;
            for (; level < 3; level++) {
                const double delta = ((1 - rp.field_max) / 2) * (2 - 1.0 / (1 << (level - 1)));
                if (q < rp.field_max + delta) break;
            }
            want[c] = level;
        } else if (q <= rp.field_min && levels[c]) coarsen[c] = 1;
    }
    // neighbours across faces (real cells only)
    const MeshTopo& t = s.topo;
    auto for_neighbours = [&](u32 c, const std::function<void(u32)>& fn) {
        for (u32 q = t.cellStart[c]; q < t.cellStart[c + 1]; q++) {
            const u32 fi = t.cellFaces[q];
            const u32 o = (t.FOC[fi] == c) ? t.FNC[fi] : t.FOC[fi];
            if (o < nB) fn(o);
        }
    };
    // buffer zone (field.cpp:728-756)
    for (int b = 0; b < rp.buffer_zone; b++)
        for (int level = 1; level <= rp.max_level; level++) {
            std::vector<uint8_t> buf(nB, 0);
            for (u32 c = 0; c < nB; c++) {
                if (want[c] != level) continue;
                for_neighbours(c, [&](u32 o) { if (want[o] < want[c]) buf[o] = 1; });
            }
            for (u32 c = 0; c < nB; c++) if (buf[c]) want[c]++;
        }
    for (u32 c = 0; c < nB; c++)
        if (want[c] && (long)nB <= rp.limit && levels[c] < rp.max_level && levels[c] < want[c]) { refine[c] = 1; coarsen[c] = 0; }
    // whole families only (field.cpp:773-787)
    std::vector<uint8_t> inFamily(nB, 0);
    for (const auto& fam : families) {
        bool all = true;
        for (u32 c : fam) all &= (coarsen[c] != 0);
        for (u32 c : fam) { inFamily[c] = 1; if (!all) coarsen[c] = 0; }
    }
    for (u32 c = 0; c < nB; c++) if (!inFamily[c]) coarsen[c] = 0;
    // 2:1 balance across faces (field.cpp:790-824), repeated until nothing changes; a family that loses a member loses all
    bool changed = true;
    while (changed) {
        changed = false;
        for (u32 c = 0; c < nB; c++) {
            if (!(refine[c] || coarsen[c])) continue;
            bool drop = false;
            for_neighbours(c, [&](u32 o) {
                const int lo = levels[o] - coarsen[o] + refine[o];
                if (refine[c]) drop |= (levels[c] > lo);
                else drop |= (levels[c] < lo);
            });
            if (drop) { refine[c] = 0; coarsen[c] = 0; changed = true; }
        }
        for (const auto& fam : families) {
            bool all = true;
            for (u32 c : fam) all &= (coarsen[c] != 0);
            if (!all) for (u32 c : fam) if (coarsen[c]) { coarsen[c] = 0; changed = true; }
        }
    }
    // owner cells of paired CYCLIC faces go together (field.cpp:826-858), then whole families again (field.cpp:860-875)
    pair_cyclic_owners(s, refine, coarsen, false);
    for (const auto& fam : families) {
        bool all = true;
        for (u32 c : fam) all &= (coarsen[c] != 0);
        if (!all) for (u32 c : fam) coarsen[c] = 0;
    }
}

// (patch, neighbor) of every CYCLIC condition the field files state, each pair once (field.cpp:828-838)
static std::vector<std::array<std::string, 2>> cyclic_pairs(const EulerSolver& s) {
    std::vector<std::array<std::string, 2>> out;
    for (const std::vector<BCond>* l : {&s.file_bc_rho, &s.file_bc_U, &s.file_bc_T, &s.file_bc_p})
        for (const BCond& b : *l) {
            if (b.type != "CYCLIC") continue;
            bool seen = false;
            for (const auto& pr : out) seen |= (pr[0] == b.patch || pr[1] == b.patch);
            if (!seen) out.push_back({b.patch, b.neighbor});
        }
    return out;
}

// the owner cells of paired CYCLIC faces are refined together and coarsened together (field.cpp:826-858): face j of a patch is paired with
// face j of its neighbor patch.  check_only: report a violation instead of repairing it (explicit flags)
static bool pair_cyclic_owners(const EulerSolver& s, std::vector<uint8_t>& refine, std::vector<uint8_t>& coarsen, bool check_only) {
    bool consistent = true;
    for (const auto& pr : cyclic_pairs(s)) {
        auto i1 = s.topo.boundaries.find(pr[0]), i2 = s.topo.boundaries.find(pr[1]);
        if (i1 == s.topo.boundaries.end() || i2 == s.topo.boundaries.end() || i1->second.size() != i2->second.size())
            throw Error("CYCLIC patches " + pr[0] + "/" + pr[1] + " missing or of different size");
        for (size_t j = 0; j < i1->second.size(); j++) {
            const u32 c1 = s.topo.FOC[i1->second[j]], c2 = s.topo.FOC[i2->second[j]];
            if (check_only) { consistent &= (refine[c1] == refine[c2]) && (coarsen[c1] == coarsen[c2]); continue; }
            if (refine[c1] && !refine[c2]) { refine[c2] = 1; coarsen[c2] = 0; }
            else if (refine[c2] && !refine[c1]) { refine[c1] = 1; coarsen[c1] = 0; }
            else if (refine[c1] && refine[c2]) {}
            else if (!coarsen[c1] || !coarsen[c2]) coarsen[c1] = coarsen[c2] = 0;
        }
    }
    return consistent;
}

// After a regrid the faces of a CYCLIC patch and of its neighbor patch are paired by position again (field.h:2662-2664 pairs face j with
// face j): the neighbor's list is reordered so that face j lies opposite face j of the patch (same bounding-box centre up to the translation
// between the two patches).  Throws when the two sides were not refined alike.
static void repair_cyclic_order(Grid& g, const std::vector<std::array<std::string, 2>>& pairs) {
    auto centre = [&](u32 f) {
        Vec3 lo{1e300, 1e300, 1e300}, hi{-1e300, -1e300, -1e300};
        for (u32 k = g.facetStart[f]; k < g.facetStart[f + 1]; k++)
            for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], g.V[g.facetVerts[k]][d]); hi[d] = std::max(hi[d], g.V[g.facetVerts[k]][d]); }
        return Vec3{(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2};
    };
    for (const auto& pr : pairs) {
        auto i1 = g.boundaries.find(pr[0]), i2 = g.boundaries.find(pr[1]);
        if (i1 == g.boundaries.end() || i2 == g.boundaries.end() || i1->second.size() != i2->second.size())
            throw Error("regrid: CYCLIC patches " + pr[0] + "/" + pr[1] + " were not refined alike (paired faces must be refined together)");
        const std::vector<u32>&A = i1->second;
        std::vector<u32>& B = i2->second;
        if (A.empty()) continue;
        std::vector<Vec3> ca(A.size()), cb(B.size());
        Vec3 ma{0, 0, 0}, mb{0, 0, 0}, lo{1e300, 1e300, 1e300}, hi{-1e300, -1e300, -1e300};
        for (size_t j = 0; j < A.size(); j++) {
            ca[j] = centre(A[j]); cb[j] = centre(B[j]);
            for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], ca[j][d]); hi[d] = std::max(hi[d], ca[j][d]); }
        }
        // translation between the patches: difference of the patches' own bounding-box centres (independent of how the faces are listed)
        Vec3 la{1e300, 1e300, 1e300}, ha{-1e300, -1e300, -1e300}, lb = la, hb = ha;
        for (size_t j = 0; j < A.size(); j++)
            for (int d = 0; d < 3; d++) {
                la[d] = std::min(la[d], ca[j][d]); ha[d] = std::max(ha[d], ca[j][d]);
                lb[d] = std::min(lb[d], cb[j][d]); hb[d] = std::max(hb[d], cb[j][d]);
            }
        double scale = 0;
        for (int d = 0; d < 3; d++) { ma[d] = (la[d] + ha[d]) / 2; mb[d] = (lb[d] + hb[d]) / 2; scale = std::max(scale, std::max(ha[d] - la[d], std::fabs(mb[d] - ma[d]))); }
        const double tol = 1e-9 * (scale > 0 ? scale : 1.0);
        std::vector<u32> nb(B.size(), MAX_INT);
        std::vector<uint8_t> used(B.size(), 0);
        for (size_t j = 0; j < A.size(); j++) {
            size_t best = B.size();
            // same position first (the usual case: both lists come out of the forest in the same order), then a search
            auto close = [&](size_t k) {
                double e = 0;
                for (int d = 0; d < 3; d++) e = std::max(e, std::fabs((cb[k][d] - mb[d]) - (ca[j][d] - ma[d])));
                return e <= tol;
            };
            if (!used[j] && close(j)) best = j;
            else for (size_t k = 0; k < B.size() && best == B.size(); k++) if (!used[k] && close(k)) best = k;
            if (best == B.size())
                throw Error("regrid: CYCLIC patches " + pr[0] + "/" + pr[1] + " were not refined alike (paired faces must be refined together)");
            used[best] = 1;
            nb[j] = B[best];
        }
        B = nb;
    }
}

// ---------------------------------------------------------------------------------------------------------
// the solver on the regridded mesh
// ---------------------------------------------------------------------------------------------------------
void EulerSolver::copy_run_parameters(EulerSolver& n) const {
    n.ctl = ctl; n.dir = dir; n.meshName = meshName;
    for (int d = 0; d < 3; d++) { n.nop[d] = nop[d]; n.decomp_n[d] = decomp_n[d]; }
    n.viscosity = viscosity; n.Pr = Pr; n.T0 = T0; n.P0 = P0; n.cp = cp; n.cv = cv; n.dt = dt; n.gravity = gravity;
    n.buoyancy = buoyancy; n.diffusion = diffusion; n.binary_out = binary_out;
    n.time_scheme = time_scheme; n.problem_init = "NONE";
    n.start_step = start_step; n.end_step = end_step; n.write_interval = write_interval; n.decomp_type = decomp_type;
    n.refine_params = refine_params; n.amr_step = amr_step;
    n.mass0 = mass0; n.energy0 = energy0; n.volume0 = volume0;
    n.vtk_fields = vtk_fields; n.vtk_cell_value = vtk_cell_value; n.vtk_polyhedral = vtk_polyhedral; n.vtk_on_dump = vtk_on_dump;
    n.launch_nonce = launch_nonce;
    n.conv_scheme = conv_scheme; n.blend_factor = blend_factor; n.cyclic_patches = cyclic_patches;
    n.convection = convection; n.conv_init = conv_init; n.scalar0 = scalar0; n.conv_end_step = conv_end_step;
    n.topo.spherical = topo.spherical; n.topo.sphere_radius = topo.sphere_radius; n.topo.sphere_height = topo.sphere_height;
}

std::unique_ptr<EulerSolver> EulerSolver::regridded(const std::vector<uint8_t>& refine, const std::vector<uint8_t>& coarsen) {
    if (nranks > 1) throw Error("EulerSolver::regridded: the in-memory regrid runs on one partition (repartitioning a regridded mesh is not built)");
    for (const std::vector<BCond>* l : {&bc_rho, &bc_U, &bc_T, &bc_p})
        for (const BCond& b : *l)
            if (b.held)
                throw Error("EulerSolver::regridded: patch " + b.patch + " has no boundary condition for one of the fields; the values its boundary "
                            "cells keep cannot follow an in-memory regrid");
    for (const std::vector<BCond>* l : {&file_bc_rho, &file_bc_U, &file_bc_T, &file_bc_p})
        for (const BCond& b : *l)
        {
            if (!b.fixed.empty()) throw Error("EulerSolver::regridded: boundary conditions with frozen per-face values cannot follow a regrid");
            // CALC_DIRICHLET freezes the values the patch holds when the condition is first applied; the set-up below runs on blank fields
            // (the state arrives afterwards, on the device), so such a patch would freeze zeros -- the reference re-reads the refined
            // fields and freezes those.  Refused rather than silently wrong (the shipped AMR cases use NEUMANN / SYMMETRY / CYCLIC)
            if (b.type == "CALC_DIRICHLET")
                throw Error("EulerSolver::regridded: a CALC_DIRICHLET patch (" + b.patch + ") cannot follow an in-memory regrid: its frozen values "
                            "would be taken from the blank set-up fields");
        }
    // the reference refines the owner cells of paired CYCLIC faces together (field.cpp:826-858; amr_tag_cells does) and pairs the faces by
    // their position in the two patches (field.h:2662-2664; repair_cyclic_order below).  Flags that split a pair are refused before the
    // forest is touched
    const std::vector<std::array<std::string, 2>> cyc = cyclic_pairs(*this);
    if (!cyc.empty()) {
        std::vector<uint8_t> r = refine, c = coarsen;
        if (r.size() != topo.nBCS || c.size() != topo.nBCS) throw Error("EulerSolver::regridded: one refine and one coarsen flag per cell expected");
        if (!pair_cyclic_owners(*this, r, c, true))
            throw Error("EulerSolver::regridded: the owner cells of paired CYCLIC faces must be refined (or coarsened) together");
    }
    if (!forest) throw Error("EulerSolver::regridded: no AMR forest (the mesh must come from set_mesh/load_mesh of a conforming hexahedral grid)");
    std::unique_ptr<EulerSolver> n(new EulerSolver());
    copy_run_parameters(*n);
    n->forest = forest;
    const bool verbose = std::getenv("NSEM_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        const auto t1 = std::chrono::steady_clock::now();
        if (verbose) std::printf("regrid: %-28s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    n->last_maps = forest->regrid(refine, coarsen);
    lap("forest regrid");
    Grid g = forest->grid();
    repair_cyclic_order(g, cyc);
    lap("grid emission");
    if (keep_regrid_grid) { n->regrid_grid = std::make_shared<Grid>(g); n->keep_regrid_grid = true; }
    n->topo.load(g);
    lap("topology (MeshTopo::load)");
    Basis b(nop);
    n->geo.build(n->topo, b);
    lap("node geometry");
    auto blank = [](int comps, const std::vector<BCond>& bcs) {
        FieldFile f;
        f.comps = comps;
        f.inits.push_back({"uniform", std::vector<double>(comps, 0.0)});
        f.bcs = bcs;
        return f;
    };
    n->set_fields(blank(1, file_bc_rho), blank(3, file_bc_U), blank(1, file_bc_T), blank(1, file_bc_p));
    lap("blank fields + BC tables");
    n->setup();                                         // reference state, gravity, BC tables on the new mesh (the fields are overwritten below)
    n->mass0 = mass0; n->energy0 = energy0; n->volume0 = volume0; n->scalar0 = scalar0;
    lap("set-up (reference state)");
    if (ctx) {
        n->attach_device(device_id);
        lap("attach (mesh upload)");
        n->adopt_refined_state(*this, n->last_maps.refineMap, n->last_maps.coarseMap, n->last_maps.cellMap, true);
        lap("device transfer + restart");
    }
    return n;
}

void EulerSolver::write_amr_grid(long dump) const {
    if (!forest) throw Error("EulerSolver::write_amr_grid: no AMR forest");
    const std::string base = dir + "/" + meshName + "_" + std::to_string(dump);
    // the reference writes the refined grid in the dump format (Mesh::write_mesh under write_format, field.cpp:74,930)
    if (binary_out) write_grid_binary(base + ".bin", forest->grid());
    else write_grid_text(base + ".txt", forest->grid());
    forest->save(base + ".forest");
}

std::unique_ptr<EulerSolver> EulerSolver::regridded_by_indicator() {
    if (!forest) throw Error("EulerSolver::regridded_by_indicator: no AMR forest");
    if (ctx) download();
    std::vector<uint8_t> refine, coarsen;
    if (topo.spherical) {
        // Prepare::refineMesh loads a spherical mesh twice (field.cpp:638-645): projected, for the volumes the field transfer uses, and
        // then as the grid file has it -- the cube shell, still with the spherical corrections of calcGeometry and of the node placement --
        // and it is THAT load's node volumes calcQOI weighs the indicator with (field.cpp:606-620)
        MeshTopo flat;
        flat.spherical = true; flat.sphere_radius = topo.sphere_radius; flat.sphere_height = topo.sphere_height; flat.no_extrude = true;
        flat.load(forest->grid());
        if (flat.nBCS != topo.nBCS) throw Error("regridded_by_indicator: the forest's grid and the solver's mesh differ");
        Geometry fg;
        fg.build(flat, Basis(nop));
        amr_tag_cells(*this, refine_params, forest->levels(), forest->families(), refine, coarsen, &fg.cV);
    } else {
        amr_tag_cells(*this, refine_params, forest->levels(), forest->families(), refine, coarsen, nullptr);
    }
    return regridded(refine, coarsen);
}

}  // namespace nsemh
