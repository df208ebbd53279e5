// partition.cpp -- domain decomposition: one mesh partition per GPU.
//
// Mirrors Prepare::decomposeMesh (src/field/field.cpp:1086-1257) in memory instead of through grid<r>/ files:
//   * cell -> part by METIS k-way on the element graph (field.cpp:1010-1080: ncon 1, objective CUT, IPTYPE EDGE,
//     UFACTOR 30), by bounding-box slabs (decomposeXYZ, :976-1006) or by cell index (decomposeIndex, :967-972);
//   * per part: vertices, facets and cells keep their ascending global order; physical patches are filtered; every
//     cut face goes into the patch `interMesh_<me>_<peer>` in ascending global face order, so both sides list the
//     shared faces in the same order (what ASYNC_COMM relies on, field.h:2267-2323).
// METIS 5 comes from the static library bundled with the CUDA toolkit (64-bit idx_t, 32-bit real_t); the option
// indices are those of METIS 5.1.0's metis.h.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "nsem_host.h"

#ifdef NSEM_WITH_METIS
extern "C" {
int METIS_SetDefaultOptions(int64_t* options);
int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                        int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* edgecut,
                        int64_t* part);
}
#endif

namespace nsemh {

static void face_cells(const Grid& g, std::vector<u32>& foc, std::vector<u32>& fnc) {
    foc.assign(g.nFacets(), MAX_INT);
    fnc.assign(g.nFacets(), MAX_INT);
    for (u32 c = 0; c < g.nCells(); c++)
        for (u32 q = g.cellStart[c]; q < g.cellStart[c + 1]; q++) {
            const u32 f = g.cellFaces[q];
            if (foc[f] == MAX_INT) foc[f] = c;
            else fnc[f] = c;
        }
}

static std::vector<u32> partition_cells_raw(const Grid& g, int nparts, const std::string& method, const int nxyz[3],
                                            const std::vector<u32>* cluster_root);

// Cells joined by non-conforming faces form clusters that must stay in one part (both sides of a mortar are evaluated by one rank,
// field.cpp:1215-1220): root[c] = the smallest cell index of c's cluster.
static std::vector<u32> mortar_clusters(const Grid& g, const std::vector<u32>* face_mortar, const std::vector<std::array<u32, 2>>* together) {
    std::vector<u32> foc, fnc, root(g.nCells());
    face_cells(g, foc, fnc);
    for (u32 c = 0; c < g.nCells(); c++) root[c] = c;
    auto find = [&](u32 c) { while (root[c] != c) { root[c] = root[root[c]]; c = root[c]; } return c; };
    auto join = [&](u32 x, u32 y) {
        const u32 a = find(x), b = find(y);
        if (a != b) root[std::max(a, b)] = std::min(a, b);
    };
    if (face_mortar)
        for (u32 f = 0; f < g.nFacets(); f++)
            if ((*face_mortar)[f] != 0 && fnc[f] != MAX_INT) join(foc[f], fnc[f]);
    // pairs of FACES whose owner cells must share a part: the two sides of a CYCLIC pair of patches (the ghost value of one is the owner
    // value at the other, applyExplicitBCs field.h:2683-2687 -- a local read)
    if (together)
        for (const auto& pr : *together)
            if (pr[0] < g.nFacets() && pr[1] < g.nFacets() && foc[pr[0]] != MAX_INT && foc[pr[1]] != MAX_INT) join(foc[pr[0]], foc[pr[1]]);
    for (u32 c = 0; c < g.nCells(); c++) root[c] = find(c);
    return root;
}

std::vector<u32> partition_cells(const Grid& g, int nparts, const std::string& method, const int nxyz[3],
                                 const std::vector<u32>* face_mortar, const std::vector<std::array<u32, 2>>* together) {
    std::vector<u32> part;
    if (together && together->empty()) together = nullptr;
    if ((face_mortar || together) && nparts > 1) {
        // the reference weighs mortar faces 1000 in the METIS graph (field.cpp:1037-1047), which makes cutting one unlikely but not
        // impossible (a tight balance on a few hundred cells does it).  Here the clusters are contracted first: METIS sees one vertex of
        // weight |cluster| per cluster (the edge weights count the faces between clusters), the other methods place a cluster where its
        // first cell goes -- a non-conforming face cannot be cut
        const std::vector<u32> root = mortar_clusters(g, face_mortar, together);
        if (method == "METIS") part = partition_cells_raw(g, nparts, method, nxyz, &root);
        else {
            part = partition_cells_raw(g, nparts, method, nxyz, nullptr);
            for (u32 c = 0; c < g.nCells(); c++) part[c] = part[root[c]];
        }
    } else {
        part = partition_cells_raw(g, nparts, method, nxyz, nullptr);
    }
    if (face_mortar && nparts > 1) {
        // a non-conforming face must not be cut: both sides of a mortar are evaluated by one rank (field.cpp:1215-1220)
        std::vector<u32> foc, fnc;
        face_cells(g, foc, fnc);
        for (u32 f = 0; f < g.nFacets(); f++)
            if ((*face_mortar)[f] != 0 && fnc[f] != MAX_INT && part[foc[f]] != part[fnc[f]])
                throw Error("decomposition cuts the non-conforming face " + std::to_string(f) + ": use METIS (edge weight 1000 on mortar faces) or fewer parts");
    }
    return part;
}

static std::vector<u32> partition_cells_raw(const Grid& g, int nparts, const std::string& method, const int nxyz[3],
                                            const std::vector<u32>* cluster_root) {
    const u32 nc = g.nCells();
    std::vector<u32> part(nc, 0);
    if (nparts <= 1) return part;
    if (method == "CELLID") {
        const u32 per = std::max<u32>(1, nc / nparts);
        for (u32 i = 0; i < nc; i++) part[i] = std::min<u32>(i / per, nparts - 1);
        return part;
    }
    if (method == "XYZ") {
        if (nxyz[0] * nxyz[1] * nxyz[2] != nparts) throw Error("Error in XYZ decomposition: use " + std::to_string(nxyz[0] * nxyz[1] * nxyz[2]) + " ranks");
        Vec3 lo{1e300, 1e300, 1e300}, hi{-1e300, -1e300, -1e300};
        for (const auto& v : g.V)
            for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], v[d]); hi[d] = std::max(hi[d], v[d]); }
        for (u32 c = 0; c < nc; c++) {
            // centroid of the cell's face-vertex cloud
            Vec3 C{0, 0, 0};
            double cnt = 0;
            for (u32 q = g.cellStart[c]; q < g.cellStart[c + 1]; q++) {
                const u32 f = g.cellFaces[q];
                for (u32 r = g.facetStart[f]; r < g.facetStart[f + 1]; r++) {
                    for (int d = 0; d < 3; d++) C[d] += g.V[g.facetVerts[r]][d];
                    cnt += 1;
                }
            }
            u32 id[3];
            for (int d = 0; d < 3; d++) {
                const double x = (C[d] / cnt - lo[d]) / ((hi[d] - lo[d]) / nxyz[d]);
                id[d] = (u32)std::min<double>(std::max(0.0, std::floor(x)), nxyz[d] - 1);
            }
            part[c] = (id[0] * nxyz[1] + id[1]) * nxyz[2] + id[2];
        }
        return part;
    }
    if (method == "METIS") {
#ifdef NSEM_WITH_METIS
        std::vector<u32> foc, fnc;
        face_cells(g, foc, fnc);
        // vertices of the graph: cells, or clusters of cells joined by non-conforming faces (cluster_root)
        std::vector<u32> vid(nc);
        std::vector<int64_t> vw;
        if (cluster_root) {
            std::vector<u32> id(nc, MAX_INT);
            for (u32 c = 0; c < nc; c++) {
                const u32 r = (*cluster_root)[c];
                if (id[r] == MAX_INT) { id[r] = (u32)vw.size(); vw.push_back(0); }
                vid[c] = id[r];
                vw[vid[c]]++;
            }
        } else {
            for (u32 c = 0; c < nc; c++) vid[c] = c;
        }
        const u32 nvx = cluster_root ? (u32)vw.size() : nc;
        if ((int)nvx < nparts) throw Error("decomposition: " + std::to_string(nvx) + " clusters of cells joined by non-conforming faces cannot fill " + std::to_string(nparts) + " parts");
        // edges between different vertices, parallel faces merged into one weighted edge
        std::vector<std::pair<uint64_t, int64_t>> ed;
        if (cluster_root) {
            ed.reserve((size_t)g.nFacets() * 2);
            for (u32 f = 0; f < g.nFacets(); f++)
                if (fnc[f] != MAX_INT && vid[foc[f]] != vid[fnc[f]]) {
                    const u32 a = vid[foc[f]], b = vid[fnc[f]];
                    ed.push_back({((uint64_t)a << 32) | b, 1});
                    ed.push_back({((uint64_t)b << 32) | a, 1});
                }
            std::sort(ed.begin(), ed.end());
        }
        std::vector<int64_t> deg(nvx + 1, 0), adj, wgt;
        bool weighted = false;
        if (!cluster_root) {
            // conforming grid: the element graph with the neighbours in face order (one edge per face, field.cpp:1010-1047)
            for (u32 f = 0; f < g.nFacets(); f++)
                if (fnc[f] != MAX_INT) { deg[foc[f] + 1]++; deg[fnc[f] + 1]++; }
            for (u32 c = 0; c < nc; c++) deg[c + 1] += deg[c];
            adj.assign(deg[nc], 0); wgt.assign(deg[nc], 1);
            std::vector<int64_t> fill(deg.begin(), deg.end() - 1);
            for (u32 f = 0; f < g.nFacets(); f++)
                if (fnc[f] != MAX_INT) { adj[fill[foc[f]]++] = fnc[f]; adj[fill[fnc[f]]++] = foc[f]; }
        }
        for (size_t i = 0; i < ed.size();) {
            size_t j = i;
            int64_t w = 0;
            while (j < ed.size() && ed[j].first == ed[i].first) { w += ed[j].second; j++; }
            adj.push_back((int64_t)(ed[i].first & 0xffffffffu));
            wgt.push_back(w);
            weighted = weighted || w != 1;
            deg[(ed[i].first >> 32) + 1]++;
            i = j;
        }
        if (cluster_root) for (u32 v = 0; v < nvx; v++) deg[v + 1] += deg[v];
        int64_t opt[40];
        METIS_SetDefaultOptions(opt);
        opt[1] = 0;      // METIS_OPTION_OBJTYPE = METIS_OBJTYPE_CUT
        opt[3] = 2;      // METIS_OPTION_IPTYPE  = METIS_IPTYPE_EDGE
        opt[6] = 200;    // METIS_OPTION_NITER
        opt[7] = nc > 200000 ? 1 : 100;   // METIS_OPTION_NCUTS (the reference's 100 cuts are unaffordable on 10^6 cells)
        // METIS_OPTION_UFACTOR: the reference allows 3 % imbalance (30, field.cpp:1061); on one GPU per part the step time is the LARGEST
        // part's, so every per cent of imbalance is a per cent of scaling efficiency (measured: 2.9 % on 8 parts of a box).  Default 1
        // (0.1 %); NSEM_METIS_UFACTOR=30 restores the reference's value.  The partition is not part of the arithmetic: any valid k-way
        // partition gives the same answer bit for bit (tests/mp_gpu_check.py).
        { const char* uf = std::getenv("NSEM_METIS_UFACTOR"); opt[16] = uf ? std::max(1, std::atoi(uf)) : 1; }
        opt[17] = 0;     // METIS_OPTION_NUMBERING: C style
        int64_t nv = nvx, ncon = 1, np = nparts, cut = 0;
        std::vector<int64_t> p64(nvx);
        const int rc = METIS_PartGraphKway(&nv, &ncon, deg.data(), adj.data(), cluster_root ? vw.data() : nullptr, nullptr, weighted ? wgt.data() : nullptr, &np,
                                           nullptr, nullptr, opt, &cut, p64.data());
        if (rc != 1) throw Error("METIS_PartGraphKway failed with code " + std::to_string(rc));
        for (u32 c = 0; c < nc; c++) part[c] = (u32)p64[vid[c]];
        return part;
#else
        throw Error("this build has no METIS (libmetis_static.a of the CUDA toolkit was not found)");
#endif
    }
    throw Error("unknown decomposition type " + method + " (METIS, XYZ, CELLID)");
}

Partition extract_partition(const Grid& g, const std::vector<u32>& part, int rank, int nparts) {
    Partition P;
    std::vector<u32> foc, fnc;
    face_cells(g, foc, fnc);
    const u32 nf = g.nFacets(), nv = (u32)g.V.size();
    std::vector<u32> vloc(nv, MAX_INT), floc(nf, MAX_INT);
    std::vector<char> vuse(nv, 0), fuse(nf, 0);
    for (u32 c = 0; c < g.nCells(); c++) {
        if ((int)part[c] != rank) continue;
        P.cellGlobal.push_back(c);
        for (u32 q = g.cellStart[c]; q < g.cellStart[c + 1]; q++) {
            const u32 f = g.cellFaces[q];
            fuse[f] = 1;
            for (u32 r = g.facetStart[f]; r < g.facetStart[f + 1]; r++) vuse[g.facetVerts[r]] = 1;
        }
    }
    u32 cnt = 0;
    for (u32 v = 0; v < nv; v++) if (vuse[v]) { vloc[v] = cnt++; P.grid.V.push_back(g.V[v]); }
    cnt = 0;
    for (u32 f = 0; f < nf; f++)
        if (fuse[f]) {
            floc[f] = cnt++;
            for (u32 r = g.facetStart[f]; r < g.facetStart[f + 1]; r++) P.grid.facetVerts.push_back(vloc[g.facetVerts[r]]);
            P.grid.facetStart.push_back((u32)P.grid.facetVerts.size());
        }
    for (u32 c : P.cellGlobal) {
        for (u32 q = g.cellStart[c]; q < g.cellStart[c + 1]; q++) P.grid.cellFaces.push_back(floc[g.cellFaces[q]]);
        P.grid.cellStart.push_back((u32)P.grid.cellFaces.size());
    }
    for (const auto& kv : g.boundaries) {
        std::vector<u32> b;
        for (u32 f : kv.second) if (floc[f] != MAX_INT) b.push_back(floc[f]);
        if (!b.empty()) P.grid.boundaries[kv.first] = b;
    }
    std::vector<std::vector<u32>> cut(nparts);
    for (u32 f = 0; f < nf; f++) {
        if (fnc[f] == MAX_INT) continue;
        const int co = (int)part[foc[f]], cn = (int)part[fnc[f]];
        if (co == cn) continue;
        if (co == rank) cut[cn].push_back(floc[f]);
        else if (cn == rank) cut[co].push_back(floc[f]);
    }
    for (int p = 0; p < nparts; p++)
        if (!cut[p].empty()) {
            P.grid.boundaries["interMesh_" + std::to_string(rank) + "_" + std::to_string(p)] = cut[p];
            P.peers.push_back(p);
        }
    return P;
}

}  // namespace nsemh
