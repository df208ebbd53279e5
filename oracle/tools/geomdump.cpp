/* geomdump -- TEST INFRASTRUCTURE ONLY.
 *
 * Links against the UNMODIFIED NebulaSEM reference objects (compiled from
 * /root/reference by oracle/build_ref.sh) and dumps the geometry/connectivity
 * arrays that Mesh::LoadMesh (src/field/field.cpp:95-167) builds, so that the
 * repo's own mesh/DG-geometry code (host C++ and the numpy restatement) can be
 * compared array by array with the reference:
 *   cC cV (field.h:1355-1367), fN fC fI FO FN, faceIndices/allFaces
 *   (field.cpp:178-193), gFOC gFNC gFMC gFaceID (mesh.h:84-106), DG::Jinv,
 *   dpsi/xgl/wgl/psiRef/psiCor (dg.cpp:481-719), boundary patches.
 *
 * usage: geomdump ./controls out.bin      (run inside a prepared case dir)
 *
 * File format: sequence of records  [u32 namelen][name][u32 dtype(0=f64,1=u32)]
 * [u64 count][payload].
 */
#include "field.h"
#include "mp.h"
#include "wrapper.h"
#include <cstdio>
#include <cstdint>

static FILE* out;
static void rec(const char* name, uint32_t dtype, uint64_t count, const void* data) {
    uint32_t n = (uint32_t)strlen(name);
    fwrite(&n, 4, 1, out);
    fwrite(name, 1, n, out);
    fwrite(&dtype, 4, 1, out);
    fwrite(&count, 8, 1, out);
    fwrite(data, dtype == 0 ? 8 : 4, count, out);
}
static void rec_u32(const char* name, const std::vector<Int>& v) {
    rec(name, 1, v.size(), v.empty() ? nullptr : &v[0]);
}

int main(int argc, char* argv[]) {
    MP mp(argc, argv);
    /* Solver::Initialize insists that argv[0] contains the solver name; pass "euler" etc. through argv[0] games */
    std::string solver = "euler";
    if (argc > 3) solver = argv[3];
    std::string fake0 = std::string("./") + solver;
    char* av[2] = {const_cast<char*>(fake0.c_str()), argv[1]};
    Solver::Initialize(2, av);
    Util::read_params(Solver::input, false);
    Mesh::LoadMesh(0);

    using namespace Mesh;
    using namespace DG;
    out = fopen(argv[2], "wb");
    uint32_t dims[16] = {NPX, NPY, NPZ, NP, NPF, gBCS, gNCells, gNFacets, gNVertices,
                         gBCSfield, gALLfield, (uint32_t)cC.size(), (uint32_t)fN.size(), 0, 0, 0};
    rec("dims", 1, 16, dims);
    rec("cC", 0, (uint64_t)gALLfield * 3, &cC[0]);
    rec("cV", 0, (uint64_t)gALLfield, &cV[0]);
    rec("Jinv", 0, (uint64_t)gBCSfield * 9, &Jinv[0]);
    rec("fN", 0, (uint64_t)gNFacets * NPF * 3, &fN[0]);
    rec("fC", 0, (uint64_t)gNFacets * NPF * 3, &fC[0]);
    rec("fI", 0, (uint64_t)gNFacets * NPF, &fI[0]);
    rec("FO", 1, (uint64_t)gNFacets * NPF, &FO[0]);
    rec("FN", 1, (uint64_t)gNFacets * NPF, &FN[0]);
    rec("faceIndices0", 1, gNCells, faceIndices[0]);
    rec("faceIndices1", 1, gNCells, faceIndices[1]);
    rec("allFaces", 1, faceIndices[1][gNCells - 1], allFaces);
    rec_u32("gFOC", gFOC);
    rec_u32("gFNC", gFNC);
    rec_u32("gFMC", gFMC);
    {
        std::vector<Int> flat;
        forEach(gFaceID, i) forEach(gFaceID[i], j) flat.push_back(gFaceID[i][j]);
        rec_u32("faceID", flat);
    }
    rec("gFC", 0, (uint64_t)gFC.size() * 3, &gFC[0]);
    rec("gFN", 0, (uint64_t)gFN.size() * 3, &gFN[0]);
    rec("gCV", 0, (uint64_t)gCV.size(), &gCV[0]);
    rec("gCC", 0, (uint64_t)gCC.size() * 3, &gCC[0]);
    rec("vertices", 0, (uint64_t)gVertices.size() * 3, &gVertices[0]);
    {
        std::vector<Int> fsz, fv;
        forEach(gFacets, i) { fsz.push_back(gFacets[i].size()); forEach(gFacets[i], j) fv.push_back(gFacets[i][j]); }
        rec_u32("facet_sizes", fsz);
        rec_u32("facet_verts", fv);
    }
    for (int d = 0; d < 3; d++) {
        Int n = (d == 0) ? NPX : (d == 1) ? NPY : NPZ;
        char nm[32];
        snprintf(nm, 32, "xgl%d", d);  rec(nm, 0, n, xgl[d]);
        snprintf(nm, 32, "wgl%d", d);  rec(nm, 0, n, wgl[d]);
        snprintf(nm, 32, "psi%d", d);  rec(nm, 0, n * n, psi[d]);
        snprintf(nm, 32, "dpsi%d", d); rec(nm, 0, n * n, dpsi[d]);
        for (int h = 0; h < 2; h++) {
            snprintf(nm, 32, "psiRef%d", d * 2 + h); rec(nm, 0, n * n, psiRef[d * 2 + h]);
            snprintf(nm, 32, "psiCor%d", d * 2 + h); rec(nm, 0, n * n, psiCor[d * 2 + h]);
        }
    }
    /* boundary patches in map (alphabetical) order */
    forEachIt(gBoundaries, it) {
        std::string nm = std::string("patch:") + it->first;
        rec_u32(nm.c_str(), it->second);
    }
    fclose(out);
    return 0;
}
