/* refinedump -- TEST INFRASTRUCTURE ONLY.
 *
 * Links against the UNMODIFIED NebulaSEM reference objects (oracle/build_ref.sh) and drives ONE regrid of a case
 * directory with an explicit list of cells to refine / coarsen, through the reference's own MeshObject::refineMesh
 * (src/mesh/mesh.cpp:2216-2748) and MeshField::refineField (src/field/field.h:1863-2015), in the call order of
 * Prepare::refineMesh (src/field/field.cpp:625-955) minus its tagging (calcQOI, thresholds, buffer zone, balance).
 * It dumps what refineField was given -- refineMap / coarseMap / cellMap, old and new cell volumes and centroids, the
 * old node coordinates, psiRef / psiCor / wgl -- and the field arrays before the transfer; the transferred fields are
 * what the reference itself writes (<field><step>.bin).  The repo's numpy restatement (oracle/amr.py) and the CUDA
 * transfer (nsem_refine_state) are compared with these.
 *
 * usage: refinedump ./controls out.bin cells.txt [step]     (inside a prepared case dir)
 *   cells.txt: "r <n> i0 i1 ..." cells to split (one level, all AMR directions), "c <n> j0 j1 ..." cells to merge
 *   (a family is merged only when all its siblings are listed and are leaves, like field.cpp:860-875).
 *
 * Record format as geomdump: [u32 namelen][name][u32 dtype(0=f64,1=u32)][u64 count][payload].
 */
#include "field.h"
#include "mp.h"
#include "wrapper.h"
#include "prepare.h"
#include <cstdio>
#include <cstdint>

static FILE* out;
static void rec(const char* name, uint32_t dtype, uint64_t count, const void* data) {
    uint32_t n = (uint32_t)strlen(name);
    fwrite(&n, 4, 1, out);
    fwrite(name, 1, n, out);
    fwrite(&dtype, 4, 1, out);
    fwrite(&count, 8, 1, out);
    if (count) fwrite(data, dtype == 0 ? 8 : 4, count, out);
}
static void rec_u32(const char* name, const std::vector<Int>& v) { rec(name, 1, v.size(), v.empty() ? nullptr : &v[0]); }

int main(int argc, char* argv[]) {
    MP mp(argc, argv);
    if (argc < 4) { fprintf(stderr, "usage: refinedump ./controls out.bin cells.txt [step]\n"); return 2; }
    std::string fake0 = "./euler";
    char* av[2] = {const_cast<char*>(fake0.c_str()), argv[1]};
    Solver::Initialize(2, av);
    Util::read_params(Solver::input, false);
    const Int step = (argc > 4) ? atoi(argv[4]) : 0;

    using namespace Mesh;
    using namespace DG;
    Mesh::amr_direction = Controls::refine_params.dir;
    LoadMesh(step, false, false);
    ScalarVector oldCV = gCV;
    VectorVector oldCC = gCC;
    Prepare::createFields(BaseField::fieldNames, step);
    Prepare::readFields(BaseField::fieldNames, step);
    gCells.erase(gCells.begin() + gBCS, gCells.end());

    /* the AMR tree of the last regrid, or the trivial one */
    {
        std::stringstream path;
        path << "amrTree_" << step << ".bin";
        if (System::exists(path.str())) {
            Util::ifstream_bin is(path.str());
            is >> gAmrTree;
        } else {
            gAmrTree.resize(gCells.size());
            forEach(gCells, i) gAmrTree[i].id = i;
        }
    }

    IntVector cCells, rCellsL, rLevelL, rDirsL;
    cCells.assign(gBCS, 0);
    {
        FILE* f = fopen(argv[3], "r");
        if (!f) { fprintf(stderr, "refinedump: cannot open %s\n", argv[3]); return 2; }
        char kind;
        unsigned n;
        while (fscanf(f, " %c %u", &kind, &n) == 2) {
            for (unsigned k = 0; k < n; k++) {
                unsigned c;
                if (fscanf(f, "%u", &c) != 1 || c >= gBCS) { fprintf(stderr, "refinedump: bad cell list\n"); return 2; }
                if (kind == 'r') { rCellsL.push_back(c); rLevelL.push_back(1); rDirsL.push_back(7); }
                else cCells[c] = 1;
            }
        }
        fclose(f);
    }
    /* only complete families of leaves may merge */
    forEach(gAmrTree, i) {
        Node& n = gAmrTree[i];
        if (!n.nchildren) continue;
        bool all = true;
        for (Int j = 0; j < n.nchildren; j++) {
            Node& cn = gAmrTree[n.cid + j];
            if (cn.nchildren || !cCells[cn.id]) all = false;
        }
        if (!all)
            for (Int j = 0; j < n.nchildren; j++) cCells[gAmrTree[n.cid + j].id] = 0;
    }

    out = fopen(argv[2], "wb");
    uint32_t dims[8] = {NPX, NPY, NPZ, NP, gBCS, gBCSfield, 0, 0};
    rec("cC_old", 0, (uint64_t)gBCSfield * 3, &cC[0]);
    rec("oldCV", 0, oldCV.size(), &oldCV[0]);
    rec("oldCC", 0, (uint64_t)oldCC.size() * 3, &oldCC[0]);
    forEach(BaseField::fieldNames, i) {
        BaseField* bf = BaseField::findField(BaseField::fieldNames[i]);
        if (!bf) continue;
        std::string nm = "pre:" + BaseField::fieldNames[i];
        if (ScalarCellField* s = dynamic_cast<ScalarCellField*>(bf)) rec(nm.c_str(), 0, gBCSfield, &(*s)[0]);
        else if (VectorCellField* v = dynamic_cast<VectorCellField*>(bf)) rec(nm.c_str(), 0, (uint64_t)gBCSfield * 3, &(*v)[0]);
    }

    IntVector refineMap, coarseMap, cellMap;
    gMesh.refineMesh(cCells, rCellsL, rLevelL, rDirsL, refineMap, coarseMap, cellMap);
    const Int nCells = gCells.size();
    gMesh.addBoundaryCells();
    gMesh.fixHexCells();
    gMesh.calcGeometry();
    VectorVector newCC = gCC;

    dims[6] = nCells;
    rec("dims", 1, 8, dims);
    rec_u32("refineMap", refineMap);
    rec_u32("coarseMap", coarseMap);
    rec_u32("cellMap", cellMap);
    rec_u32("cCells", cCells);
    rec_u32("rCells", rCellsL);
    rec("newCC", 0, (uint64_t)newCC.size() * 3, &newCC[0]);
    rec("newCV", 0, gCV.size(), &gCV[0]);
    for (int d = 0; d < 3; d++) {
        Int n = (d == 0) ? NPX : (d == 1) ? NPY : NPZ;
        char nm[32];
        snprintf(nm, 32, "wgl%d", d); rec(nm, 0, n, wgl[d]);
        for (int h = 0; h < 2; h++) {
            snprintf(nm, 32, "psiRef%d", d * 2 + h); rec(nm, 0, n * n, psiRef[d * 2 + h]);
            snprintf(nm, 32, "psiCor%d", d * 2 + h); rec(nm, 0, n * n, psiCor[d * 2 + h]);
        }
    }
    fclose(out);

    /* the reference transfers and writes every field, then the tree and the grid */
    forEachIt(BaseField::allFields, it)
        (*it)->refineField(step, refineMap, coarseMap, cellMap, nCells, oldCV, oldCC, newCC);
    {
        std::stringstream path;
        path << "amrTree_" << step << ".bin";
        Util::ofstream_bin os(path.str());
        os << gAmrTree;
    }
    {
        std::stringstream path;
        path << gMeshName << "_" << step << ".bin";
        Util::ofstream_bin os(path.str());
        os << gMesh;
    }
    BaseField::destroyFields();
    return 0;
}
