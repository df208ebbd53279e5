// io.cpp -- NebulaSEM on-disk formats: controls, grid, field files (text and binary).
//
// The reference parses both encodings with the same templated grammar; the binary stream
// (Util::ofstream_bin/ifstream_bin, src/util/util.h:191-256) writes strings as 1-byte length + bytes with all
// whitespace stripped (empty strings emit nothing), single chars as [1][c], Int as u32, Scalar as f64.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

#include "nsem_host.h"

namespace nsemh {

static bool file_exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}

// ---------------------------------------------------------------------------------------------------------
// Tokens
// ---------------------------------------------------------------------------------------------------------
Tokens Tokens::from_text(const std::string& text) {
    Tokens t;
    std::string cur;
    bool comment = false;
    for (char ch : text) {
        if (comment) {
            if (ch == '\n') comment = false;
            continue;
        }
        if (ch == '#') {           // Util::nextc skips '#' comments (util.cpp:18-29)
            comment = true;
            if (!cur.empty()) { t.tok_.push_back(cur); cur.clear(); }
            continue;
        }
        if (ch == '{' || ch == '}') {
            if (!cur.empty()) { t.tok_.push_back(cur); cur.clear(); }
            t.tok_.push_back(std::string(1, ch));
        } else if (std::isspace((unsigned char)ch)) {
            if (!cur.empty()) { t.tok_.push_back(cur); cur.clear(); }
        } else {
            cur.push_back(ch);
        }
    }
    if (!cur.empty()) t.tok_.push_back(cur);
    return t;
}

Tokens Tokens::from_binary(std::vector<unsigned char> bytes) {
    Tokens t;
    t.binary = true;
    t.bin_ = std::move(bytes);
    return t;
}

Tokens Tokens::open(const std::string& path_noext) {
    if (file_exists(path_noext + ".txt")) {
        std::ifstream is(path_noext + ".txt");
        std::stringstream ss;
        ss << is.rdbuf();
        return from_text(ss.str());
    }
    if (file_exists(path_noext + ".bin")) {
        std::ifstream is(path_noext + ".bin", std::ios::binary);
        std::vector<unsigned char> b((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
        return from_binary(std::move(b));
    }
    throw Error("cannot open " + path_noext + ".{txt,bin}");
}

bool Tokens::eof() const { return binary ? pos_ >= bin_.size() : pos_ >= tok_.size(); }

std::string Tokens::word() {
    if (eof()) throw Error("unexpected end of file");
    if (!binary) return tok_[pos_++];
    const size_t n = bin_[pos_];
    if (pos_ + 1 + n > bin_.size()) throw Error("truncated binary string");
    std::string s((const char*)&bin_[pos_ + 1], n);
    pos_ += 1 + n;
    return s;
}
u32 Tokens::uint() {
    if (!binary) return (u32)std::stoul(word());
    if (pos_ + 4 > bin_.size()) throw Error("truncated binary u32");
    u32 v;
    std::memcpy(&v, &bin_[pos_], 4);
    pos_ += 4;
    return v;
}
double Tokens::real() {
    if (!binary) return std::stod(word());
    if (pos_ + 8 > bin_.size()) throw Error("truncated binary f64");
    double v;
    std::memcpy(&v, &bin_[pos_], 8);
    pos_ += 8;
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// Controls: `name { key value... }` blocks (Util::read_params, util.cpp:32-79)
// ---------------------------------------------------------------------------------------------------------
static bool looks_like_key(const std::string& w) { return !w.empty() && (std::isalpha((unsigned char)w[0]) || w[0] == '_'); }
static bool complete(const std::vector<std::string>& v) {
    return !v.empty() && std::count(v.begin(), v.end(), "{") == std::count(v.begin(), v.end(), "}");
}

Controls Controls::read(const std::string& path) {
    std::ifstream is(path);
    if (!is) throw Error("cannot open controls file " + path);
    std::stringstream ss;
    ss << is.rdbuf();
    Tokens t = Tokens::from_text(ss.str());
    Controls c;
    while (!t.eof()) {
        const std::string name = t.word();
        if (t.word() != "{") throw Error("controls: expected '{' after block name " + name);
        auto& blk = c.blocks[name];
        std::string key;
        int depth = 1;
        while (depth) {
            const std::string w = t.word();
            if (w == "{") { depth++; blk[key].push_back(w); }
            else if (w == "}") { depth--; if (depth) blk[key].push_back(w); }
            else if (depth == 1 && looks_like_key(w) && (key.empty() || complete(blk[key]))) { key = w; blk[key]; }
            else blk[key].push_back(w);
        }
    }
    return c;
}
bool Controls::has(const std::string& b, const std::string& k) const {
    auto it = blocks.find(b);
    return it != blocks.end() && it->second.count(k) && !it->second.at(k).empty();
}
std::string Controls::str(const std::string& b, const std::string& k, const std::string& def) const {
    return has(b, k) ? blocks.at(b).at(k)[0] : def;
}
double Controls::num(const std::string& b, const std::string& k, double def) const {
    return has(b, k) ? std::stod(blocks.at(b).at(k)[0]) : def;
}
long Controls::integer(const std::string& b, const std::string& k, long def) const {
    return has(b, k) ? std::stol(blocks.at(b).at(k)[0]) : def;
}
bool Controls::yes(const std::string& b, const std::string& k, bool def) const {
    if (!has(b, k)) return def;
    std::string s = blocks.at(b).at(k)[0];
    std::transform(s.begin(), s.end(), s.begin(), ::toupper);
    return s == "YES" || s == "1" || s == "TRUE";
}
Vec3 Controls::vec(const std::string& b, const std::string& k, Vec3 def) const {
    if (!has(b, k) || blocks.at(b).at(k).size() < 3) return def;
    const auto& v = blocks.at(b).at(k);
    return Vec3{std::stod(v[0]), std::stod(v[1]), std::stod(v[2])};
}

// ---------------------------------------------------------------------------------------------------------
// Grid (MeshObject::readTextMesh / writeTextMesh, mesh.h:158-220)
// ---------------------------------------------------------------------------------------------------------
Grid read_grid(const std::string& path_noext) {
    Tokens t = Tokens::open(path_noext);
    Grid g;
    const u32 nv = t.uint();
    t.sym();
    g.V.resize(nv);
    for (u32 i = 0; i < nv; i++)
        for (int d = 0; d < 3; d++) g.V[i][d] = t.real();
    t.sym();
    const u32 nf = t.uint();
    t.sym();
    for (u32 i = 0; i < nf; i++) {
        const u32 n = t.uint();
        t.sym();
        for (u32 j = 0; j < n; j++) g.facetVerts.push_back(t.uint());
        t.sym();
        g.facetStart.push_back((u32)g.facetVerts.size());
    }
    t.sym();
    const u32 nc = t.uint();
    t.sym();
    for (u32 i = 0; i < nc; i++) {
        const u32 n = t.uint();
        t.sym();
        for (u32 j = 0; j < n; j++) g.cellFaces.push_back(t.uint());
        t.sym();
        g.cellStart.push_back((u32)g.cellFaces.size());
    }
    t.sym();
    const u32 nb = t.uint();
    t.sym();
    for (u32 i = 0; i < nb; i++) {
        const std::string name = t.word();
        const u32 n = t.uint();
        t.sym();
        std::vector<u32> faces(n);
        for (u32 j = 0; j < n; j++) faces[j] = t.uint();
        t.sym();
        auto& dst = g.boundaries[name];
        dst.insert(dst.begin(), faces.begin(), faces.end());   // mesh.h:176-177 inserts at the front
    }
    return g;
}

void write_grid_text(const std::string& path, const Grid& g) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw Error("cannot write " + path);
    std::fprintf(f, "%u\n{\n", (u32)g.V.size());
    for (const auto& v : g.V) std::fprintf(f, "%.17g %.17g %.17g\n", v[0], v[1], v[2]);
    std::fprintf(f, "}\n%u\n{\n", g.nFacets());
    for (u32 i = 0; i < g.nFacets(); i++) {
        std::fprintf(f, "%u{ ", g.facetStart[i + 1] - g.facetStart[i]);
        for (u32 j = g.facetStart[i]; j < g.facetStart[i + 1]; j++) std::fprintf(f, "%u ", g.facetVerts[j]);
        std::fprintf(f, "}\n");
    }
    std::fprintf(f, "}\n%u\n{\n", g.nCells());
    for (u32 i = 0; i < g.nCells(); i++) {
        std::fprintf(f, "%u{ ", g.cellStart[i + 1] - g.cellStart[i]);
        for (u32 j = g.cellStart[i]; j < g.cellStart[i + 1]; j++) std::fprintf(f, "%u ", g.cellFaces[j]);
        std::fprintf(f, "}\n");
    }
    std::fprintf(f, "}\n%u\n{\n", (u32)g.boundaries.size());
    for (const auto& kv : g.boundaries) {
        std::fprintf(f, "%s %u\n{ ", kv.first.c_str(), (u32)kv.second.size());
        for (u32 x : kv.second) std::fprintf(f, "%u ", x);
        std::fprintf(f, "}\n");
    }
    std::fprintf(f, "}\n");
    std::fclose(f);
}

// the same grammar through Util::ofstream_bin (util.h:191-226): u32 counts, f64 coordinates, every symbol and name a 1-byte length + bytes
void write_grid_binary(const std::string& path, const Grid& g) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw Error("cannot write " + path);
    auto sym = [&](const std::string& w) { const unsigned char n = (unsigned char)w.size(); std::fwrite(&n, 1, 1, f); std::fwrite(w.data(), 1, n, f); };
    auto u = [&](u32 v) { std::fwrite(&v, 4, 1, f); };
    u((u32)g.V.size()); sym("{");
    for (const auto& v : g.V) std::fwrite(v.data(), 8, 3, f);
    sym("}");
    u(g.nFacets()); sym("{");
    for (u32 i = 0; i < g.nFacets(); i++) {
        u(g.facetStart[i + 1] - g.facetStart[i]); sym("{");
        std::fwrite(g.facetVerts.data() + g.facetStart[i], 4, g.facetStart[i + 1] - g.facetStart[i], f);
        sym("}");
    }
    sym("}");
    u(g.nCells()); sym("{");
    for (u32 i = 0; i < g.nCells(); i++) {
        u(g.cellStart[i + 1] - g.cellStart[i]); sym("{");
        std::fwrite(g.cellFaces.data() + g.cellStart[i], 4, g.cellStart[i + 1] - g.cellStart[i], f);
        sym("}");
    }
    sym("}");
    u((u32)g.boundaries.size()); sym("{");
    for (const auto& kv : g.boundaries) {
        sym(kv.first); u((u32)kv.second.size()); sym("{");
        std::fwrite(kv.second.data(), 4, kv.second.size(), f);
        sym("}");
    }
    sym("}");
    if (std::fclose(f) != 0) throw Error("short write of " + path);
}

// ---------------------------------------------------------------------------------------------------------
// Field files (readInternal_/readBoundary_ field.h:1412-1552; writeInternal_/writeBoundary_ :1588-1663)
// ---------------------------------------------------------------------------------------------------------
FieldFile read_field(const std::string& path_noext, int comps_expected) {
    Tokens t = Tokens::open(path_noext);
    FieldFile ff;
    if (t.word() != "size") throw Error(path_noext + ": expected 'size'");
    ff.comps = (int)t.uint();
    if (ff.comps != comps_expected) throw Error(path_noext + ": unexpected number of components");
    if (t.word() != "internal") throw Error(path_noext + ": Internal field not found");
    const u32 n = t.uint();
    t.sym();
    const int c = ff.comps;
    auto reals = [&](FieldFile::Init& in, int k) { for (int q = 0; q < k; q++) in.a.push_back(t.real()); };
    if (n <= 4) {
        for (u32 q = 0; q < n; q++) {
            FieldFile::Init in;
            in.kind = t.word();
            if (in.kind == "uniform") reals(in, c);
            else if (in.kind == "cosine" || in.kind == "cosine2" || in.kind == "gaussian" || in.kind == "linear") reals(in, 2 * c + 6);
            else if (in.kind == "gaussian-outside") reals(in, 2 * c + 5);
            else if (in.kind == "hydrostatic") reals(in, c + 2);
            else throw Error("Unknown initialization name: " + in.kind);
            ff.inits.push_back(in);
        }
    } else {
        ff.values.resize((size_t)n * c);
        for (auto& v : ff.values) v = t.real();
    }
    t.sym();
    if (t.word() != "boundary") throw Error(path_noext + ": expected 'boundary'");
    const u32 nb = t.uint();
    t.sym();
    for (u32 q = 0; q < nb; q++) {
        BCond bc;
        bc.patch = t.word();
        t.sym();
        while (true) {
            const std::string key = t.word();
            if (key == "}") break;
            if (key == "type") bc.type = t.word();
            else if (key == "value") for (int d = 0; d < c; d++) bc.value[d] = t.real();
            else if (key == "shape") bc.shape = t.real();
            else if (key == "tvalue") for (int d = 0; d < c; d++) bc.tvalue[d] = t.real();
            else if (key == "tshape") bc.tshape = t.real();
            else if (key == "dir") for (int d = 0; d < 3; d++) bc.dir[d] = t.real();
            else if (key == "zMin") bc.zMin = t.real();
            else if (key == "zMax") bc.zMax = t.real();
            else if (key == "neighbor") bc.neighbor = t.word();
            else if (key == "fixed") {
                const u32 m = t.uint();
                t.sym();
                bc.fixed.resize((size_t)m * c);
                for (auto& v : bc.fixed) v = t.real();
                t.sym();
            } else if (key == "E" || key == "kappa" || key == "ks" || key == "cks") t.real();
        }
        ff.bcs.push_back(bc);
    }
    return ff;
}

namespace {
struct BinOut {
    FILE* f;
    void str(const std::string& s0) {
        std::string s;
        for (char ch : s0) if (!std::isspace((unsigned char)ch)) s.push_back(ch);
        if (s.empty()) return;
        const unsigned char n = (unsigned char)s.size();
        std::fwrite(&n, 1, 1, f);
        std::fwrite(s.data(), 1, n, f);
    }
    void u(u32 v) { std::fwrite(&v, 4, 1, f); }
    void d(double v) { std::fwrite(&v, 8, 1, f); }
};
}  // namespace

void write_field(const std::string& path_noext, bool binary, int comps, const double* v, uint64_t n_nodes,
                 const std::vector<BCond>& all_bcs) {
    // conditions the set-up added for patches the field file does not list (hold_unlisted_patches) are not part of the file
    std::vector<BCond> bcs;
    for (const auto& b : all_bcs) if (!b.held) bcs.push_back(b);
    auto near_zero = [](double x) { return std::fabs(x) <= 1e-7; };
    if (binary) {
        FILE* f = std::fopen((path_noext + ".bin").c_str(), "wb");
        if (!f) throw Error("cannot write " + path_noext + ".bin");
        BinOut o{f};
        o.str("size"); o.u((u32)comps);
        o.str("internal"); o.u((u32)n_nodes); o.str("{");
        std::fwrite(v, 8, (size_t)n_nodes * comps, f);
        o.str("}");
        o.str("boundary"); o.u((u32)bcs.size()); o.str("{");
        for (const auto& b : bcs) {
            o.str(b.patch); o.str("{");
            o.str("type"); o.str(b.type);
            double mv = 0, mt = 0;
            for (int d = 0; d < comps; d++) { mv += b.value[d] * b.value[d]; mt += b.tvalue[d] * b.tvalue[d]; }
            if (!near_zero(std::sqrt(mv))) { o.str("value"); for (int d = 0; d < comps; d++) o.d(b.value[d]); }
            if (!near_zero(b.shape)) { o.str("shape"); o.d(b.shape); }
            if (!near_zero(std::sqrt(mt))) { o.str("tvalue"); for (int d = 0; d < comps; d++) o.d(b.tvalue[d]); }
            if (!near_zero(b.tshape)) { o.str("tshape"); o.d(b.tshape); }
            if (!(near_zero(b.dir[0]) && near_zero(b.dir[1]) && near_zero(b.dir[2] - 1))) { o.str("dir"); for (int d = 0; d < 3; d++) o.d(b.dir[d]); }
            if (b.zMax > 0) { o.str("zMin"); o.d(b.zMin); o.str("zMax"); o.d(b.zMax); }
            if (!b.neighbor.empty()) { o.str("neighbor"); o.str(b.neighbor); }
            o.str("}");
        }
        o.str("}");
        std::fclose(f);
    } else {
        FILE* f = std::fopen((path_noext + ".txt").c_str(), "w");
        if (!f) throw Error("cannot write " + path_noext + ".txt");
        std::fprintf(f, "size %d\ninternal %llu\n{\n", comps, (unsigned long long)n_nodes);
        for (uint64_t i = 0; i < n_nodes; i++) {
            for (int d = 0; d < comps; d++) std::fprintf(f, "%.12g ", v[i * comps + d]);
            std::fprintf(f, "\n");
        }
        std::fprintf(f, "}\nboundary %u\n{\n", (u32)bcs.size());
        for (const auto& b : bcs) {
            std::fprintf(f, "%s\n{\n\ttype %s\n", b.patch.c_str(), b.type.c_str());
            double mv = 0, mt = 0;
            for (int d = 0; d < comps; d++) { mv += b.value[d] * b.value[d]; mt += b.tvalue[d] * b.tvalue[d]; }
            auto put = [&](const char* key, const double* x, int n) {      // scalars print bare, vectors with a space after every component
                std::fprintf(f, "\t%s ", key);
                if (n == 1) std::fprintf(f, "%.12g", x[0]);
                else for (int d = 0; d < n; d++) std::fprintf(f, "%.12g ", x[d]);
                std::fprintf(f, "\n");
            };
            if (!near_zero(std::sqrt(mv))) put("value", b.value, comps);
            if (!near_zero(b.shape)) put("shape", &b.shape, 1);
            if (!near_zero(std::sqrt(mt))) put("tvalue", b.tvalue, comps);
            if (!near_zero(b.tshape)) put("tshape", &b.tshape, 1);
            if (!(near_zero(b.dir[0]) && near_zero(b.dir[1]) && near_zero(b.dir[2] - 1))) put("dir", b.dir.data(), 3);
            if (b.zMax > 0) { put("zMin", &b.zMin, 1); put("zMax", &b.zMax, 1); }
            if (!b.neighbor.empty()) std::fprintf(f, "\tneighbor %s\n", b.neighbor.c_str());
            std::fprintf(f, "}\n");
        }
        std::fprintf(f, "}\n");
        std::fclose(f);
    }
}

// ---------------------------------------------------------------------------------------------------------
// VTK (Vtk::write_vtk, src/vtk/vtk.cpp:41-287, the DG branch `NPMAT > 1` :125-286; field blocks MeshField::writeVtkCellAll, field.h:1059-1071)
// ---------------------------------------------------------------------------------------------------------
void write_vtk(const std::string& path, const Basis& b, const double* cC, u32 nBCS, const std::vector<VtkField>& fields,
               bool write_cell_value, bool write_polyhedral) {
    const u32 NPX = (u32)b.NPX, NPY = (u32)b.NPY, NPZ = (u32)b.NPZ, NP = (u32)b.NP;
    if (NP <= 1) throw Error("write_vtk: the finite-volume branch (one node per cell) of Vtk::write_vtk is outside the dGSEM path");
    std::ofstream of(path);
    if (!of) throw Error("cannot write " + path);
    std::vector<char> buf(1 << 22);
    of.rdbuf()->pubsetbuf(buf.data(), (std::streamsize)buf.size());
    const uint64_t nNodes = (uint64_t)nBCS * NP;
    of << (write_polyhedral ? "# vtk DataFile Version 2.0\n" : "# vtk DataFile Version 1.0\n");
    of << "Data produced by FV/DG solver consisting of fields and grid.\nASCII\nDATASET UNSTRUCTURED_GRID\n";
    of << "POINTS " << nNodes << " double\n";
    of.precision(12);
    for (uint64_t i = 0; i < nNodes; i++) of << cC[i * 3] << " " << cC[i * 3 + 1] << " " << cC[i * 3 + 2] << " \n";
    of.precision(6);
    // every element is cut into the sub-cells spanned by neighbouring LGL nodes: points, lines, quads or hexahedra (vtk.cpp:136-264)
    const int ext[3] = {NPX > 1, NPY > 1, NPZ > 1};
    const int dim = ext[0] + ext[1] + ext[2];
    const u32 nr = NPX > 1 ? NPX - 1 : 1, ns = NPY > 1 ? NPY - 1 : 1, nt = NPZ > 1 ? NPZ - 1 : 1;
    const uint64_t ncells = (uint64_t)nBCS * nr * ns * nt;
    const int nverts = 1 << dim;
    static const int vtk_type[4] = {1, 3, 9, 12};
    auto idx = [&](u32 c, u32 r, u32 s, u32 t) { return (uint64_t)c * NP + (uint64_t)r * NPY * NPZ + (uint64_t)s * NPZ + t; };
    // corner offsets (dr,ds,dt) in the reference's order: the first extended axis runs first, quads and hexahedra counter-clockwise
    int axes[3], na = 0;
    for (int d = 0; d < 3; d++) if (ext[d]) axes[na++] = d;
    int off[8][3] = {};
    if (dim == 1) off[1][axes[0]] = 1;
    else if (dim >= 2) {
        static const int q[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
        for (int k = 0; k < nverts; k++) {
            off[k][axes[0]] = q[k & 3][0];
            off[k][axes[1]] = q[k & 3][1];
            if (dim == 3) off[k][axes[2]] = k >> 2;
        }
    }
    of << "CELLS " << ncells << " " << ncells * (uint64_t)(nverts + 1) << "\n";
    for (u32 c = 0; c < nBCS; c++)
        for (u32 r = 0; r < nr; r++)
            for (u32 s = 0; s < ns; s++)
                for (u32 t = 0; t < nt; t++) {
                    of << nverts << " ";
                    for (int k = 0; k < nverts; k++) of << idx(c, r + off[k][0], s + off[k][1], t + off[k][2]) << " ";
                    of << "\n";
                }
    of << "CELL_TYPES " << ncells << "\n";
    for (uint64_t i = 0; i < ncells; i++) of << vtk_type[dim] << "\n";
    if (write_cell_value) {
        of << "CELL_DATA " << ncells << "\nFIELD attributes 1\ncellID  1 " << ncells << " int\n";
        for (uint64_t i = 0; i < ncells; i++) of << i << "\n";
    }
    of << "POINT_DATA " << nNodes << "\nFIELD attributes " << fields.size() << "\n";
    // forEachCellField order (field.h:1192-1197): the scalar fields first, then the vector fields, each in the order of the list
    for (int want : {1, 3, 6, 9})
        for (const auto& f : fields) {
            if (f.comps != want) continue;
            of << f.name << " " << f.comps << " " << nNodes << " double\n";
            for (uint64_t i = 0; i < nNodes; i++) {
                if (f.comps == 1) of << f.v[i] << "\n";
                else {
                    for (int d = 0; d < f.comps; d++) of << f.v[i * f.comps + d] << " ";
                    of << "\n";
                }
            }
            of << "\n";
        }
    of.close();
    if (!of) throw Error("short write of " + path);
}

}  // namespace nsemh
