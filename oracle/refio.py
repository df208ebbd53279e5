"""Readers/writers for the NebulaSEM on-disk formats -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this package.

Formats restated from the reference:
  * binary stream grammar      src/util/util.h:191-256   (strings = 1-byte length + bytes with whitespace stripped,
                                                           empty strings emit nothing, chars = [1][c], Int = u32, Scalar = f64)
  * grid file                  src/mesh/mesh.h:158-220   (vertices, facets, cells, boundaries)
  * field file                 src/field/field.h:1412-1552 (read), :1588-1663 (write)
  * geomdump records           oracle/tools/geomdump.cpp (own tool linked against the reference objects)
"""
from __future__ import annotations

import re
import struct
from dataclasses import dataclass, field

import numpy as np


# --------------------------------------------------------------------------------------------
# token streams: the reference parses text and binary files with the same templated grammar
# --------------------------------------------------------------------------------------------
class TextTokens:
    """Whitespace-delimited tokens of a text file; '{'/'}' glued to a count ("4{0 1 2 3}") are split off."""

    def __init__(self, text: str):
        # strip '#' comments (Util::nextc, util.cpp:18-29, only does so where it is called; the grid and field
        # files written by the reference carry no comments, block-mesh inputs do)
        text = re.sub(r"#[^\n]*", " ", text)
        text = text.replace("{", " { ").replace("}", " } ")
        self.tok = text.split()
        self.pos = 0

    def word(self) -> str:
        t = self.tok[self.pos]
        self.pos += 1
        return t

    def u32(self) -> int:
        return int(self.word())

    def f64(self) -> float:
        return float(self.word())

    def sym(self) -> str:
        return self.word()

    def f64_array(self, n: int) -> np.ndarray:
        a = np.array(self.tok[self.pos:self.pos + n], dtype=np.float64)
        self.pos += n
        return a

    def u32_array(self, n: int) -> np.ndarray:
        a = np.array(self.tok[self.pos:self.pos + n], dtype=np.int64)
        self.pos += n
        return a

    def eof(self) -> bool:
        return self.pos >= len(self.tok)


class BinTokens:
    def __init__(self, data: bytes):
        self.b = data
        self.pos = 0

    def word(self) -> str:
        n = self.b[self.pos]
        s = self.b[self.pos + 1:self.pos + 1 + n].decode()
        self.pos += 1 + n
        return s

    sym = word

    def u32(self) -> int:
        v = struct.unpack_from("<I", self.b, self.pos)[0]
        self.pos += 4
        return v

    def f64(self) -> float:
        v = struct.unpack_from("<d", self.b, self.pos)[0]
        self.pos += 8
        return v

    def f64_array(self, n: int) -> np.ndarray:
        a = np.frombuffer(self.b, dtype="<f8", count=n, offset=self.pos).copy()
        self.pos += 8 * n
        return a

    def u32_array(self, n: int) -> np.ndarray:
        a = np.frombuffer(self.b, dtype="<u4", count=n, offset=self.pos).astype(np.int64)
        self.pos += 4 * n
        return a

    def eof(self) -> bool:
        return self.pos >= len(self.b)


def open_tokens(path_noext: str):
    """Mirror of the .txt-then-.bin probing in Mesh::LoadMesh (field.cpp:108-117) / MeshField::read (field.h:1578-1584)."""
    import os
    if os.path.exists(path_noext + ".txt") and not path_noext.endswith("__force_bin__"):
        with open(path_noext + ".txt") as f:
            return TextTokens(f.read())
    with open(path_noext + ".bin", "rb") as f:
        return BinTokens(f.read())


# --------------------------------------------------------------------------------------------
# grid files
# --------------------------------------------------------------------------------------------
@dataclass
class Grid:
    vertices: np.ndarray                      # (nv,3) f64
    facets: list                              # list of int lists (vertex ids)
    cells: list                               # list of int lists (facet ids), real cells only
    boundaries: dict = field(default_factory=dict)   # name -> list of facet ids (insertion = file order)


def read_grid(path_noext: str, prefer_bin: bool = False) -> Grid:
    import os
    if prefer_bin or not os.path.exists(path_noext + ".txt"):
        with open(path_noext + ".bin", "rb") as f:
            t = BinTokens(f.read())
    else:
        with open(path_noext + ".txt") as f:
            t = TextTokens(f.read())
    nv = t.u32(); t.sym()
    verts = t.f64_array(3 * nv).reshape(nv, 3)
    t.sym()
    nf = t.u32(); t.sym()
    facets = []
    for _ in range(nf):
        n = t.u32(); t.sym()
        facets.append([int(x) for x in t.u32_array(n)])
        t.sym()
    t.sym()
    nc = t.u32(); t.sym()
    cells = []
    for _ in range(nc):
        n = t.u32(); t.sym()
        cells.append([int(x) for x in t.u32_array(n)])
        t.sym()
    t.sym()
    nb = t.u32(); t.sym()
    bnd = {}
    for _ in range(nb):
        name = t.word()
        n = t.u32(); t.sym()
        faces = [int(x) for x in t.u32_array(n)]
        t.sym()
        # readTextMesh inserts at the FRONT of an existing list of the same name (mesh.h:176-177)
        bnd[name] = faces + bnd.get(name, [])
    return Grid(verts, facets, cells, bnd)


def write_grid_text(path: str, g: Grid) -> None:
    """Text grid in the grammar of MeshObject::writeTextMesh (mesh.h:203-220)."""
    out = []
    out.append(f"{len(g.vertices)}\n{{")
    for v in g.vertices:
        out.append(f"{float(v[0])!r} {float(v[1])!r} {float(v[2])!r}")
    out.append("}")
    out.append(f"{len(g.facets)}\n{{")
    for f in g.facets:
        out.append(f"{len(f)}{{ " + " ".join(str(x) for x in f) + " }")
    out.append("}")
    out.append(f"{len(g.cells)}\n{{")
    for c in g.cells:
        out.append(f"{len(c)}{{ " + " ".join(str(x) for x in c) + " }")
    out.append("}")
    out.append(f"{len(g.boundaries)}\n{{")
    for name, faces in g.boundaries.items():
        out.append(f"{name} {len(faces)}\n{{ " + " ".join(str(x) for x in faces) + " }")
    out.append("}")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


# --------------------------------------------------------------------------------------------
# field files
# --------------------------------------------------------------------------------------------
@dataclass
class BC:
    """One `patch { type ... }` entry (BCondition, field.h:144-267)."""
    patch: str
    kind: str = ""
    value: np.ndarray | None = None
    shape: float = 0.0
    tvalue: np.ndarray | None = None
    tshape: float = 0.0
    dir: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 1.0]))
    zMin: float = 0.0
    zMax: float = 0.0
    neighbor: str = ""
    fixed: np.ndarray | None = None


@dataclass
class FieldFile:
    comps: int
    initializers: list | None          # list of (name, params...) when `internal N<=4`
    values: np.ndarray | None          # (N, comps) raw node values otherwise
    bcs: list


def read_field(path_noext: str) -> FieldFile:
    t = open_tokens(path_noext)
    assert t.word() == "size"
    comps = t.u32()
    assert t.word() == "internal"
    n = t.u32(); t.sym()
    inits, values = None, None
    if n <= 4:
        inits = []
        for _ in range(n):
            kind = t.word()
            if kind == "uniform":
                inits.append((kind, t.f64_array(comps)))
            elif kind in ("cosine", "cosine2", "gaussian", "linear"):
                value = t.f64_array(comps); pert = t.f64_array(comps)
                center = t.f64_array(3); radius = t.f64_array(3)
                inits.append((kind, value, pert, center, radius))
            elif kind == "gaussian-outside":
                value = t.f64_array(comps); pert = t.f64_array(comps)
                center = t.f64_array(3); r1 = t.f64(); r2 = t.f64()
                inits.append((kind, value, pert, center, r1, r2))
            elif kind == "hydrostatic":
                p0 = t.f64_array(comps); scale = t.f64(); expon = t.f64()
                inits.append((kind, p0, scale, expon))
            else:
                raise ValueError(f"unknown initializer {kind}")
    else:
        values = t.f64_array(n * comps).reshape(n, comps)
    t.sym()
    assert t.word() == "boundary"
    nb = t.u32(); t.sym()
    bcs = []
    for _ in range(nb):
        bc = BC(patch=t.word()); t.sym()
        bc.value = np.zeros(comps); bc.tvalue = np.zeros(comps)
        while True:
            key = t.word()
            if key == "}":
                break
            if key == "type":
                bc.kind = t.word()
            elif key == "value":
                bc.value = t.f64_array(comps)
            elif key == "shape":
                bc.shape = t.f64()
            elif key == "tvalue":
                bc.tvalue = t.f64_array(comps)
            elif key == "tshape":
                bc.tshape = t.f64()
            elif key == "dir":
                bc.dir = t.f64_array(3)
            elif key == "zMin":
                bc.zMin = t.f64()
            elif key == "zMax":
                bc.zMax = t.f64()
            elif key == "neighbor":
                bc.neighbor = t.word()
            elif key == "fixed":
                m = t.u32(); t.sym()
                bc.fixed = t.f64_array(m * comps).reshape(m, comps); t.sym()
            elif key in ("E", "kappa", "ks", "cks"):
                t.f64()
        bcs.append(bc)
    return FieldFile(comps, inits, values, bcs)


def read_field_values(path_noext: str) -> np.ndarray:
    """Raw node values (N, comps) of a dump written by the reference (`rho1.bin`, ...)."""
    ff = read_field(path_noext)
    assert ff.values is not None, "field file holds initializers, not node values"
    return ff.values


# --------------------------------------------------------------------------------------------
# geomdump records
# --------------------------------------------------------------------------------------------
def read_geomdump(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (n,) = struct.unpack_from("<I", data, pos); pos += 4
        name = data[pos:pos + n].decode(); pos += n
        dtype, count = struct.unpack_from("<IQ", data, pos); pos += 12
        if dtype == 0:
            arr = np.frombuffer(data, dtype="<f8", count=count, offset=pos).copy(); pos += 8 * count
        else:
            arr = np.frombuffer(data, dtype="<u4", count=count, offset=pos).astype(np.int64); pos += 4 * count
        out[name] = arr
    return out


# --------------------------------------------------------------------------------------------
# controls files
# --------------------------------------------------------------------------------------------
def read_controls(path: str) -> dict:
    """`name { key value... }` blocks with '#' comments (Util::read_params, util.cpp:32-79). Values stay token lists."""
    with open(path) as f:
        t = TextTokens(f.read())
    blocks = {}
    while not t.eof():
        name = t.word()
        assert t.sym() == "{", f"expected '{{' after {name}"
        blk = {}
        key = None
        depth = 1
        while depth:
            w = t.word()
            if w == "{":
                depth += 1
                blk[key].append(w)
            elif w == "}":
                depth -= 1
                if depth:
                    blk[key].append(w)
            elif depth == 1 and re.match(r"^[A-Za-z_]", w) and (key is None or _complete(blk[key])):
                key = w
                blk[key] = []
            else:
                blk[key].append(w)
        blocks[name] = blk
    return blocks


def _complete(vals: list) -> bool:
    """A key's value list is complete once it is non-empty and its braces balance."""
    return len(vals) > 0 and vals.count("{") == vals.count("}")
