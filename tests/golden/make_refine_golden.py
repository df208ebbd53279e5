"""Generate the AMR field-transfer golden vectors in tests/golden/refine_field/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists and oracle/build_ref.sh has been run):
    python tests/golden/make_refine_golden.py
For each case (2-D order 4 from examples/atmo/srtb-amr, 3-D order 2 from examples/atmo/srtb-3d on a 4^3 mesh):
1. the reference's `mesh` + `euler` run 20 steps on the fixed mesh, so that rho, U, T, p all carry structure;
2. oracle/_ref/parity/refinedump (own driver linked against the reference objects, oracle/tools/refinedump.cpp) regrids
   the case twice through the reference's MeshObject::refineMesh and MeshField::refineField (field.h:1863-2015):
   pass 1 splits a block of cells, pass 2 merges two (one in 3-D) of the new families back, splits two coarse cells and
   one fine cell (a level-2 cell), so that copy, refinement and coarsening all occur with fields that are no longer
   polynomial on the parent;
3. per pass one .npz: what refineField was given (maps, volumes, centroids, old node coordinates, psiRef/psiCor/wgl),
   the fields before, and the fields the reference wrote after the transfer;
4. per stage (0 = uniform grid, 1 = after pass 1, 2 = after pass 2) a fixed-mesh case directory <name>/stage<k>/ with the
   grid and the field files of that stage as the reference wrote them (index 0), which the C++ host opens in the GPU tests.
The reference ships no golden vectors of its own.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refio, run_ref  # noqa: E402
from make_amr_golden import FIXED_CONTROLS, edit_controls  # noqa: E402

REF = os.environ.get("NSEM_REFERENCE", "/root/reference")
NSTEPS = 20
FIELDS = ("rho", "U", "T", "p")


def refinedump(case, cells_txt, out_name):
    open(os.path.join(case, "cells.txt"), "w").write(cells_txt)
    subprocess.check_call([run_ref.ref_bin("refinedump"), "./controls", out_name, "cells.txt", "1"], cwd=case, stdout=subprocess.DEVNULL)
    d = refio.read_geomdump(os.path.join(case, out_name))
    for n in FIELDS:
        d["post:" + n] = refio.read_field_values(os.path.join(case, f"{n}1"))
    return d


def save_stage(case, out_dir, name, k, grid_index, orders):
    """controls + grid + fields of dump 1 as a fixed-mesh case that starts at step 0"""
    st = os.path.join(out_dir, name, f"stage{k}")
    os.makedirs(st, exist_ok=True)
    shutil.copy(os.path.join(case, f"grid_{grid_index}.bin"), os.path.join(st, "grid_0.bin"))
    for n in FIELDS:
        shutil.copy(os.path.join(case, f"{n}1.bin"), os.path.join(st, f"{n}0.bin"))
    with open(os.path.join(st, "controls"), "w") as fh:
        fh.write(FIXED_CONTROLS.format(n=NSTEPS, **orders))


def pack(d):
    out = {}
    for k, v in d.items():
        out[k.replace(":", "_")] = v
    return out


def cells_line(kind, ids):
    return f"{kind} {len(ids)} " + " ".join(str(int(i)) for i in ids) + "\n"


def main():
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refine_field")
    os.makedirs(out_dir, exist_ok=True)
    cases = {
        # name: (example, mesh edit, controls edits, (nx, ny, nz), block of element indices to split in pass 1)
        "2d_o4": ("srtb-amr", None, dict(npx=4, npy=4, npz=0), (10, 1, 10), [(ex, 0, ez) for ex in range(3, 7) for ez in range(2, 6)]),
        "3d_o2": ("srtb-3d", ("wall 3{6 6 6}", "wall 3{4 4 4}"), dict(npx=2, npy=2, npz=2), (4, 4, 4),
                  [(ex, ey, ez) for ex in (1, 2) for ey in (1, 2) for ez in (1, 2)]),
    }
    for name, (example, mesh_edit, ctl, n, block) in cases.items():
        d = tempfile.mkdtemp(prefix="golden_refine_")
        try:
            a = os.path.join(d, "case")
            shutil.copytree(os.path.join(REF, "examples", "atmo", example), a)
            if mesh_edit:
                txt = open(os.path.join(a, "bubble")).read()
                assert mesh_edit[0] in txt
                open(os.path.join(a, "bubble"), "w").write(txt.replace(mesh_edit[0], mesh_edit[1]))
            edit_controls(os.path.join(a, "controls"), end_step=NSTEPS, write_interval=NSTEPS, amr_step=None, **ctl)
            subprocess.check_call([run_ref.ref_bin("mesh"), "bubble", "-o", "grid_0.bin"], cwd=a, stdout=subprocess.DEVNULL)
            run_ref.run_euler(a, variant="parity")                     # fields of dump 1 on the uniform grid
            save_stage(a, out_dir, name, 0, 0, ctl)
            cid = lambda e: (e[0] * n[1] + e[1]) * n[2] + e[2]          # block-generated element order (hexMesh.cpp:322-342)
            # pass 1: split the block
            p1 = refinedump(a, cells_line("r", [cid(e) for e in block]), "pass1.bin")
            save_stage(a, out_dir, name, 1, 1, ctl)
            # pass 2: merge the first (and in 2-D also the last) family again, split two untouched coarse cells and the child of a
            # middle family that lies nearest the centre of the block (all its face neighbours are level-1 cells)
            rm, cm = p1["refineMap"].astype(np.int64), p1["cellMap"].astype(np.int64)
            fams, i = [], 0
            while i < len(rm):
                nch = rm[i]
                fams.append((rm[i + 1], cm[rm[i + 2:i + 2 + nch]]))
                i += nch + 2
            merge = list(fams[0][1]) + (list(fams[-1][1]) if name == "2d_o4" else [])
            newCC = p1["newCC"].reshape(-1, 3)
            centre = newCC[np.concatenate([f[1] for f in fams])].mean(axis=0)
            mid = fams[len(fams) // 2 + (1 if name == "2d_o4" else 0)][1]
            if name == "3d_o2":
                mid = fams[-1][1]
            fine = mid[np.argmin(np.linalg.norm(newCC[mid] - centre, axis=1))]
            coarse_old = [cid((0, 0, 0)), cid((n[0] - 1, n[1] - 1, n[2] - 1))]
            split = [cm[c] for c in coarse_old] + [fine]
            p2 = refinedump(a, cells_line("r", split) + cells_line("c", merge), "pass2.bin")
            save_stage(a, out_dir, name, 2, 1, ctl)
            for tag, p in (("pass1", p1), ("pass2", p2)):
                np.savez_compressed(os.path.join(out_dir, f"{name}_{tag}.npz"), **pack(p))
                dims = p["dims"]
                print(name, tag, "cells", dims[4], "->", dims[6], "refineMap", len(p["refineMap"]), "coarseMap", len(p["coarseMap"]))
        finally:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
