"""Generate the non-conforming (AMR, mortar-face) golden case in tests/golden/srtb_amr/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists and oracle/build_ref.sh has been run):
    python tests/golden/make_amr_golden.py
1. examples/atmo/srtb-amr (2-D rising bubble, order 4, `amr_step 1`, max_level 2) is meshed with the reference's `mesh`
   and started with the reference's `euler`: the initial regrid (before step 1) refines 32 of the 100 cells
   and leaves a NON-CONFORMING grid of 196 cells with 56 mortar sub-faces in grid_0.bin, plus the fields transferred to
   it (rho0/U0/T0/p0.bin).
2. Those files + the controls without `amr_step` are the fixed-mesh case stored here (inputs); the reference is run on
   it for NSTEPS steps and its binary dumps are stored as expected.npz (outputs).
The reference ships no golden vectors of its own; these dumps pin the oracle's mortar operators
(scatter/gather_non_conforming, field.h:2019-2248; psiRef/psiCor, dg.cpp:550-590).
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import run_ref  # noqa: E402

NSTEPS = 20
REF = os.environ.get("NSEM_REFERENCE", "/root/reference")


FIXED_CONTROLS = """general
{{
    solver euler
    mesh grid
    write_format BINARY
    rho 1.177
    viscosity 1.5
    state TRANSIENT
    start_step 0
    end_step {n}
    write_interval {n}
    dt 0.00125
    n_deferred 0
    convection_scheme RUSANOV
    nonortho_scheme OVER_RELAXED
    time_scheme AB1
    blend_factor 1
    parallel_method BLOCKED
    method PCG
    preconditioner DIAG
    tolerance 1e-5
    max_iterations 6400
    SOR_omega 1.7
    probe 0 {{}}
    gravity 0 -9.80606 0
    npx {npx}
    npy {npy}
    npz {npz}
}}
prepare
{{
    fields 4 {{ U T p rho }}
}}
euler
{{
    velocity_UR 0.5
    pressure_UR 0.8
    t_UR 0.8
    diffusion YES
    buoyancy YES
}}
"""


def edit_controls(path, **kv):
    txt = open(path).read()
    txt = re.sub(r"^\s*print_time.*\n", "", txt, flags=re.M)
    txt = re.sub(r"^\s*write_format.*$", "    write_format BINARY", txt, flags=re.M)
    for k, v in kv.items():
        if v is None:
            txt = re.sub(rf"^\s*{k}\s.*\n", "", txt, flags=re.M)
        else:
            txt = re.sub(rf"^\s*{k}\s.*$", f"    {k} {v}", txt, flags=re.M)
    open(path, "w").write(txt)


# name -> (example directory, mesh-file edit, controls edits for the regrid run, orders of the fixed-mesh controls)
CASES = {
    # 2-D, order 4: 100 -> 196 cells, 56 mortar sub-faces (two per coarse face)
    "srtb_amr": ("srtb-amr", None, {}, dict(npx=4, npy=4, npz=0)),
    # 3-D, order 2: 4^3 = 64 -> 372 cells (44 refined), four mortar sub-faces per coarse face
    "srtb3d_amr": ("srtb-3d", ("wall 3{6 6 6}", "wall 3{4 4 4}"),
                   dict(max_level=2, field_min=0.1, field_max=0.4, npx=2, npy=2, npz=2), dict(npx=2, npy=2, npz=2)),
}


def main(name):
    example, mesh_edit, ctl_edit, orders = CASES[name]
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), name)
    d = tempfile.mkdtemp(prefix="golden_amr_")
    try:
        a = os.path.join(d, "amr")
        shutil.copytree(os.path.join(REF, "examples", "atmo", example), a)
        if mesh_edit:
            txt = open(os.path.join(a, "bubble")).read()
            assert mesh_edit[0] in txt
            open(os.path.join(a, "bubble"), "w").write(txt.replace(mesh_edit[0], mesh_edit[1]))
        edit_controls(os.path.join(a, "controls"), end_step=1, write_interval=1, **ctl_edit)
        subprocess.check_call([run_ref.ref_bin("mesh"), "bubble", "-o", "grid_0.bin"], cwd=a, stdout=subprocess.DEVNULL)
        run_ref.run_euler(a, variant="parity")                    # initial regrid -> non-conforming grid_0.bin + fields
        os.makedirs(out_dir, exist_ok=True)
        for f in ("grid_0.bin", "rho0.bin", "U0.bin", "T0.bin", "p0.bin"):
            shutil.copy(os.path.join(a, f), os.path.join(out_dir, f))
        with open(os.path.join(out_dir, "controls"), "w") as fh:        # the fixed-mesh controls (no amr_step, no refinement block)
            fh.write(FIXED_CONTROLS.format(n=NSTEPS, **orders))
        f = os.path.join(d, "fixed")
        shutil.copytree(out_dir, f)
        run_ref.run_euler(f, variant="parity")
        dump = run_ref.read_dump(f, 1)
        np.savez_compressed(os.path.join(out_dir, "expected.npz"), nsteps=NSTEPS, rho=dump["rho"], U=dump["U"], T=dump["T"], p=dump["p"])
        print({k: v.shape for k, v in dump.items()})
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        main(nm)
