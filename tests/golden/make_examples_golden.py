"""Generate tests/golden/examples/: the reference's OWN euler example directories (every non-spherical `solver euler` case under
/root/reference/examples), meshed by the reference's `mesh` tool and run for NSTEPS steps by the UNMODIFIED reference binary
(oracle/_ref/parity/{mesh,euler}).

Run in the build container (where /root/reference exists and oracle/build_ref.sh has been run):
    python tests/golden/make_examples_golden.py
Each fixture = the example's case files as the reference reads them (controls, rho/U/T/p0.txt) with only the run length changed
(end_step = write_interval = NSTEPS, write_format BINARY, amr_step removed: fixed mesh), the grid the reference's mesher made of the
example's block file (grid_0.bin) and the reference's dump after NSTEPS steps (expected.npz).  The spherical examples (acoustic-sphere*,
hydro-sphere) are outside the path.
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

from oracle import refio  # noqa: E402
from oracle.mesh import MeshTopo  # noqa: E402

EXAMPLES = "/root/reference/examples"
NSTEPS = 3
AMR_CASES = ("isentropic", "srtb-3d", "srtb-amr", "srtb-amr-hill")    # the amr_step examples whose regrid the reference survives (it
                                                                       # segfaults on srtb-amr-zaxis in this build)
CASES = ["isentropic", "atmo/ctbs", "atmo/dc", "atmo/lrtb", "atmo/srtb", "atmo/srtb-3d", "atmo/srtb-amr", "atmo/srtb-amr-hill",
         "atmo/srtb-amr-zaxis", "atmo/srtb-curved", "atmo/srtb-inclined"]


def main():
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "examples")
    for ex in CASES:
        name = os.path.basename(ex)
        d = os.path.join(tempfile.mkdtemp(prefix="ex_golden_"), name)
        shutil.copytree(os.path.join(EXAMPLES, ex), d)
        try:
            block = [f for f in os.listdir(d) if f != "controls" and not f.endswith((".txt", ".sh"))][0]
            m = subprocess.run([run_ref.ref_bin("mesh"), block, "-o", "grid_0.bin"], cwd=d, capture_output=True, text=True, timeout=600)
            if m.returncode != 0 or not os.path.exists(os.path.join(d, "grid_0.bin")):
                raise RuntimeError(m.stdout[-1000:] + m.stderr[-1000:])
            ctl = open(os.path.join(d, "controls")).read()
            ctl = re.sub(r"(?m)^(\s*)end_step\s+\d+", rf"\g<1>end_step {NSTEPS}", ctl)
            ctl = re.sub(r"(?m)^(\s*)write_interval\s+\d+", rf"\g<1>write_interval {NSTEPS}", ctl)
            ctl = re.sub(r"(?m)^(\s*)write_format\s+\w+", r"\g<1>write_format BINARY", ctl)
            ctl = re.sub(r"(?m)^\s*amr_step\s+\d+\s*\n", "", ctl)
            open(os.path.join(d, "controls"), "w").write(ctl)
            dst = os.path.join(out_dir, name)
            shutil.rmtree(dst, ignore_errors=True)
            os.makedirs(dst)
            for f in ("controls", "grid_0.bin", "rho0.txt", "U0.txt", "T0.txt", "p0.txt"):      # before the run: problem_init rewrites the *0 files
                shutil.copy(os.path.join(d, f), os.path.join(dst, f))
            run_ref.run_euler(d, variant="parity", timeout=900)
            dump = run_ref.read_dump(d, 1)
            np.savez_compressed(os.path.join(dst, "expected.npz"), nsteps=NSTEPS, rho=dump["rho"], U=dump["U"], T=dump["T"], p=dump["p"])
            print(name, {k: v.shape for k, v in dump.items()})
            if name in AMR_CASES:
                # the example as it ships (amr_step kept): the reference regrids before step 1 (Prepare::refineMesh: tagging by the
                # refinement{} block + MeshObject::refineMesh) and overwrites grid_0; keep the cells of that grid (centroid, volume)
                shutil.rmtree(d)
                shutil.copytree(os.path.join(EXAMPLES, ex), d)
                shutil.copy(os.path.join(dst, "grid_0.bin"), d)
                ctl = open(os.path.join(d, "controls")).read()
                ctl = re.sub(r"(?m)^(\s*)end_step\s+\d+", r"\g<1>end_step 1", ctl)
                ctl = re.sub(r"(?m)^(\s*)write_interval\s+\d+", r"\g<1>write_interval 1", ctl)
                ctl = re.sub(r"(?m)^(\s*)write_format\s+\w+", r"\g<1>write_format BINARY", ctl)
                open(os.path.join(d, "controls"), "w").write(ctl)
                run_ref.run_euler(d, variant="parity", timeout=900)
                topo = MeshTopo(refio.read_grid(os.path.join(d, "grid_0"))).load()
                nb = topo.nBCS
                np.savez_compressed(os.path.join(dst, "initial_regrid.npz"), CC=np.asarray(topo.CC)[:nb], CV=np.asarray(topo.CV)[:nb],
                                    n_facets=len(topo.facets), n_mortar=int(np.count_nonzero(np.asarray(topo.FMC))),
                                    refinement=re.findall(r"(?s)refinement\s*\{(.*?)\}", ctl)[0])
                print("   initial regrid:", nb, "cells", len(topo.facets), "facets")
        finally:
            shutil.rmtree(os.path.dirname(d), ignore_errors=True)


if __name__ == "__main__":
    main()
