"""Golden fixtures for AMR RUNS by the UNMODIFIED reference binary (oracle/_ref/parity/{mesh,euler}): examples/atmo/srtb-amr as it ships
(BASELINE configs[4]; `amr_step 1`, `write_interval 50`, 100 steps) and a reduced examples/atmo/acoustic-sphere-amr-dg (cubed sphere).

    python tests/golden/make_amr_run_golden.py          (build container: /root/reference + oracle/build_ref.sh)

What the reference does with these controls (iteration.h:94-147, euler.cpp:57-287): the set-up's START branch on the coarse grid and a
dump of all fields (the zero-iteration pass, iteration.h:19-23,44-48), Prepare::refineMesh(0) (tags by the refinement{} block, refines
grid_0 and every field file), then steps 1-50 on that grid from the set-up's RESTART branch, dump 1, Prepare::refineMesh(1), steps 51-100,
dump 2.  The fixture keeps the case files before the run, and of the run: the grid of every regrid (cells only: centroid + volume) and
dump 2 with the node coordinates of its grid (from the oracle's geometry of the reference's grid_1, bit-equal to the reference's)."""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refio, run_ref  # noqa: E402
from oracle.dg import Basis, Geometry  # noqa: E402
from oracle.euler import Params  # noqa: E402
from oracle.mesh import MeshTopo  # noqa: E402

EXAMPLES = "/root/reference/examples/atmo"


def load_topo(grid, params):
    topo = MeshTopo(grid)
    topo.spherical, topo.sphere_radius, topo.sphere_height = params.is_spherical, params.sphere_radius, params.sphere_height
    return topo.load()


def main():
    make("srtb-amr", 100, 50)
    # a cubed-sphere AMR run: examples/atmo/acoustic-sphere-amr-dg at its 10 x 10 cells per panel, order 2, 10 steps per dump (the example:
    # order 4, 240): regrid of the pressure pulse before step 1 (2-D refinement, the radial axis of every cell is never split) and after dump 1
    make("acoustic-sphere-amr-dg", 20, 10, divisions=(10, 10, 1), nop=(2, 2, 0))


def make(name, NSTEPS, INTERVAL, divisions=None, nop=None):
    EX = os.path.join(EXAMPLES, name)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amr_run", name)
    d = os.path.join(tempfile.mkdtemp(prefix="amr_run_"), name)
    shutil.copytree(EX, d)
    for f in os.listdir(d):
        os.chmod(os.path.join(d, f), 0o644)
    block = [f for f in os.listdir(d) if f != "controls" and not f.endswith((".txt", ".sh"))][0]
    if divisions:
        blk = open(os.path.join(d, block)).read()
        blk, cnt = re.subn(r"(?m)^(8\{[^}]*\}\s+linear\s+)3\{\d+ \d+ \d+\}", r"\g<1>3{%d %d %d}" % divisions, blk)
        assert cnt == 6
        open(os.path.join(d, block), "w").write(blk)
    m = subprocess.run([run_ref.ref_bin("mesh"), block, "-o", "grid_0.bin"], cwd=d, capture_output=True, text=True, timeout=600)
    assert m.returncode == 0, m.stdout[-1000:] + m.stderr[-1000:]
    ctl = open(os.path.join(d, "controls")).read()
    if nop:
        for k, v in zip(("npx", "npy", "npz"), nop):
            ctl = re.sub(rf"(?m)^(\s*){k}\s+\d+", rf"\g<1>{k} {v}", ctl)
    ctl = re.sub(r"(?m)^(\s*)end_step\s+\d+", rf"\g<1>end_step {NSTEPS}", ctl)
    ctl = re.sub(r"(?m)^(\s*)write_interval\s+\d+", rf"\g<1>write_interval {INTERVAL}", ctl)
    ctl = re.sub(r"(?m)^(\s*)write_format\s+\w+", r"\g<1>write_format BINARY", ctl)
    assert re.search(r"(?m)^\s*amr_step\s+1\s*$", ctl), "the example ships with amr_step 1"
    open(os.path.join(d, "controls"), "w").write(ctl)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for f in ("controls", "grid_0.bin", "rho0.txt", "U0.txt", "T0.txt", "p0.txt"):
        shutil.copy(os.path.join(d, f), os.path.join(out, f))
    wall, log = run_ref.run_euler(d, variant="parity", timeout=1800)
    print("\n".join(l for l in log.splitlines() if "Refining" in l or "loss" in l)[-1500:])
    nop = [int(re.search(rf"(?m)^\s*{k}\s+(\d+)", ctl).group(1)) for k in ("npx", "npy", "npz")]
    params = Params.from_controls(refio.read_controls(os.path.join(d, "controls")))
    keep = {}
    for k in (0, 1):
        topo = load_topo(refio.read_grid(os.path.join(d, f"grid_{k}"), prefer_bin=True), params)
        nb = topo.nBCS
        keep[f"grid{k}_CC"] = np.asarray(topo.CC)[:nb]
        keep[f"grid{k}_CV"] = np.asarray(topo.CV)[:nb]
        if k == 1:
            geo = Geometry(topo, Basis(nop))
            keep["node_xyz"] = np.asarray(geo.cC)[:geo.gBCSfield]
            keep["node_cV"] = np.asarray(geo.cV)[:geo.gBCSfield]
            keep["NP"] = geo.gBCSfield // nb
    # the same run stopped at the first dump (end_step 50: no second regrid, so <field>1.bin is the state after 50 steps on the grid of
    # the initial regrid; in the full run Prepare::refineMesh(1) overwrites those files with the transferred fields)
    d50 = d + "_50"
    shutil.copytree(out, d50)
    os.chmod(os.path.join(d50, "controls"), 0o644)
    open(os.path.join(d50, "controls"), "w").write(re.sub(r"(?m)^(\s*)end_step\s+\d+", rf"\g<1>end_step {INTERVAL}", ctl))
    run_ref.run_euler(d50, variant="parity", timeout=1800)
    topo0 = load_topo(refio.read_grid(os.path.join(d50, "grid_0"), prefer_bin=True), params)
    geo0 = Geometry(topo0, Basis(nop))
    n0 = geo0.gBCSfield
    half = run_ref.read_dump(d50, 1)
    for f in ("rho0.txt", "U0.txt", "T0.txt", "p0.txt"):      # the case's own analytic initialisers: the .bin files next to them are read
        os.remove(os.path.join(d50, f))
    init = run_ref.read_dump(d50, 0)            # <field>0.bin after Prepare::refineMesh(0): the start-branch state transferred to the regridded grid
    keep.update(init_rho=init["rho"][:n0], init_U=init["U"][:n0], init_T=init["T"][:n0], init_p=init["p"][:n0])
    keep.update(half_node_xyz=np.asarray(geo0.cC)[:n0], half_rho=half["rho"][:n0], half_U=half["U"][:n0], half_T=half["T"][:n0])
    dump = run_ref.read_dump(d, NSTEPS // INTERVAL)
    n = keep["node_xyz"].shape[0]
    assert dump["rho"].shape[0] >= n
    np.savez_compressed(os.path.join(out, "expected.npz"), nsteps=NSTEPS, interval=INTERVAL, rho=dump["rho"][:n], U=dump["U"][:n], T=dump["T"][:n],
                        p=dump["p"][:n], **keep)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in keep.items()}, "wall", wall)
    print(sorted(os.listdir(d)))


if __name__ == "__main__":
    main()
