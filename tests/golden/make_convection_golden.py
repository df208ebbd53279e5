"""Golden fixture for the explicit scalar advection app: examples/atmo/advection-leveque of the reference (2-D, order 4,
LeVeque's deformational wind re-evaluated every step, RUSANOV) run by the UNMODIFIED reference binary oracle/_ref/parity/convection.

    python tests/golden/make_convection_golden.py        (build container: /root/reference + oracle/build_ref.sh)

The example's controls are kept except: time_scheme AB1 (the one-stage scheme of the path; the example ships AB2), amr_step removed (fixed
mesh), write_format BINARY, end_step = write_interval = 40.  A second fixture freezes the wind (problem_init NONE) with the example's
field files replaced by a rotating wind given as raw node values."""
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

EX = "/root/reference/examples/atmo/advection-leveque"
NSTEPS = 40


def main():
    make("advection-leveque", "AB1")
    make("advection-leveque-ab2", "AB2")       # the scheme the example ships with: two residuals kept (field.h:3789-3806, 3885-3905)
    make("advection-leveque-ab4", "AB4")       # the start-up sequence AB1, AB2, AB3, then AB4
    # examples/transport/scalar: a cosine bell carried round a periodic 1-D line of 20 elements (order 4 along x only: 5 x 1 x 1 nodes, CYCLIC
    # inlet/outlet, frozen uniform wind, BDF1), 40 steps
    make("transport-scalar", "BDF1", example="/root/reference/examples/transport/scalar")
    # examples/transport/wave2d: a Gaussian carried across a 2-D box with the BLENDED face value (0.6 central + 0.4 upwind by the sign of
    # the facet flux, field.h:3427-3437) instead of RUSANOV; also with UDS and CDS
    make("transport-wave2d", "BDF1", example="/root/reference/examples/transport/wave2d", block="simple")
    make("transport-wave2d-uds", "BDF1", example="/root/reference/examples/transport/wave2d", block="simple", edits={"convection_scheme": "UDS"})
    make("transport-wave2d-cds", "BDF1", example="/root/reference/examples/transport/wave2d", block="simple", edits={"convection_scheme": "CDS"})
    make_amr()
    # examples/transport/wave2d-amr-dg as shipped: UDS face values (also on the 2:1 faces), AB1, regrid every dump
    make_amr("transport-wave2d-amr", 40, 20, example="/root/reference/examples/transport/wave2d-amr-dg", scheme="AB1", block="simple")
    # the same on the cubed sphere: examples/atmo/advection-sphere-amr (Lauritzen's wind, BDF1, 2-D refinement that never splits the radial
    # axis) at 8 x 8 cells per panel, order 2, one 12-day period in 480 steps; dumps every 20 steps, regrids before step 1 and after dump 12
    make_amr("advection-sphere-amr", 480, 20, example="/root/reference/examples/atmo/advection-sphere-amr", divisions=(8, 8, 1),
             edits={"dt": 2160, "npx": 2, "npy": 2}, amr_step=12, scheme="BDF1")


def make(name, scheme, example=EX, block=None, edits=None):
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "convection", name)
    d = os.path.join(tempfile.mkdtemp(prefix="conv_golden_"), "advection-leveque")
    shutil.copytree(example, d)
    for f in os.listdir(d):
        os.chmod(os.path.join(d, f), 0o644)
    block = block or [f for f in os.listdir(d) if f != "controls" and not f.endswith((".txt", ".sh"))][0]
    m = subprocess.run([run_ref.ref_bin("mesh"), block, "-o", "grid_0.bin"], cwd=d, capture_output=True, text=True, timeout=600)
    assert m.returncode == 0, m.stdout[-1000:] + m.stderr[-1000:]
    os.chmod(os.path.join(d, "controls"), 0o644)
    ctl = open(os.path.join(d, "controls")).read()
    for k, v in (edits or {}).items():
        assert re.search(rf"(?m)^(\s*){k}\s+\S+", ctl), k
        ctl = re.sub(rf"(?m)^(\s*){k}\s+\S+", rf"\g<1>{k} {v}", ctl)
    ctl = re.sub(r"(?m)^(\s*)end_step\s+\d+", rf"\g<1>end_step {NSTEPS}", ctl)
    ctl = re.sub(r"(?m)^(\s*)write_interval\s+\d+", rf"\g<1>write_interval {NSTEPS}", ctl)
    ctl = re.sub(r"(?m)^(\s*)write_format\s+\w+", r"\g<1>write_format BINARY", ctl)
    ctl = re.sub(r"(?m)^(\s*)time_scheme\s+\w+", rf"\g<1>time_scheme {scheme}", ctl)
    ctl = re.sub(r"(?m)^\s*amr_step\s+\d+\s*\n", "", ctl)
    open(os.path.join(d, "controls"), "w").write(ctl)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for f in ("controls", "grid_0.bin", "U0.txt", "T0.txt"):
        shutil.copy(os.path.join(d, f), os.path.join(out, f))
        os.chmod(os.path.join(out, f), 0o644)
    exe = run_ref.ref_bin("convection")
    r = subprocess.run([exe, "./controls"], cwd=d, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    print("\n".join(l for l in r.stdout.splitlines() if "loss" in l)[-300:])
    T = refio.read_field_values(os.path.join(d, "T1"))[:, 0]
    U = refio.read_field_values(os.path.join(d, "U1"))
    np.savez_compressed(os.path.join(out, "expected.npz"), nsteps=NSTEPS, T=T, U=U)
    print(name, T.shape, U.shape, "T range", T.min(), T.max())


if __name__ == "__main__":
    main()


def make_amr(name="advection-leveque-amr", nsteps=40, interval=20, example=EX, divisions=None, edits=None, amr_step=1, scheme="AB2", block=None):
    """An AMR RUN of the convection app: the example exactly as it ships (AB2, amr_step 1, max_level 2, buffer_zone 2) for `nsteps` steps with
    a dump (and a regrid) every `interval`.  Kept: the case before the run, the cells of the grid of every regrid (centroid, volume), the
    scalar after the first `interval` steps (the same run stopped there) and at the end, with the node positions of their grids (oracle
    geometry of the reference's grids)."""
    from oracle.dg import Basis, Geometry
    from oracle.mesh import MeshTopo
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "convection", name)
    from oracle.euler import Params
    d = os.path.join(tempfile.mkdtemp(prefix="conv_amr_"), "case")
    shutil.copytree(example, d)
    for f in os.listdir(d):
        os.chmod(os.path.join(d, f), 0o644)
    block = block or [f for f in os.listdir(d) if f != "controls" and not f.endswith((".txt", ".sh"))][0]
    if divisions:
        blk = open(os.path.join(d, block)).read()
        blk, cnt = re.subn(r"(?m)^(8\{[^}]*\}\s+linear\s+)3\{\d+ \d+ \d+\}", r"\g<1>3{%d %d %d}" % divisions, blk)
        assert cnt == 6
        open(os.path.join(d, block), "w").write(blk)
    m = subprocess.run([run_ref.ref_bin("mesh"), block, "-o", "grid_0.bin"], cwd=d, capture_output=True, text=True, timeout=600)
    assert m.returncode == 0, m.stdout[-1000:] + m.stderr[-1000:]
    ctl = open(os.path.join(d, "controls")).read()
    ctl = re.sub(r"(?m)^(\s*)end_step\s+\d+", rf"\g<1>end_step {nsteps}", ctl)
    ctl = re.sub(r"(?m)^(\s*)write_interval\s+\d+", rf"\g<1>write_interval {interval}", ctl)
    if re.search(r"(?m)^\s*write_format", ctl):
        ctl = re.sub(r"(?m)^(\s*)write_format\s+\w+", r"\g<1>write_format BINARY", ctl)
    for k, v in (edits or {}).items():
        assert re.search(rf"(?m)^(\s*){k}\s+\S+", ctl), k
        ctl = re.sub(rf"(?m)^(\s*){k}\s+\S+", rf"\g<1>{k} {v}", ctl)
    assert re.search(r"(?m)^\s*amr_step\s+1\s*$", ctl) and re.search(rf"(?m)^\s*time_scheme\s+{scheme}\s*$", ctl)
    ctl = re.sub(r"(?m)^(\s*)amr_step\s+1", rf"\g<1>amr_step {amr_step}", ctl)
    open(os.path.join(d, "controls"), "w").write(ctl)
    params = Params.from_controls(refio.read_controls(os.path.join(d, "controls")))
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for f in ("controls", "grid_0.bin", "U0.txt", "T0.txt"):
        shutil.copy(os.path.join(d, f), os.path.join(out, f))
    for f in os.listdir(d):          # only the case files travel into the runs (a second block file must not be picked up)
        if f not in ("controls", "grid_0.bin", "U0.txt", "T0.txt"):
            os.remove(os.path.join(d, f))
    half = d + "_half"
    shutil.copytree(out, half)
    # the wind's period is end_step * dt, so the run stopped at the first dump keeps end_step and gets a kill switch instead: run the
    # full case twice is not possible either (the second regrid overwrites dump 1) -- so the half-way state comes from a run whose second
    # regrid is switched off by amr_step 2
    open(os.path.join(half, "controls"), "w").write(re.sub(r"(?m)^(\s*)amr_step\s+\d+", rf"\g<1>amr_step {2 * amr_step}", ctl))
    nop = [int(re.search(rf"(?m)^\s*{k}\s+(\d+)", ctl).group(1)) for k in ("npx", "npy", "npz")]
    keep = {}
    for tag, dd, gk, dump in (("half", half, 0, amr_step), ("end", d, amr_step, nsteps // interval)):
        r = subprocess.run([run_ref.ref_bin("convection"), "./controls"], cwd=dd, capture_output=True, text=True, timeout=1800)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        print(tag, [l for l in r.stdout.splitlines() if "Refining " in l])
        topo = MeshTopo(refio.read_grid(os.path.join(dd, f"grid_{gk}"), prefer_bin=True))
        topo.spherical, topo.sphere_radius, topo.sphere_height = params.is_spherical, params.sphere_radius, params.sphere_height
        topo.load()
        geo = Geometry(topo, Basis(nop))
        n = geo.gBCSfield
        keep[f"{tag}_CC"] = np.asarray(topo.CC)[: topo.nBCS]
        keep[f"{tag}_xyz"] = np.asarray(geo.cC)[:n]
        keep[f"{tag}_cV"] = np.asarray(geo.cV)[:n]
        keep[f"{tag}_T"] = refio.read_field_values(os.path.join(dd, f"T{dump}"))[:n, 0]
    # how far the reference's own -O2 and -O3 builds are apart at the end of this run (SURVEY finding 6): the device test allows 3 x that
    # where it exceeds 1e-11 (a deformational flow over a whole period amplifies rounding differences)
    fast = d + "_fast"
    shutil.copytree(out, fast)
    rf = subprocess.run([run_ref.ref_bin("convection", "fast"), "./controls"], cwd=fast, capture_output=True, text=True, timeout=1800,
                        env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert rf.returncode == 0, rf.stdout[-2000:] + rf.stderr[-2000:]
    Tf = refio.read_field_values(os.path.join(fast, f"T{nsteps // interval}"))[:, 0]
    if Tf.shape[0] >= keep["end_T"].shape[0]:
        Tf = Tf[: keep["end_T"].shape[0]]
        keep["spread_T"] = float(np.linalg.norm(Tf - keep["end_T"]) / np.linalg.norm(keep["end_T"]))
    else:
        keep["spread_T"] = float("nan")          # the two builds did not even end on the same grid
    print("   -O2 vs -O3 spread of T at the end:", keep["spread_T"])
    np.savez_compressed(os.path.join(out, "expected.npz"), nsteps=nsteps, interval=interval, amr_step=amr_step, NP=geo.gBCSfield // topo.nBCS, **keep)
    print(name, {k: getattr(v, "shape", v) for k, v in keep.items()})
