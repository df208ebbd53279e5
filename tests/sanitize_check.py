"""compute-sanitizer target (no pytest, no torch): a few steps of every kernel family on small meshes.
    compute-sanitizer --tool memcheck python tests/sanitize_check.py
Covers: v4 sweeps (conforming 3-D), v4 + mortar kernels (3-D non-conforming), v1 + mortar kernels (2-D non-conforming and 3-D with
NSEM_MORTAR_V1=1), v1 2-D, the boundary/ghost-trace kernels, upload/download, the pipelined transfers and the AMR field-transfer kernels
(copy / merge / split + restart pass), the run schedule of sweep A, the scalar-advection mode (winds, AB update, face-value schemes, 1-D)
and the cubed-sphere cases.  NSEM_SANITIZE_ONLY=amr runs the AMR group alone."""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nebulasem_b200 import host  # noqa: E402


def run(s, nsteps, tag):
    s.attach(0)
    info = s.kernel_info
    s.step(nsteps)
    s.download()
    rho, U, T, p = s.state()
    ok = np.isfinite(rho).all() and np.isfinite(U).all() and np.isfinite(T).all()
    print("SANITIZE", tag, "|", info, "| finite:", bool(ok), flush=True)
    s.close()
    return ok


def run_amr():
    """both regrids of the 3-D golden case (copy, split, merge), restart pass, a few steps on the result; pipelined transfers"""
    gold = os.path.join(ROOT, "tests", "golden", "refine_field")
    ok = True
    for k in (1, 2):
        g = np.load(os.path.join(gold, f"3d_o2_pass{k}.npz"))
        old = host.Solver.open_case(os.path.join(gold, "3d_o2", f"stage{k - 1}"))
        new = host.Solver.open_case(os.path.join(gold, "3d_o2", f"stage{k}"))
        old.attach(0)
        new.attach(0)
        new.adopt_refined_state(old, g["refineMap"], g["coarseMap"], g["cellMap"], restart=True)
        new.step(2)
        new.upload_async()
        new.step(1)
        new.download_async()
        new.sync()
        fin = all(np.isfinite(x).all() for x in new.state_out())
        print("SANITIZE", f"amr transfer pass {k}", "|", new.kernel_info, "| finite:", bool(fin), flush=True)
        ok &= fin
        old.close()
        new.close()
    return ok


def main():
    ok = True
    if os.environ.get("NSEM_SANITIZE_ONLY") == "amr":
        ok = run_amr()
        print("SANITIZE_DONE", "OK" if ok else "NOT FINITE", flush=True)
        sys.exit(0 if ok else 1)
    ok &= run(host.Solver.synthetic("bubble3d", 3, 3, 3, 4), 3, "bubble3d 3^3 order 4")
    ok &= run(host.Solver.synthetic("bubble2d", 4, 1, 4, 4), 3, "bubble2d 4x4 order 4")
    ok &= run(host.Solver.synthetic("hill3d", 6, 2, 4, 3), 3, "hill3d 6x2x4 order 3")
    with tempfile.TemporaryDirectory() as d:
        for fixture, env in (("srtb3d_amr", "0"), ("srtb3d_amr", "1"), ("srtb_amr", "0")):
            os.environ["NSEM_MORTAR_V1"] = env
            c = os.path.join(d, fixture + env)
            shutil.copytree(os.path.join(ROOT, "tests", "golden", fixture), c)
            ok &= run(host.Solver.open_case(c), 3, f"{fixture} NSEM_MORTAR_V1={env}")
    ok &= run_amr()
    # round 2: runs of consecutive elements per CTA (RunIter with run > 1 needs >= 4 runs per CTA: 18^3 elements), S kept with the state,
    # and the scalar-advection mode of the sweeps with the wind re-evaluated on the device (the reference's own LeVeque example)
    ok &= run(host.Solver.synthetic("bubble3d", 18, 18, 18, 2), 2, "bubble3d 18^3 order 2 (runs of elements per CTA)")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "leveque")
        shutil.copytree(os.path.join(ROOT, "tests", "golden", "convection", "advection-leveque"), c)
        ok &= run(host.Solver.open_case(c), 3, "convection advection-leveque")
    # later in round 2: cubed-sphere cases (per-node gravity, the UNLISTED continuation of rho's boundary cells, the Lauritzen wind kernel),
    # the Adams-Bashforth update, the 1-D instantiation and the upwind / blended face values of the plain-load sweep
    with tempfile.TemporaryDirectory() as d:
        for sub, name in (("sphere", "hydro-sphere"), ("sphere", "advection-sphere"), ("convection", "advection-leveque-ab4"),
                          ("convection", "transport-scalar"), ("convection", "transport-wave2d")):
            c = os.path.join(d, name)
            shutil.copytree(os.path.join(ROOT, "tests", "golden", sub, name), c)
            ok &= run(host.Solver.open_case(c), 5, f"{sub}/{name}")
    print("SANITIZE_DONE", "OK" if ok else "NOT FINITE", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
