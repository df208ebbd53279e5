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


def make(name, scheme):
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "convection", name)
    d = os.path.join(tempfile.mkdtemp(prefix="conv_golden_"), "advection-leveque")
    shutil.copytree(EX, d)
    block = [f for f in os.listdir(d) if f != "controls" and not f.endswith((".txt", ".sh"))][0]
    m = subprocess.run([run_ref.ref_bin("mesh"), block, "-o", "grid_0.bin"], cwd=d, capture_output=True, text=True, timeout=600)
    assert m.returncode == 0, m.stdout[-1000:] + m.stderr[-1000:]
    os.chmod(os.path.join(d, "controls"), 0o644)
    ctl = open(os.path.join(d, "controls")).read()
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
