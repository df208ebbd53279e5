"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference binary (oracle/_ref/parity/euler).

Run in the build container (where /root/reference exists and oracle/build_ref.sh has been run):
    python tests/golden/make_golden.py
Each fixture = the reference's binary dumps rho/U/T/p after N steps of a case from oracle/cases.py, stored with the
case name, its keyword arguments and the step count.  The reference ships no golden vectors of its own
(tests/run_tests.sh only checks that examples run), so these dumps are what pins the oracle.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cases, run_ref  # noqa: E402

FIXTURES = [
    ("bubble2d_n10_o4_s100", "bubble2d", dict(n=10, order=4), 100),     # BASELINE configs[0]: examples/atmo/srtb, 100 steps
    ("bubble3d_n3_o4_s20", "bubble3d", dict(n=3, order=4), 20),
    ("bubble3d_n3_o2_s20", "bubble3d", dict(n=3, order=2), 20),
    ("vortex_n6_o3_s50", "vortex", dict(n=6, order=3), 50),
    ("vortex_n4_o6_s20", "vortex", dict(n=4, order=6), 20),
    ("hill3d_6x2x4_o3_s10", "hill3d", dict(nx=6, ny=2, nz=4, order=3), 10),
    # north_star: "after 100 RK steps" -- 3-D cases at the full step count (round 2)
    ("bubble3d_n3_o4_s100", "bubble3d", dict(n=3, order=4), 100),
    ("hill3d_6x2x4_o3_s100", "hill3d", dict(nx=6, ny=2, nz=4, order=3), 100),
    ("vortex_n4_o4_s100", "vortex", dict(n=4, order=4), 100),
]


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for fname, case, kw, nsteps in FIXTURES:
        d = tempfile.mkdtemp(prefix="golden_")
        try:
            cases.CASES[case](**kw).write(d, nsteps)
            run_ref.run_euler(d, variant="parity")
            dump = run_ref.read_dump(d, 1)
            # the reference against itself: the same sources built with its release flags (-O3, FMA contraction; oracle/_ref/fast) on the
            # same case.  SURVEY finding 6: the self-relative momentum error of ANY faithful implementation is only meaningful against this
            # spread (p - p_ref cancels in the early bubble momentum), so the parity tests assert rhoU_self <= 3 x spread
            d2 = tempfile.mkdtemp(prefix="golden_fast_")
            try:
                cases.CASES[case](**kw).write(d2, nsteps)
                run_ref.run_euler(d2, variant="fast", threads=1)
                fast = run_ref.read_dump(d2, 1)
            finally:
                shutil.rmtree(d2, ignore_errors=True)
            T0 = 300.0
            ctl = open(os.path.join(d, "controls")).read().split()
            if "T0" in ctl:
                T0 = float(ctl[ctl.index("T0") + 1])
            rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
            spread = dict(rho=rel(fast["rho"], dump["rho"]), rhoU_self=rel(fast["rho"][:, None] * fast["U"], dump["rho"][:, None] * dump["U"]),
                          rhoTheta=rel(fast["rho"] * (fast["T"] + T0), dump["rho"] * (dump["T"] + T0)))
            np.savez_compressed(os.path.join(out_dir, fname + ".npz"), case=case, kwargs=repr(kw), nsteps=nsteps,
                                rho=dump["rho"], U=dump["U"], T=dump["T"], p=dump["p"],
                                spread_rho=spread["rho"], spread_rhoU_self=spread["rhoU_self"], spread_rhoTheta=spread["rhoTheta"])
            print(fname, {k: v.shape for k, v in dump.items()}, "O2-vs-O3 spread:", spread)
        finally:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
