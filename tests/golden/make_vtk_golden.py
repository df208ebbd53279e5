"""Generate tests/golden/vtk/: small cases run by the UNMODIFIED reference (oracle/_ref/parity/euler), converted by the reference's own
`prepare ./controls -vtk -start 1` (Vtk::write_vtk, src/vtk/vtk.cpp).

Run in the build container (where /root/reference exists and oracle/build_ref.sh has been run):
    python tests/golden/make_vtk_golden.py
Each fixture directory holds what the conversion reads (controls, grid_0.txt, rho/U/T/p1.bin) and what it wrote (grid1.vtk.gz), so that
the test can redo the conversion with `euler ./controls -vtk -start 1` and compare byte by byte.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cases, run_ref  # noqa: E402

FIXTURES = [
    ("bubble2d_n3_o3", "bubble2d", dict(n=3, order=3), 5, ""),                    # NPY = 1: quadrilateral sub-cells
    ("bubble3d_n2_o2", "bubble3d", dict(n=2, order=2), 4, ""),                    # hexahedral sub-cells
    ("hill3d_3x1x2_o2", "hill3d", dict(nx=3, ny=1, nz=2, order=2), 3,             # terrain-following nodes, no cellID block
     "vtk\n{\n    write_cell_value NO\n}\n"),
]


def reference_vtk(case_dir: str, index: int = 1) -> bytes:
    """The file the reference's prepare makes of dump <index> of case_dir."""
    out = subprocess.run([run_ref.ref_bin("prepare"), "./controls", "-vtk", "-start", str(index)], cwd=case_dir, capture_output=True,
                         text=True, timeout=300)
    if out.returncode != 0:
        raise RuntimeError(out.stdout[-1000:] + out.stderr[-1000:])
    with open(os.path.join(case_dir, f"grid{index}.vtk"), "rb") as f:
        return f.read()


def main():
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vtk")
    for fname, case, kw, nsteps, extra in FIXTURES:
        d = tempfile.mkdtemp(prefix="vtk_golden_")
        try:
            cases.CASES[case](**kw).write(d, nsteps)
            with open(os.path.join(d, "controls"), "a") as f:
                f.write(extra)
            run_ref.run_euler(d, variant="parity")
            vtk = reference_vtk(d)
            dst = os.path.join(out_dir, fname)
            os.makedirs(dst, exist_ok=True)
            for f in ("controls", "grid_0.txt", "rho1.bin", "U1.bin", "T1.bin", "p1.bin"):
                shutil.copy(os.path.join(d, f), os.path.join(dst, f))
            with open(os.path.join(dst, "grid1.vtk.gz"), "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as g:
                g.write(vtk)
            print(fname, len(vtk), "bytes of VTK")
        finally:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
