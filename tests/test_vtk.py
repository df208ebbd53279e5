"""VTK output: `euler ./controls -vtk` and EulerSolver::write_vtk against the files the reference's own
`prepare ./controls -vtk` wrote (Vtk::write_vtk, src/vtk/vtk.cpp:125-286) -- byte for byte.  Host only, no GPU."""
import glob
import gzip
import os
import shutil
import subprocess

import pytest

from nebulasem_b200 import build
from oracle import cases as ocases
from oracle import run_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vtk")
FIXTURES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*")) if os.path.isdir(p))


def convert(case_dir, *args):
    out = subprocess.run([build.EULER_BIN, "./controls", "-vtk", *args], cwd=case_dir, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-500:] + out.stderr[-500:]
    return out.stdout


def test_fixtures_present():
    assert FIXTURES == ["bubble2d_n3_o3", "bubble3d_n2_o2", "hill3d_3x1x2_o2"]


@pytest.mark.parametrize("name", FIXTURES)
def test_vtk_byte_identical_to_reference_fixture(tmp_path, name):
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(GOLD, name), d)
    with gzip.open(os.path.join(d, "grid1.vtk.gz"), "rb") as f:
        want = f.read()
    assert "Converting result to VTK format." in convert(d, "-start", "1")
    got = open(os.path.join(d, "grid1.vtk"), "rb").read()
    assert got == want
    if name.startswith("hill3d"):
        assert b"CELL_DATA" not in got                  # vtk{write_cell_value NO}
    else:
        assert b"cellID  1 " in got


def test_solver_object_writes_the_same_file(tmp_path):
    """write_vtk of an opened case (the call the run loop makes after a download with NSEM_VTK=1) == the command-line conversion."""
    from nebulasem_b200 import host
    d = str(tmp_path / "case")
    shutil.copytree(os.path.join(GOLD, "bubble2d_n3_o3"), d)
    with gzip.open(os.path.join(d, "grid1.vtk.gz"), "rb") as f:
        want = f.read()
    s = host.Solver.open_case(d, 1)          # grid_0 is the newest grid <= dump 1 (findLastRefinedGrid, field.cpp:79-91)
    s.write_vtk(7)
    s.close()
    # open_case runs the set-up (p from rho, euler.cpp:150-162), so p may differ in the last printed digit: compare everything but p
    got = open(os.path.join(d, "grid7.vtk"), "rb").read()
    cut = lambda b: b[:b.index(b"\np 1 ")] + b[b.index(b"\nrho 1 "):]
    assert cut(got) == cut(want)


def test_range_of_dumps_and_field_selection(tmp_path):
    """-start i -stop j converts dumps i..j-1; prepare{fields} selects and orders the fields (scalars first, then vectors)."""
    d = str(tmp_path / "case")
    shutil.copytree(os.path.join(GOLD, "bubble3d_n2_o2"), d)
    for f in ("rho", "U", "T", "p"):
        shutil.copy(os.path.join(d, f + "1.bin"), os.path.join(d, f + "2.bin"))
    ctl = open(os.path.join(d, "controls")).read().replace("fields 4 { U T p rho }", "fields 2 { U rho }")
    assert "fields 2 { U rho }" in ctl
    open(os.path.join(d, "controls"), "w").write(ctl)
    convert(d, "-start", "1", "-stop", "3")
    a, b = (open(os.path.join(d, f"grid{k}.vtk"), "rb").read() for k in (1, 2))
    assert a == b and b"FIELD attributes 2\nrho 1 216 double" in a and b"\nU 3 216 double" in a and b"\nT 1 " not in a
    assert a.index(b"rho 1 216") < a.index(b"U 3 216")


def test_unknown_option_is_an_error(tmp_path):
    d = str(tmp_path / "case")
    shutil.copytree(os.path.join(GOLD, "bubble3d_n2_o2"), d)
    out = subprocess.run([build.EULER_BIN, "./controls", "-vtx"], cwd=d, capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "unknown option" in out.stderr


@pytest.mark.skipif(not run_ref.have_ref("parity"), reason="oracle/_ref/parity not built")
@pytest.mark.parametrize("case,kw,nsteps", [
    ("vortex", dict(n=3, order=4), 3),
    ("bubble2d", dict(n=4, order=2), 2),
])
def test_vtk_byte_identical_to_reference_live(tmp_path, case, kw, nsteps):
    """Fresh cases through the reference's euler + prepare and through this repo's converter."""
    d = str(tmp_path / case)
    ocases.CASES[case](**kw).write(d, nsteps)
    run_ref.run_euler(d, variant="parity")
    out = subprocess.run([run_ref.ref_bin("prepare"), "./controls", "-vtk", "-start", "1"], cwd=d, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    os.rename(os.path.join(d, "grid1.vtk"), os.path.join(d, "ref1.vtk"))
    convert(d, "-start", "1")
    assert open(os.path.join(d, "grid1.vtk"), "rb").read() == open(os.path.join(d, "ref1.vtk"), "rb").read()


@pytest.mark.gpu
def test_binary_resumes_from_the_reference_dump_and_writes_vtk_on_dumps(tmp_path):
    """`start_step 4` on the reference's dump 1 (fixture): the binary runs steps 5..8 on the GPU, writes dump 2 and, with NSEM_VTK=1,
    grid2.vtk from the downloaded state.  Dump 2 against the oracle set up on dump 1's fields (which is bit-identical to the reference
    resumed the same way, tests/test_oracle_vs_reference.py); the VTK against the conversion of that dump."""
    import numpy as np

    from oracle import case as ocase
    from oracle import refio
    from tests.helpers import conserved_errors
    d = str(tmp_path / "resumed")
    shutil.copytree(os.path.join(GOLD, "bubble3d_n2_o2"), d)
    d0 = str(tmp_path / "as_step0")
    os.makedirs(d0)
    for f in ("rho", "U", "T", "p"):
        shutil.copy(os.path.join(d, f + "1.bin"), os.path.join(d0, f + "0.bin"))
    shutil.copy(os.path.join(d, "grid_0.txt"), d0)
    shutil.copy(os.path.join(d, "controls"), d0)
    orc = ocase.load_case(d0, exact_order=False)
    orc.run(4)
    ctl = open(os.path.join(d, "controls")).read()
    open(os.path.join(d, "controls"), "w").write(ctl.replace("start_step 0", "start_step 4").replace("end_step 4", "end_step 8"))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([build.EULER_BIN, "./controls"], cwd=d, env=dict(env, NSEM_VTK="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-800:]
    # the reference's diagnostics lines (euler.cpp:276-281) at the dump
    assert "Courant number: Max: " in out.stdout and "Mass loss: " in out.stdout and "Energy loss " in out.stdout, out.stdout[-800:]
    rho, U, T = (refio.read_field_values(os.path.join(d, f + "2")) for f in ("rho", "U", "T"))
    err = conserved_errors(orc, rho[:, 0], U, T[:, 0])
    print(err)
    assert err["rho"] <= 1e-11 and err["rhoTheta"] <= 1e-11 and err["rhoU_scaled"] <= 1e-11
    on_dump = open(os.path.join(d, "grid2.vtk"), "rb").read()
    os.rename(os.path.join(d, "grid2.vtk"), os.path.join(d, "on_dump.vtk"))
    convert(d, "-start", "2")
    converted = open(os.path.join(d, "grid2.vtk"), "rb").read()
    cut = lambda b: b[:b.index(b"\np 1 ")] + b[b.index(b"\nrho 1 "):]     # the conversion's set-up recomputes p from rho
    assert cut(on_dump) == cut(converted) and b"POINT_DATA 216" in on_dump


@pytest.mark.skipif(not run_ref.have_ref("parity"), reason="oracle/_ref/parity not built")
@pytest.mark.parametrize("fixture", ["srtb_amr", "srtb3d_amr"])
def test_vtk_of_a_non_conforming_grid_byte_identical_live(tmp_path, fixture):
    """The grids the reference's own regrid wrote (2:1 hanging nodes, merged sides): same file as the reference's prepare."""
    d = str(tmp_path / fixture)
    shutil.copytree(os.path.join(os.path.dirname(GOLD), fixture), d)
    out = subprocess.run([run_ref.ref_bin("prepare"), "./controls", "-vtk", "-start", "0"], cwd=d, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    os.rename(os.path.join(d, "grid0.vtk"), os.path.join(d, "ref0.vtk"))
    convert(d, "-start", "0")
    assert open(os.path.join(d, "grid0.vtk"), "rb").read() == open(os.path.join(d, "ref0.vtk"), "rb").read()


@pytest.mark.skipif(not run_ref.have_ref("parity"), reason="oracle/_ref/parity not built")
@pytest.mark.parametrize("sub,name,exe", [
    ("sphere", "hydro-sphere", "euler"),                       # 3-D shell; no rho0 file: a name without a file is not converted
    ("sphere", "acoustic-sphere", "euler"),                    # 2-D surface on the sphere
    ("sphere", "acoustic-sphere-regridded", "euler"),          # the reference's own regridded sphere (2:1 faces)
    ("sphere", "advection-sphere", "convection"),              # the convection app's fields (T in the rho slot) on the sphere
    ("convection", "advection-leveque", "convection"),
    ("convection", "transport-scalar", "convection"),          # 1-D: the line cells of Vtk::write_vtk
    ("convection", "transport-wave2d", "convection"),
])
def test_vtk_of_sphere_and_convection_cases_byte_identical_live(tmp_path, sub, name, exe):
    """`<app> ./controls -vtk` against `prepare ./controls -vtk` of the reference on the round-2 fixtures.  Prepare::convertVTK loads the mesh with
    remove_empty = false (prepare.cpp:12): a 2-D cell keeps its two empty faces and its node placement takes its corners from them -- on the
    sphere, where the placement's radial rescale is not symmetric in the element axes, the files of `prepare -vtk` hold nodes a metre away
    from the solver's own; the converter loads the same way (EulerSolver::vtk_mode)."""
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(os.path.dirname(GOLD), sub, name), d)
    out = subprocess.run([run_ref.ref_bin("prepare"), "./controls", "-vtk", "-start", "0"], cwd=d, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    os.rename(os.path.join(d, "grid0.vtk"), os.path.join(d, "ref0.vtk"))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([os.path.join(os.path.dirname(build.EULER_BIN), exe), "./controls", "-vtk", "-start", "0"], cwd=d, env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert open(os.path.join(d, "grid0.vtk"), "rb").read() == open(os.path.join(d, "ref0.vtk"), "rb").read()
