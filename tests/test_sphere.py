"""Cubed-sphere shells (general{is_spherical YES}) against fixtures made by the UNMODIFIED reference binaries
(tests/golden/make_sphere_golden.py -> tests/golden/sphere/<case>/expected.npz: the reference's geometry arrays after Mesh::LoadMesh and
its dump after nsteps steps):

  hydro-sphere      3-D shell, order 2, radial gravity + hydrostatic reference state, a warm blob placed with the great-circle initialiser,
                    no rho0 file (boundary cells of rho keep their set-up values)
  acoustic-sphere   one radial layer with both shells deleted (a 2-D surface), order 3, viscosity 30
  advection-sphere  scalar advection with Lauritzen's deformational wind re-evaluated every step

CPU: the oracle (ExtrudeMesh, curved-element geometry, radial node placement, spherical initialisers, radial gravity, Lauritzen winds)
against the reference arrays and dumps; the C++ host against the oracle, bit for bit.  GPU (-m gpu): the device run through the host
library against the reference's dump."""
import os
import shutil

import numpy as np
import pytest

from oracle import case as ocase
from tests.helpers import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sphere")
GEOM = (("cC", lambda g: g.cC), ("cV", lambda g: g.cV), ("Jinv", lambda g: g.Jinv), ("fN", lambda g: g.fN), ("fC", lambda g: g.fC),
        ("fI", lambda g: g.fI), ("gFN", lambda g: g.topo.FNv), ("gFC", lambda g: g.topo.FC), ("gCV", lambda g: g.topo.CV),
        ("gCC", lambda g: g.topo.CC), ("vertices", lambda g: g.topo.V))
TOL = 1e-11          # north_star: relative L2 against the reference after the run


def load(name):
    exp = np.load(os.path.join(GOLD, name, "expected.npz"))
    if name == "advection-sphere":
        return exp, ocase.load_convection_case(os.path.join(GOLD, name), exact_order=False)
    return exp, ocase.load_case(os.path.join(GOLD, name), exact_order=False)


@pytest.mark.parametrize("name", ["hydro-sphere", "acoustic-sphere", "advection-sphere"])
def test_oracle_sphere_geometry_is_bit_identical_to_the_reference(name):
    exp, orc = load(name)
    g = orc.g
    assert g.spherical
    for nm, get in GEOM:
        assert np.array_equal(np.asarray(get(g)).ravel(), exp[nm].ravel()), nm
    assert np.array_equal(g.FO, exp["FO"]) and np.array_equal(g.FN, exp["FN"])
    # the shell: the vertices on concentric spheres between the two radii, the volumes add up to the shell's (curved elements, not flat hexahedra)
    r = np.sqrt((g.topo.V ** 2).sum(axis=1))
    ri, ro = orc.p.sphere_radius, orc.p.sphere_radius + orc.p.sphere_height
    assert abs(r.min() - ri) < 1e-6 and abs(r.max() - ro) < 1e-6
    layers = np.unique(np.round(r, 3))
    assert len(layers) <= 3 and np.allclose(np.diff(layers), (ro - ri) / (len(layers) - 1))      # concentric shells, evenly spaced
    shell = 4.0 / 3.0 * np.pi * (ro ** 3 - ri ** 3)
    assert abs(g.topo.CV[: g.nBCS].sum() - shell) <= 2e-3 * shell


@pytest.mark.parametrize("name", ["hydro-sphere", "acoustic-sphere"])
def test_oracle_euler_on_the_sphere_matches_the_reference_binary(name):
    exp, orc = load(name)
    nb = orc.gB
    if name == "hydro-sphere":
        # gravity points to the centre, the geopotential is the height above the inner shell (euler.cpp:109-111)
        assert np.allclose(np.einsum("ij,ij->i", orc.gvec[:nb], orc.g.cC[:nb]) / np.linalg.norm(orc.g.cC[:nb], axis=1), -9.80606, rtol=1e-12)
        assert orc.gh[:nb].min() >= -9.80606 * orc.p.sphere_height * (1 + 1e-9) and orc.gh[:nb].max() <= 1e-6
    orc.run(int(exp["nsteps"]))
    for nm, a in (("rho", orc.rho), ("U", orc.U), ("T", orc.T), ("p", orc.pp)):
        assert rel_l2(a[:nb], exp[nm]) <= 1e-12, nm


def test_oracle_lauritzen_advection_matches_the_reference_binary():
    exp, orc = load("advection-sphere")
    nb = orc.gB
    # the start branch evaluates the wind at Iteration::get_step() * dt = dt and writes it to dump 0 (convection.cpp:96-102)
    assert np.array_equal(orc.wind(orc.p.dt, orc.end_step * orc.p.dt)[:nb], exp["U_start"])
    orc.run(int(exp["nsteps"]))
    assert np.array_equal(orc.U[:nb], exp["U"])
    assert rel_l2(orc.T[:nb], exp["T"]) <= 1e-14


@pytest.mark.parametrize("name", ["hydro-sphere", "acoustic-sphere", "advection-sphere"])
def test_host_sphere_geometry_and_setup_bit_equal_to_oracle(tmp_path, name):
    from nebulasem_b200 import host
    exp, orc = load(name)
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(GOLD, name), d)
    s = host.Solver.open_case(d)
    g = orc.g
    assert (s.nBCS, s.nCells, s.nFacets) == (g.nBCS, g.nCells, g.nFacets)
    for nm, ref in (("cC", g.cC), ("cV", g.cV), ("Jinv", g.Jinv), ("fN", g.fN), ("fC", g.fC), ("fI", g.fI), ("faceNormal", g.topo.FNv),
                    ("faceCenter", g.topo.FC)):
        assert np.array_equal(s.f64(nm), np.asarray(ref).ravel()), nm
    for nm, ref in (("FO", g.FO), ("FN", g.FN)):
        assert np.array_equal(s.u32(nm), np.asarray(ref, dtype=np.uint32).ravel()), nm
    rho, U, T, p = s.state()
    if name == "advection-sphere":
        assert np.array_equal(rho, orc.T) and np.array_equal(U, orc.U)          # the scalar lives in the rho slot
    else:
        assert np.array_equal(rho, orc.rho) and np.array_equal(U, orc.U) and np.array_equal(T, orc.T) and np.array_equal(p, orc.pp)
        assert np.array_equal(s.f64("rho_ref"), orc.rho_ref) and np.array_equal(s.f64("p_ref"), orc.p_ref)
        assert np.array_equal(s.f64("g"), orc.gvec.ravel())
    s.close()


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hydro-sphere", "acoustic-sphere", "advection-sphere"])
def test_device_run_on_the_sphere_matches_the_reference_binary(tmp_path, name):
    from nebulasem_b200 import host
    exp = np.load(os.path.join(GOLD, name, "expected.npz"))
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(GOLD, name), d)
    s = host.Solver.open_case(d)
    s.attach(0)
    n0 = s.launch_count
    s.step(int(exp["nsteps"]))
    s.download()
    rho, U, T, p = s.state()
    launches = s.launch_count - n0
    info = s.kernel_info
    nb = s.nBCS * s.NP
    if name == "advection-sphere":
        errs = {"T": rel_l2(rho[:nb], exp["T"]), "U": rel_l2(U[:nb], exp["U"])}
        spread = 0.0
    else:
        # the conserved variables, as in tests/test_gpu_parity.py: momentum against the scale rho * c0 and, self-relative, within 3 x the
        # distance between the reference's own -O2 and -O3 builds on this case (the fixture holds it) where that exceeds 1e-11
        T0, c0 = 300.0, np.sqrt(1004.67 / 715.5 * (1004.67 - 715.5) * 300.0)
        errs = {"rho": rel_l2(rho[:nb], exp["rho"]),
                "rhoTheta": rel_l2(rho[:nb] * (T[:nb] + T0), exp["rho"] * (exp["T"] + T0)),
                "rhoU_scaled": rel_l2(rho[:nb, None] * U[:nb], exp["rho"][:, None] * exp["U"], scale=np.linalg.norm(exp["rho"]) * c0),
                "p": rel_l2(p[:nb], exp["p"], scale=np.linalg.norm(exp["p"]) + 1e-9 * 101325.0 * np.sqrt(nb))}
        spread = float(exp["spread_rhoU_self"])
        errs_self = rel_l2(rho[:nb, None] * U[:nb], exp["rho"][:, None] * exp["U"])
        print("   rhoU self-relative", errs_self, "reference -O2 vs -O3:", spread)
    print(name, info, "launches", launches, "rel L2 vs the reference:", errs)
    s.close()
    assert launches >= 2 * int(exp["nsteps"])
    for k, e in errs.items():
        assert np.isfinite(e) and e <= TOL, (k, e)
    if name != "advection-sphere":
        assert errs_self <= max(TOL, 3.0 * spread), (errs_self, spread)
