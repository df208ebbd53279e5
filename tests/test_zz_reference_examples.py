"""The reference's OWN euler examples (every non-spherical `solver euler` directory under examples/: isentropic, atmo/{ctbs, dc, lrtb,
srtb, srtb-3d, srtb-amr, srtb-amr-hill, srtb-amr-zaxis, srtb-curved, srtb-inclined}), as fixtures made by
tests/golden/make_examples_golden.py: the example's case files, the grid the reference's mesher made of its block file, and the dump of the
unmodified reference binary after 3 steps.

CPU: the numpy oracle is bit-identical to the reference on every one of them; the C++ host reads the same files and arrives at the same
geometry and set-up state, bit for bit.  GPU: the CUDA path from those files against the reference's dump (<= 1e-11, north_star).
(Named test_zz_* so that the GPU part runs after the core parity tests.)"""
import glob
import os
import shutil

import numpy as np
import pytest

from oracle import case as ocase
from tests.helpers import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "examples")
NAMES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*")) if os.path.isdir(p))
TOL = 1e-11


def test_all_example_fixtures_present():
    assert NAMES == ["ctbs", "dc", "isentropic", "lrtb", "srtb", "srtb-3d", "srtb-amr", "srtb-amr-hill", "srtb-amr-zaxis", "srtb-curved",
                     "srtb-inclined"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_and_host_setup_on_the_reference_examples(tmp_path, name):
    from nebulasem_b200 import host
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(GOLD, name), d)
    exp = np.load(os.path.join(d, "expected.npz"))
    orc = ocase.load_case(d, exact_order=True)
    s = host.Solver.open_case(d)
    g = orc.g
    assert (s.nBCS, s.nCells, s.nFacets) == (g.nBCS, g.nCells, g.nFacets)
    for nm, ref in (("cC", g.cC), ("cV", g.cV), ("Jinv", g.Jinv), ("fN", g.fN), ("fC", g.fC), ("fI", g.fI)):
        assert np.array_equal(s.f64(nm), np.asarray(ref).ravel()), nm
    for nm, ref in (("FO", g.FO), ("FN", g.FN), ("allFaces", g.allFaces)):
        assert np.array_equal(s.u32(nm), np.asarray(ref, dtype=np.uint32).ravel()), nm
    rho, U, T, p = s.state()
    assert np.array_equal(rho, orc.rho) and np.array_equal(U, orc.U) and np.array_equal(T, orc.T) and np.array_equal(p, orc.pp)
    assert np.array_equal(s.f64("rho_ref"), orc.rho_ref) and np.array_equal(s.f64("p_ref"), orc.p_ref)
    s.close()
    orc.run(int(exp["nsteps"]))
    nb = orc.gB
    assert np.array_equal(orc.rho[:nb], exp["rho"]) and np.array_equal(orc.U[:nb], exp["U"])
    assert np.array_equal(orc.T[:nb], exp["T"]) and np.array_equal(orc.pp[:nb], exp["p"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_runs_the_reference_examples(tmp_path, name):
    from nebulasem_b200 import host
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(GOLD, name), d)
    exp = np.load(os.path.join(d, "expected.npz"))
    s = host.Solver.open_case(d)
    T0, cp, cv = s.params["T0"], s.params["cp"], s.params["cv"]
    s.attach(0)
    s.step(int(exp["nsteps"]))
    s.download()
    rho, U, T, p = s.state()
    nb = s.gBCSfield
    s.close()
    c0 = np.sqrt(cp / cv * (cp - cv) * T0)
    err = dict(rho=rel_l2(rho[:nb], exp["rho"]),
               rhoTheta=rel_l2(rho[:nb] * (T[:nb] + T0), exp["rho"] * (exp["T"] + T0)),
               rhoU_scaled=rel_l2(rho[:nb, None] * U[:nb], exp["rho"][:, None] * exp["U"], scale=np.linalg.norm(exp["rho"]) * c0))
    print(name, err)
    assert err["rho"] <= TOL and err["rhoTheta"] <= TOL and err["rhoU_scaled"] <= TOL, err


@pytest.mark.parametrize("name,cells", [("srtb-amr", (100, 196)), ("srtb-3d", (216, 608)), ("srtb-amr-hill", (484, 802)),
                                        ("isentropic", (256, 436))])
def test_initial_regrid_of_the_amr_examples_matches_the_reference(tmp_path, monkeypatch, name, cells):
    """The examples that ship with `amr_step 1`: the reference regrids before step 1 (tagging by the example's refinement{} block, then
    MeshObject::refineMesh).  The in-memory regrid (amr.cpp) tags the same cells and places the new vertices where the reference does --
    also on the terrain-following srtb-amr-hill grid, whose face centres are not vertex averages (calcFaceCenter, mesh.cpp:1166-1188), and
    on the isentropic vortex with its CYCLIC patches: same cells (centroid, volume to 1e-12), same facet and mortar-face counts."""
    from nebulasem_b200 import host
    monkeypatch.setenv("NSEM_AMR", "1")                  # the fixture's controls have amr_step removed; keep the forest and refinement{}
    d = str(tmp_path / name)
    shutil.copytree(os.path.join(GOLD, name), d)
    exp = np.load(os.path.join(d, "initial_regrid.npz"))
    s = host.Solver.open_case(d)
    assert s.nBCS == cells[0]
    s.regrid()
    assert s.nBCS == cells[1] == len(exp["CV"]) and s.nFacets == int(exp["n_facets"])
    assert int((s.u32("faceMortar") > 0).sum()) == int(exp["n_mortar"])
    n = s.nBCS
    mine = np.concatenate([s.f64("gCC")[:3 * n].reshape(n, 3), s.f64("gCV")[:n, None]], axis=1)
    ref = np.concatenate([exp["CC"], exp["CV"][:, None]], axis=1)
    s.close()
    order = lambda a: a[np.lexsort(np.round(a, 6).T[::-1])]
    mine, ref = order(mine), order(ref)
    scale = np.array([np.abs(ref[:, :3]).max()] * 3 + [np.abs(ref[:, 3]).max()])      # one length scale for the coordinates (z may be 0 everywhere)
    err = np.abs(mine - ref).max(axis=0) / scale
    print(name, err)
    assert np.all(err <= 1e-12), err


def _corner_blob(s, lo, hi):
    x = s.f64("cC").reshape(-1, 3)
    r = np.linalg.norm((x - np.array([lo[0], lo[1], 0.0]))[:, :2], axis=1) / (0.3 * (hi[0] - lo[0]))
    return np.where(r < 1, 0.25 * (1 + np.cos(np.pi * r)), 0.0)


def test_periodic_faces_pair_node_by_node_after_a_regrid(tmp_path, monkeypatch):
    """applyExplicitBCs pairs slot n of face j of a CYCLIC patch with slot n of face j of the neighbor patch (field.h:2683-2687:
    cF[FN[k]] = cF[FO[fi*NPF + n]]).  After two in-memory regrids of examples/isentropic that pairing still joins nodes that are periodic
    images of each other: same owner-node coordinates up to the translation between the patches, slot by slot."""
    from nebulasem_b200 import host
    import re
    monkeypatch.setenv("NSEM_AMR", "1")
    d = str(tmp_path / "isentropic")
    shutil.copytree(os.path.join(GOLD, "isentropic"), d)
    s = host.Solver.open_case(d)
    s.enable_amr(direction=(0, 0, 1), field="T", field_min=0.15, field_max=0.4, max_level=2, buffer_zone=1)
    pairs = re.findall(r"(\w+)\s*\{\s*type\s+CYCLIC\s+neighbor\s+(\w+)", open(os.path.join(d, "U0.txt")).read())
    x = s.f64("cC").reshape(-1, 3)
    lo, hi = x.min(axis=0), x.max(axis=0)
    for cycle in range(2):
        s.set_state(T=_corner_blob(s, lo, hi))
        s.regrid()
        cC, FO = s.f64("cC").reshape(-1, 3), s.u32("FO")
        npf = len(FO) // s.nFacets
        for a, b in pairs:
            fa, fb = s.patch_faces(a), s.patch_faces(b)
            assert len(fa) == len(fb) > 16
            ka = FO.reshape(-1, npf)[fa]
            kb = FO.reshape(-1, npf)[fb]
            ok = ka < len(cC)
            assert np.array_equal(ok, kb < len(cC))
            shift = cC[kb[ok]] - cC[ka[ok]]
            assert np.abs(shift - shift[0]).max() <= 1e-9 * np.abs(hi - lo).max(), (cycle, a, b)
    s.close()


@pytest.mark.gpu
def test_device_amr_across_periodic_patches_matches_the_oracle(tmp_path, monkeypatch):
    """examples/isentropic on one B200 with a blob on the corner of the periodic domain: two regrids (to level 2) refine cells on the CYCLIC
    patches on all four periodic images and the state stays on the device (nsem_refine_state).  The transfer conserves mass; the 10 steps
    after each regrid are compared with the oracle stepping the SAME regridded grid from the SAME state, ghost cells of the re-paired
    periodic faces included (<= 1e-11).  Mass itself is NOT conserved across CYCLIC patches by the reference's scheme -- boundary faces
    interpolate with fI = 0 (field.cpp:257-270), so the two images of a periodic face see different central fluxes; the oracle on the
    unrefined example drifts by 3e-7 of the volume in 10 steps -- so the drift is compared with the oracle's, not with zero."""
    from nebulasem_b200 import host
    monkeypatch.setenv("NSEM_AMR", "1")
    d = str(tmp_path / "isentropic")
    shutil.copytree(os.path.join(GOLD, "isentropic"), d)
    s = host.Solver.open_case(d)
    s.enable_amr(direction=(0, 0, 1), field="T", field_min=0.15, field_max=0.4, max_level=2, buffer_zone=1)
    s.attach(0)
    x = s.f64("cC").reshape(-1, 3)
    lo, hi = x.min(axis=0), x.max(axis=0)

    def mass(rho):
        n = s.gBCSfield
        return float((rho[:n] * s.f64("cV")[:n]).sum())

    s.set_state(T=_corner_blob(s, lo, hi))
    s.upload()
    volume = float(s.f64("cV")[:s.gBCSfield].sum())
    s.download()
    m0 = mass(s.state()[0])
    n0 = s.nBCS
    for cycle in range(2):
        s.regrid()
        s.download()
        rho, U, T, p = [a.copy() for a in s.state()]
        assert abs(mass(rho) - m0) <= 1e-11 * volume, cycle                   # refineField's mass fix (field.h:1990-2013)
        # the oracle on the regridded grid (written like a dump of an AMR run), continued from the device's state
        s.write_amr_grid(10 + cycle)
        s.write(10 + cycle)
        orc = ocase.load_case(d, exact_order=False, step=10 + cycle)
        assert orc.g.nBCS == s.nBCS and np.array_equal(orc.g.cC.ravel(), s.f64("cC"))
        orc.rho, orc.U, orc.T, orc.pp = rho.copy(), U.copy(), T.copy(), p.copy()
        orc.run(10)
        s.step(10)
        s.download()
        rho, U, T, p = s.state()
        T0, c0 = orc.p.T0, np.sqrt(orc.gamma * orc.R * orc.p.T0)
        # every live entry: the real nodes and the ghost nodes a face refers to (a ghost cell has NP slots in the reference's layout and
        # uses the face's few; the others hold 0 and turn into 0/0 in the oracle's update)
        FN = s.u32("FN")
        live = np.zeros(len(rho), bool)
        live[:s.gBCSfield] = True
        live[FN[FN < len(rho)]] = True
        gh = live.copy()
        gh[:s.gBCSfield] = False
        assert gh.sum() >= 5 * (len(s.patch_faces("inx")) + len(s.patch_faces("outy")))
        assert rel_l2(rho[live], orc.rho[live]) <= TOL and rel_l2((rho * (T + T0))[live], (orc.rho * (orc.T + T0))[live]) <= TOL, cycle
        assert rel_l2((rho[:, None] * U)[live], (orc.rho[:, None] * orc.U)[live], scale=np.linalg.norm(orc.rho[live]) * c0) <= TOL, cycle
        assert rel_l2(rho[gh], orc.rho[gh]) <= TOL and rel_l2(T[gh] + T0, orc.T[gh] + T0) <= TOL, cycle
        assert rel_l2(U[gh], orc.U[gh], scale=np.sqrt(gh.sum()) * c0) <= TOL, cycle
        drift, drift_orc = mass(rho) - m0, mass(orc.rho) - m0
        assert abs(drift - drift_orc) <= 1e-11 * volume, (cycle, drift, drift_orc)
        m0 = mass(rho)
    assert all(np.isfinite(v).all() for v in s.state())
    lv, own = s.cell_levels(), s.u32("faceOwner")
    assert s.nBCS > n0 and lv.max() >= 1 and lv[own[s.patch_faces("inx")]].max() >= 1 and lv[own[s.patch_faces("outy")]].max() >= 1
    s.close()


def test_initial_regrid_of_the_zaxis_example_tags_what_the_reference_tags(tmp_path, monkeypatch):
    """examples/atmo/srtb-amr-zaxis (x-z plane, refinement{direction 0 1 0}): the reference binary tags 24 cells ("Refining 24 Coarsening 0")
    and then crashes inside its own regrid (both builds of oracle/_ref, so there is no grid to compare with); the in-memory regrid
    refines 24 cells in the plane: 100 - 24 + 4 * 24 = 172 cells, one level, 2:1 across faces."""
    from nebulasem_b200 import host
    monkeypatch.setenv("NSEM_AMR", "1")
    d = str(tmp_path / "zaxis")
    shutil.copytree(os.path.join(GOLD, "srtb-amr-zaxis"), d)
    s = host.Solver.open_case(d)
    assert s.nBCS == 100
    s.regrid()
    lv = s.cell_levels()
    assert s.nBCS == 172 and int((lv == 1).sum()) == 96 and lv.max() == 1
    assert abs(s.f64("gCV")[:s.nBCS].sum() - 1000.0 * 100.0 * 1000.0) <= 1e-6          # the 1 km x 100 m x 1 km slab
    s.close()
