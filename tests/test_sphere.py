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


def test_partitions_of_the_shell_are_projected_like_the_whole_mesh(tmp_path):
    """ExtrudeMesh scales by the extremes of the vertices it sees (mesh.cpp:735-741), and a part of a two-layer shell need not reach both
    cubes: every part projects with the whole grid's extremes, so its elements are the single-partition ones bit for bit (set-up included)."""
    from nebulasem_b200 import host
    d = str(tmp_path / "hydro-sphere")
    shutil.copytree(os.path.join(GOLD, "hydro-sphere"), d)
    g = host.Solver.open_case(d)
    NP = g.NP
    cC_g = g.f64("cC").reshape(-1, 3)[: g.gBCSfield].reshape(g.nBCS, -1)
    J_g = g.f64("Jinv")[: g.gBCSfield * 9].reshape(g.nBCS, -1)
    rho_g, U_g, T_g, p_g = [a[: g.gBCSfield].reshape(g.nBCS, -1) for a in g.state()]
    gv_g = g.f64("g").reshape(-1, 3)[: g.gBCSfield].reshape(g.nBCS, -1)
    seen = np.zeros(g.nBCS, dtype=int)
    nparts, one_shell_only = 5, 0
    for r in range(nparts):
        p = host.Solver.open_case(d, 0, r, nparts)
        cg = p.u32("cellGlobal")
        seen[cg] += 1
        rad = np.linalg.norm(p.f64("cC").reshape(-1, 3)[: p.gBCSfield], axis=1)
        one_shell_only += int(rad.max() - rad.min() < 0.75 * 10000.0)
        assert np.array_equal(p.f64("cC").reshape(-1, 3)[: p.gBCSfield].reshape(p.nBCS, -1), cC_g[cg])
        assert np.array_equal(p.f64("Jinv")[: p.gBCSfield * 9].reshape(p.nBCS, -1), J_g[cg])
        rho, U, T, pp = [a[: p.gBCSfield].reshape(p.nBCS, -1) for a in p.state()]
        assert np.array_equal(rho, rho_g[cg]) and np.array_equal(T, T_g[cg]) and np.array_equal(pp, p_g[cg])
        assert np.array_equal(p.f64("g").reshape(-1, 3)[: p.gBCSfield].reshape(p.nBCS, -1), gv_g[cg])
        p.close()
    assert (seen == 1).all()
    g.close()


def test_regrid_of_the_shell_tags_and_splits_like_the_reference(tmp_path):
    """The initial regrid of the reduced examples/atmo/acoustic-sphere-amr-dg (tests/golden/amr_run/, made by the reference's euler binary):
    Prepare::refineMesh weighs the indicator with the node volumes of the UN-projected load of the mesh (field.cpp:606-645) and never
    splits a cell's radial axis (uDir = unit(cell centre), mesh.cpp:1299); new vertices are placed on the cube shell and projected with
    the rest.  Same 24 cells refined (600 -> 672), same centroids and curved-element volumes."""
    from nebulasem_b200 import host
    src = os.path.join(os.path.dirname(GOLD), "amr_run", "acoustic-sphere-amr-dg")
    d = str(tmp_path / "case")
    shutil.copytree(src, d)
    exp = np.load(os.path.join(d, "expected.npz"))
    s = host.Solver.open_case(d)
    assert s.nBCS == 600
    old_nodes = s.f64("cC").reshape(-1, 3)[: s.gBCSfield].reshape(s.nBCS, s.NP, 3).copy()
    s.regrid()
    assert s.nBCS == exp["grid0_CC"].shape[0] == 672
    cc, cv = s.f64("gCC")[: 3 * s.nBCS].reshape(-1, 3), s.f64("gCV")[: s.nBCS]
    new_nodes = s.f64("cC").reshape(-1, 3)[: s.gBCSfield].reshape(s.nBCS, s.NP, 3)
    # cells the regrid does not touch keep their local axes (three of the six panels are left-handed in the block file): refineField copies
    # their values node by node, and on the sphere an element's interior nodes sit where they do because of which axis is which
    cell_map = s.u32("cellMap")[:600]               # one entry per old cell first (the new cells' entries follow, mesh.cpp:2216-2748)
    kept = np.nonzero(cell_map < s.nBCS)[0]
    assert len(kept) == 600 - 24
    assert np.abs(new_nodes[cell_map[kept]] - old_nodes[kept]).max() <= 1e-8
    # and every node of the regridded mesh is where the reference's geometry of ITS regridded grid has it
    from scipy.spatial import cKDTree
    mine = np.concatenate([np.repeat(cc, s.NP, axis=0), new_nodes.reshape(-1, 3)], axis=1)
    ref = np.concatenate([np.repeat(exp["grid0_CC"], int(exp["NP"]), axis=0), exp["half_node_xyz"]], axis=1)
    dist, idx = cKDTree(ref).query(mine)
    assert dist.max() <= 1e-7 and len(np.unique(idx)) == len(idx), dist.max()
    s.close()
    scale = np.abs(exp["grid0_CC"]).max()
    key = lambda x: np.round(x / scale * 1e6).astype(np.int64)
    oa, ob = np.lexsort(key(cc).T[::-1]), np.lexsort(key(exp["grid0_CC"]).T[::-1])
    assert np.array_equal(key(cc)[oa], key(exp["grid0_CC"])[ob])
    assert np.abs(cc[oa] - exp["grid0_CC"][ob]).max() <= 1e-12 * scale
    assert np.abs(cv[oa] / exp["grid0_CV"][ob] - 1).max() <= 1e-11


def test_regridded_shell_geometry_is_bit_identical_to_the_reference(tmp_path):
    """The reference's OWN regridded cubed sphere (24 cells refined, 2:1 faces; tests/golden/sphere/acoustic-sphere-regridded, by
    make_sphere_golden.make_regridded) and the SHA-256 of every geometry array its LoadMesh builds there: spherical corrections of calcGeometry
    on cells with more than six facets, node placement from merged sides, mortar flags.  Oracle and C++ host reproduce all of them bit for bit."""
    import hashlib
    import json

    from nebulasem_b200 import host
    src = os.path.join(GOLD, "acoustic-sphere-regridded")
    sums = json.load(open(os.path.join(src, "geom_sha256.json")))
    sha = lambda a, dt: hashlib.sha256(np.ascontiguousarray(np.asarray(a).ravel(), dtype=dt).tobytes()).hexdigest()
    orc = ocase.load_case(src, exact_order=False)
    g = orc.g
    assert [g.basis.NPX, g.basis.NPY, g.basis.NPZ, g.basis.NP, g.basis.NPF, g.nBCS] == sums["dims"][:6]
    assert np.count_nonzero(np.asarray(g.topo.FMC)) > 0
    for nm, arr in (("cC", g.cC), ("cV", g.cV), ("Jinv", g.Jinv), ("fN", g.fN), ("fC", g.fC), ("fI", g.fI), ("gFN", g.topo.FNv), ("gFC", g.topo.FC),
                    ("gCV", g.topo.CV), ("gCC", g.topo.CC)):
        assert sha(arr, "<f8") == sums[nm], "oracle " + nm
    for nm, arr in (("FO", g.FO), ("FN", g.FN), ("gFMC", g.topo.FMC), ("gFOC", g.topo.FOC), ("gFNC", g.topo.FNC)):
        assert sha(arr, "<i8") == sums[nm], "oracle " + nm
    d = str(tmp_path / "case")
    shutil.copytree(src, d)
    s = host.Solver.open_case(d)
    for nm, key in (("cC", "cC"), ("cV", "cV"), ("Jinv", "Jinv"), ("fN", "fN"), ("fC", "fC"), ("fI", "fI"), ("faceNormal", "gFN"), ("faceCenter", "gFC")):
        assert sha(s.f64(nm), "<f8") == sums[key], "host " + nm
    for nm, key in (("FO", "FO"), ("FN", "FN"), ("faceMortar", "gFMC"), ("faceOwner", "gFOC"), ("faceNeigh", "gFNC")):
        assert sha(s.u32(nm).astype(np.int64), "<i8") == sums[key], "host " + nm
    s.close()


def test_partitions_of_a_regridded_shell_keep_the_face_geometry(tmp_path):
    """The curved-element corrections visit entries 2..5 of a cell's facet list (mesh.cpp:529); a further sub-facet of a split side gets its
    area and centre from the cell on its other side -- which a partition may not hold.  Then the cell that does hold it applies them, so the
    faces of every part equal those of the whole mesh (this was 1e-6 in the solution of a two-partition AMR run before)."""
    import re

    from nebulasem_b200 import host
    d = str(tmp_path / "case")
    shutil.copytree(os.path.join(os.path.dirname(GOLD), "amr_run", "acoustic-sphere-amr-dg"), d)
    s = host.Solver.open_case(d)
    s.regrid()
    s.write_amr_grid(0)                                  # grid_0.bin is now the regridded mesh (2:1 faces along the refined patch)
    s.close()
    os.remove(os.path.join(d, "grid_0.forest"))
    ctl = re.sub(r"(?m)^\s*amr_step\s+\d+\s*\n", "", open(os.path.join(d, "controls")).read())
    open(os.path.join(d, "controls"), "w").write(ctl)

    def faces_of_cells(s):
        fb, fe, af = s.u32("faceBegin"), s.u32("faceEnd"), s.u32("allFaces")
        area = np.linalg.norm(s.f64("faceNormal").reshape(-1, 3), axis=1)
        radius = np.linalg.norm(s.f64("faceCenter").reshape(-1, 3), axis=1)
        return [np.array(sorted(zip(area[af[fb[c]:fe[c]]], radius[af[fb[c]:fe[c]]]))) for c in range(s.nBCS)]

    g = host.Solver.open_case(d)
    whole = faces_of_cells(g)
    assert max(len(f) for f in whole) > 4                # cells with a split side are there (radial faces are deleted: 4 = conforming)
    for nparts in (2, 3, 5):
        for r in range(nparts):
            p = host.Solver.open_case(d, 0, r, nparts)
            mine = faces_of_cells(p)
            for l, c in enumerate(p.u32("cellGlobal")):
                assert mine[l].shape == whole[c].shape
                assert np.abs(mine[l] / whole[c] - 1).max() <= 1e-12, (nparts, r, l, c)
            p.close()
    g.close()


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


@pytest.mark.gpu
def test_two_partitions_of_the_shell_equal_one_partition(tmp_path):
    """hydro-sphere split over two GPUs (METIS, as its controls say; halo across panel edges, rho without a condition on top/bottom in both
    parts) == the single-partition run."""
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = str(tmp_path / "hydro-sphere")
    shutil.copytree(os.path.join(GOLD, "hydro-sphere"), d)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(root, "tests", "mp_gpu_check.py"), "METIS"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MP_CHECK_CASE=d, MP_CHECK_STEPS="20"))
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0 and "MP_CHECK_OK" in out.stdout
