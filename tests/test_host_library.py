"""C++ host side (libnsem_host.so) against the oracle: file readers, topology, DG geometry, euler set-up, dump writer.
CPU only: nothing here touches the GPU."""
import os
import re

import numpy as np
import pytest

from nebulasem_b200 import capi, host
from oracle import refio
from tests.helpers import make_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,kw", [
    ("bubble2d", dict(n=4, order=4)),
    ("bubble3d", dict(n=3, order=3)),
    ("vortex", dict(n=4, order=5)),
    ("hill3d", dict(nx=5, ny=2, nz=3, order=2)),
])
def test_host_geometry_and_setup_bit_equal_to_oracle(tmp_cases, name, kw):
    orc = make_oracle(tmp_cases, name, 1, exact=False, **kw)
    s = host.Solver.open_case(orc.case_dir)
    g = orc.g
    assert (s.NPX, s.NPY, s.NPZ, s.NP, s.NPF) == (g.basis.NPX, g.basis.NPY, g.basis.NPZ, g.basis.NP, g.basis.NPF)
    assert (s.nBCS, s.nCells, s.nFacets) == (g.nBCS, g.nCells, g.nFacets)
    for nm, ref in (("cC", g.cC), ("cV", g.cV), ("Jinv", g.Jinv), ("fN", g.fN), ("fC", g.fC), ("fI", g.fI), ("faceNormal", g.topo.FNv)):
        assert np.array_equal(s.f64(nm), np.asarray(ref).ravel()), nm
    for nm, ref in (("FO", g.FO), ("FN", g.FN), ("allFaces", g.allFaces), ("faceBegin", g.faceIndices[0]), ("faceEnd", g.faceIndices[1]),
                    ("faceOwner", g.topo.FOC), ("faceNeigh", g.topo.FNC), ("faceID", np.concatenate(g.topo.faceID))):
        assert np.array_equal(s.u32(nm), np.asarray(ref, dtype=np.uint32).ravel()), nm
    rho, U, T, p = s.state()
    assert np.array_equal(rho, orc.rho) and np.array_equal(U, orc.U) and np.array_equal(T, orc.T) and np.array_equal(p, orc.pp)
    assert np.array_equal(s.f64("rho_ref"), orc.rho_ref) and np.array_equal(s.f64("p_ref"), orc.p_ref)
    m0, e0, v0 = s.totals()
    assert abs(m0 - orc.mass0) <= 1e-12 * abs(orc.mass0) and abs(v0 - orc.volume0) <= 1e-12 * abs(orc.volume0)
    s.close()


def test_host_setup_from_a_dump_bit_equal_to_oracle(tmp_path):
    """open_case(dir, 1): the fields of the reference's dump 1 on grid_0 (the newest grid <= 1), set up as a run resuming there is
    (rho from p; tests/test_oracle_vs_reference.py pins that against the reference) -- bit-equal to the oracle on the same files."""
    import shutil

    from oracle import case as ocase
    gold = os.path.join(ROOT, "tests", "golden", "vtk", "bubble3d_n2_o2")
    d = str(tmp_path / "resumed")
    shutil.copytree(gold, d)
    d2 = str(tmp_path / "as_step0")
    os.makedirs(d2)
    shutil.copy(os.path.join(d, "controls"), d2)
    shutil.copy(os.path.join(d, "grid_0.txt"), d2)
    for f in ("rho", "U", "T", "p"):
        shutil.copy(os.path.join(d, f + "1.bin"), os.path.join(d2, f + "0.bin"))
    orc = ocase.load_case(d2, exact_order=False)
    s = host.Solver.open_case(d, 1)
    rho, U, T, p = s.state()
    assert np.array_equal(rho, orc.rho) and np.array_equal(U, orc.U) and np.array_equal(T, orc.T) and np.array_equal(p, orc.pp)
    s.close()


@pytest.mark.parametrize("fixture", ["srtb_amr", "srtb3d_amr"])
def test_host_non_conforming_topology_bit_equal_to_oracle(tmp_path, fixture):
    """Non-conforming (2:1 AMR) grids written by the reference's regrid: the C++ host's general fixHexCells route (coplanar
    sub-facets grouped and merged per side, gFMC), node geometry from merged sides and the mortar projections
    psiRef/psiCor (dg.cpp:550-590) are bit-equal to the oracle, which is itself bit-equal to the reference on these
    grids (tests/test_oracle_golden.py)."""
    import shutil

    from oracle import case as ocase
    d = str(tmp_path / fixture)
    shutil.copytree(os.path.join(ROOT, "tests", "golden", fixture), d)
    orc = ocase.load_case(d, exact_order=False)
    s = host.Solver.open_case(d)
    g = orc.g
    assert (s.nBCS, s.nCells, s.nFacets) == (g.nBCS, g.nCells, g.nFacets)
    assert np.count_nonzero(np.asarray(g.topo.FMC)) > 0
    for nm, ref in (("cC", g.cC), ("cV", g.cV), ("Jinv", g.Jinv), ("fN", g.fN), ("fC", g.fC), ("fI", g.fI), ("faceNormal", g.topo.FNv),
                    ("faceCenter", g.topo.FC)):
        assert np.array_equal(s.f64(nm), np.asarray(ref).ravel()), nm
    for nm, ref in (("FO", g.FO), ("FN", g.FN), ("allFaces", g.allFaces), ("faceBegin", g.faceIndices[0]), ("faceEnd", g.faceIndices[1]),
                    ("faceOwner", g.topo.FOC), ("faceNeigh", g.topo.FNC), ("faceID", np.concatenate(g.topo.faceID)),
                    ("faceMortar", g.topo.FMC)):
        assert np.array_equal(s.u32(nm), np.asarray(ref, dtype=np.uint32).ravel()), nm
    for q in range(6):
        assert np.array_equal(s.f64(f"psiRef{q}"), g.basis.psiRef[q]), q
        assert np.array_equal(s.f64(f"psiCor{q}"), g.basis.psiCor[q]), q
    rho, U, T, p = s.state()
    assert np.array_equal(rho, orc.rho) and np.array_equal(U, orc.U) and np.array_equal(T, orc.T) and np.array_equal(p, orc.pp)
    s.close()


def test_decomposition_keeps_mortar_faces_whole():
    """The reference weighs non-conforming faces 1000 in the METIS graph (field.cpp:1037-1047) and refuses a decomposition that cuts one
    (field.cpp:1215-1220); here the cells they join are contracted into one graph vertex, so none can be cut."""
    from oracle import mesh as omesh
    gp = os.path.join(ROOT, "tests", "golden", "srtb3d_amr", "grid_0")
    g = refio.read_grid(gp)
    t = omesh.MeshTopo(g).load()
    nc, nf = len(g.cells), len(g.facets)
    FOC, FNC = np.asarray(t.FOC), np.asarray(t.FNC)
    for nparts in (2, 4):
        part, fmc = host.partition_grid(gp, nc, nf, nparts, "METIS")
        assert np.array_equal(fmc, np.asarray(t.FMC, dtype=np.uint32))       # 3-D: no face is deleted, numbering unchanged
        m = np.nonzero(fmc)[0]
        assert len(m) == 160 and not (part[FOC[m]] != part[FNC[m]]).any()
        counts = np.bincount(part, minlength=nparts)
        assert counts.min() > 0 and counts.max() <= 1.1 * nc / nparts
    # round 1 refused a decomposition by cell index here (it cut mortar faces); cells joined by 2:1 faces now travel together in every method
    part, _ = host.partition_grid(gp, nc, nf, 2, "CELLID")
    assert not (part[FOC[m]] != part[FNC[m]]).any() and np.bincount(part, minlength=2).min() > 0


def test_field_dump_written_in_the_reference_format(tmp_cases):
    orc = make_oracle(tmp_cases, "bubble3d", 1, exact=False, n=2, order=2)
    s = host.Solver.open_case(orc.case_dir)
    s.write(7)            # rho7.bin, U7.bin, T7.bin, p7.bin (write_format defaults to BINARY, field.cpp:74)
    nb = orc.gB
    for nm, ref in (("rho", orc.rho), ("U", orc.U), ("T", orc.T), ("p", orc.pp)):
        ff = refio.read_field(os.path.join(orc.case_dir, f"{nm}7"))
        vals = ff.values[:, 0] if ff.comps == 1 else ff.values
        assert np.array_equal(vals, ref[:nb])
        assert [b.kind for b in ff.bcs] == [b.kind for b in orc.bcs[nm]]
    s.close()


def test_synthetic_generator_matches_case_files(tmp_cases):
    orc = make_oracle(tmp_cases, "bubble3d", 1, exact=False, n=3, order=4)
    a = host.Solver.open_case(orc.case_dir)
    b = host.Solver.synthetic("bubble3d", 3, 3, 3, 4)
    for nm in ("cC", "cV", "Jinv", "fN", "rho", "U", "T", "p", "rho_ref", "p_ref"):
        assert np.array_equal(a.f64(nm), b.f64(nm)), nm
    for nm in ("FO", "FN", "allFaces", "faceID"):
        assert np.array_equal(a.u32(nm), b.u32(nm)), nm


def test_errors_are_reported_not_swallowed(tmp_cases):
    with pytest.raises(capi.NsemError):
        host.Solver.open_case(os.path.join(str(tmp_cases), "does_not_exist"))
    with pytest.raises(capi.NsemError, match="unknown synthetic case"):
        host.Solver.synthetic("nonsense", 2, 2, 2, 2)
    orc = make_oracle(tmp_cases, "bubble2d", 1, exact=False, n=3, order=2)
    s = host.Solver.open_case(orc.case_dir)
    with pytest.raises(capi.NsemError, match="no device attached|no CPU fallback"):
        s.step(1)        # the hot path has no CPU fallback
    # adaptive regridding (AmrIteration, iteration.h:94-147): asked for, the solver keeps an AMR forest over the grid (tests/test_amr_regrid.py);
    # a grid that is already non-conforming cannot seed it and a regrid on it is refused, not skipped
    import shutil
    amr = os.path.join(str(tmp_cases), "asks_for_amr")
    shutil.copytree(orc.case_dir, amr)
    txt = open(os.path.join(amr, "controls")).read().replace("end_step", "amr_step 1\n    end_step", 1)
    open(os.path.join(amr, "controls"), "w").write(txt)
    s = host.Solver.open_case(amr)
    assert (s.cell_levels() == 0).all()
    s.close()
    amr2 = os.path.join(str(tmp_cases), "asks_for_amr_on_a_regridded_grid")
    shutil.copytree(os.path.join(ROOT, "tests", "golden", "srtb3d_amr"), amr2)
    txt = open(os.path.join(amr2, "controls")).read().replace("end_step", "amr_step 1\n    end_step", 1)
    open(os.path.join(amr2, "controls"), "w").write(txt)
    s = host.Solver.open_case(amr2)
    with pytest.raises(capi.NsemError, match="no AMR forest"):
        s.regrid()
    s.close()


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nsem_c.h")).read()
    declared = sorted(set(re.findall(r"\b(nsem_[a-z_0-9]+)\s*\(", hdr)))
    assert set(declared) == set(capi.EXPORTS), (declared, capi.EXPORTS)
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    hl = host.load_host_library()
    for name in host.HOST_EXPORTS:
        assert hasattr(hl, name), name


def test_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.NsemError, match="no CUDA device"):
        capi.Context(0)
