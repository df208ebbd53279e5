"""The numpy oracle against the golden fixtures produced by the unmodified reference binary (CPU, no GPU)."""
import ast
import glob
import os

import numpy as np
import pytest

from tests.helpers import make_oracle, rel_l2

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_reference_dump(tmp_cases, path):
    g = np.load(path)
    case, kw, nsteps = str(g["case"]), ast.literal_eval(str(g["kwargs"])), int(g["nsteps"])
    orc = make_oracle(tmp_cases, case, nsteps, exact=True, **kw)
    orc.run(nsteps)
    nb = orc.gB
    # the exact-order oracle is bit-identical to the reference where libm is the same; 1e-13 allows another libm
    for name, mine in (("rho", orc.rho), ("U", orc.U), ("T", orc.T), ("p", orc.pp)):
        ref = g[name]
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(mine[:nb] - ref).max() <= 1e-12 * scale, name
    assert rel_l2(orc.rho[:nb], g["rho"]) <= 1e-14


@pytest.mark.parametrize("path", GOLD[:3], ids=[os.path.basename(p)[:-4] for p in GOLD[:3]])
def test_fast_order_oracle_within_tolerance(tmp_cases, path):
    """The einsum (fast) evaluation order used for the larger GPU comparisons stays within 1e-12 of the reference."""
    g = np.load(path)
    case, kw, nsteps = str(g["case"]), ast.literal_eval(str(g["kwargs"])), int(g["nsteps"])
    orc = make_oracle(tmp_cases, case, nsteps, exact=False, **kw)
    orc.run(nsteps)
    nb = orc.gB
    assert rel_l2(orc.rho[:nb], g["rho"]) <= 1e-13
    c0 = np.sqrt(orc.gamma * orc.R * orc.p.T0)
    mom_ref = g["rho"][:, None] * g["U"]
    mom = orc.rho[:nb, None] * orc.U[:nb]
    assert np.linalg.norm(mom - mom_ref) / (np.linalg.norm(g["rho"]) * c0) <= 1e-13
    th_ref = g["rho"] * (g["T"] + orc.p.T0)
    assert rel_l2(orc.rho[:nb] * (orc.T[:nb] + orc.p.T0), th_ref) <= 1e-13


@pytest.mark.parametrize("fixture,nmortar", [("srtb_amr", 56), ("srtb3d_amr", 160)])
def test_oracle_mortar_faces_reproduce_reference_dump(tmp_path, fixture, nmortar):
    """Non-conforming meshes made by make_amr_golden.py from the reference's own regrid (tests/golden/srtb_amr: 2-D order 4,
    196 cells, 56 mortar sub-faces of examples/atmo/srtb-amr; tests/golden/srtb3d_amr: 3-D order 2, 372 cells, 160 mortar
    sub-faces, four per coarse face): the oracle's scatter/gather_non_conforming and psiRef/psiCor against the
    reference's dump after 20 steps -- bit-identical where libm is the same."""
    import shutil

    from oracle import case as ocase
    src = os.path.join(os.path.dirname(__file__), "golden", fixture)
    d = str(tmp_path / fixture)
    shutil.copytree(src, d)
    g = np.load(os.path.join(src, "expected.npz"))
    orc = ocase.load_case(d, exact_order=True)
    assert len(orc.mortar_faces) == nmortar and orc.has_mortar
    orc.run(int(g["nsteps"]))
    nb = orc.gB
    for name, mine in (("rho", orc.rho), ("U", orc.U), ("T", orc.T), ("p", orc.pp)):
        ref = g[name]
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(mine[:nb] - ref).max() <= 1e-12 * scale, name
    assert np.array_equal(orc.rho[:nb], g["rho"]) or rel_l2(orc.rho[:nb], g["rho"]) <= 1e-14
    # the mortar pair is conservative: mass is conserved to rounding on the non-conforming mesh
    mass0 = None
    orc2 = ocase.load_case(d, exact_order=False)
    m0 = float((orc2.rho[:nb] * orc2.g.cV[:nb]).sum())
    orc2.run(5)
    m1 = float((orc2.rho[:nb] * orc2.g.cV[:nb]).sum())
    assert abs(m1 - m0) <= 1e-13 * abs(m0)


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5, 6, 7])
def test_mortar_projections_preserve_constants_and_integrals(order):
    """psiRef/psiCor (dg.cpp:550-590): the coarse -> fine projection reproduces constants on each half, and the fine -> coarse
    projection conserves the integral (a flux through the two halves equals the flux the coarse face receives) -- what makes the
    scatter/gather pair of a 2:1 face conservative."""
    from oracle.dg import Basis
    b = Basis((order, order, order))
    n = order + 1
    w = b.wgl[0]
    rng = np.random.default_rng(order)
    for h in range(2):
        R = b.psiRef[h].reshape(n, n)          # [in, io]
        assert np.allclose(R.sum(axis=0), 1.0, atol=1e-12)
    # gather: a sub-facet carries HALF the coarse face's area with the FULL quadrature weights (fN = gFN_sub * w_a w_b / 4), so the
    # projection of its flux must keep the weighted sum: sum_in w_in (f @ psiCor_h)[in] == sum_io w_io f[io]
    for h in range(2):
        f = rng.standard_normal(n)
        coarse = f @ b.psiCor[h].reshape(n, n)                       # coarse[in] = sum_io f[io] psiCor[h][io*n+in]
        lhs, rhs = float((w * coarse).sum()), float((w * f).sum())
        assert abs(lhs - rhs) <= 1e-12 * max(1.0, abs(rhs))
