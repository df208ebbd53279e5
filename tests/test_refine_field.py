"""AMR field transfer (MeshField::refineField, src/field/field.h:1863-2015).

Golden vectors: tests/golden/refine_field/ -- two regrids each of a 2-D order-4 and a 3-D order-2 rising-bubble case, driven through the
UNMODIFIED reference's refineMesh/refineField by oracle/tools/refinedump.cpp (tests/golden/make_refine_golden.py): pass 1 splits a block of
cells, pass 2 merges families back, splits coarse cells and one fine cell.  CPU tests pin the numpy restatement (oracle/amr.py) bit for
bit; GPU tests run the device-resident transfer (nsem_refine_state / nsem_restart_state through the C++ host) against the same vectors.
"""
import glob
import os

import numpy as np
import pytest

from oracle import amr

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "refine_field")
PASSES = [("2d_o4", 1), ("2d_o4", 2), ("3d_o2", 1), ("3d_o2", 2)]
FIELDS = (("rho", 1), ("U", 3), ("T", 1), ("p", 1))


def load(name, k):
    g = dict(np.load(os.path.join(GOLD, f"{name}_pass{k}.npz")))
    d = g["dims"]
    g["npts"] = tuple(int(x) for x in d[:3])
    g["NP"] = int(d[3])
    g["nOld"] = int(d[4])
    g["nNew"] = int(d[6])
    g["psiRef"] = [g[f"psiRef{i}"] for i in range(6)]
    g["psiCor"] = [g[f"psiCor{i}"] for i in range(6)]
    g["wgl"] = [g[f"wgl{i}"] for i in range(3)]
    return g


def oracle_transfer(g, P):
    return amr.refine_field(P, g["npts"], g["refineMap"], g["coarseMap"], g["cellMap"], g["nNew"], g["oldCV"], g["oldCC"], g["newCC"],
                            g["newCV"], g["cC_old"], g["psiRef"], g["psiCor"], g["wgl"])


def test_fixtures_cover_copy_split_and_merge():
    assert len(glob.glob(os.path.join(GOLD, "*.npz"))) == 4
    for name, k in PASSES:
        g = load(name, k)
        assert len(g["refineMap"]) > 0
        assert (len(g["coarseMap"]) > 0) == (k == 2)
        assert (g["cellMap"][:g["nOld"]] != amr.MAX_INT).sum() > 0       # copied cells


@pytest.mark.parametrize("name,k", PASSES)
def test_oracle_transfer_is_bit_identical_to_the_reference(name, k):
    g = load(name, k)
    for f, comps in FIELDS:
        pre = g["pre_" + f].reshape(-1, comps)
        post = g["post_" + f].reshape(-1, comps)[:g["nNew"] * g["NP"]]
        out = oracle_transfer(g, pre)
        assert np.array_equal(out, post), (f, np.abs(out - post).max())


@pytest.mark.parametrize("name,k", PASSES)
def test_transfer_properties(name, k):
    """Size-independent properties of the operator: a constant stays that constant (the psi tables are partitions of unity and the
    mass-fix factor of a constant is 1), and the integral of every split family equals the integral over its parent."""
    g = load(name, k)
    NP = g["NP"]
    const = np.full((g["nOld"] * NP, 1), 3.25)
    out = oracle_transfer(g, const)
    assert np.abs(out - 3.25).max() <= 1e-13
    pre = g["pre_T"].reshape(-1, 1) + 300.0                              # one-signed, so |integral| = integral
    out = oracle_transfer(g, pre).reshape(g["nNew"], NP)
    ii, jj, kk = np.meshgrid(*[np.arange(n) for n in g["npts"]], indexing="ij")
    w = (g["wgl"][0][ii] * g["wgl"][1][jj] * g["wgl"][2][kk]).ravel() / 8
    rm, cm, i = g["refineMap"].astype(np.int64), g["cellMap"].astype(np.int64), 0
    while i < len(rm):
        n, parent = rm[i], rm[i + 1]
        kids = cm[rm[i + 2:i + 2 + n]]
        old = (pre.reshape(-1, NP)[parent] * w).sum() * g["oldCV"][parent]
        new = sum((out[c] * w).sum() * g["newCV"][c] for c in kids)
        assert abs(new - old) <= 1e-12 * abs(old)
        i += n + 2


@pytest.mark.parametrize("name", ["2d_o4", "3d_o2"])
def test_host_geometry_of_the_regridded_stages_matches_the_reference(name):
    """What EulerSolver::adopt_refined_state hands to nsem_refine_state -- cell volumes, centroids, node coordinates, psiRef/psiCor of the
    C++ host -- equals what the reference handed to refineField."""
    from nebulasem_b200 import host
    for k in (1, 2):
        g = load(name, k)
        old = host.Solver.open_case(os.path.join(GOLD, name, f"stage{k - 1}"))
        new = host.Solver.open_case(os.path.join(GOLD, name, f"stage{k}"))
        nOld, nNew, NP = g["nOld"], g["nNew"], g["NP"]
        assert old.nBCS == nOld and new.nBCS == nNew
        assert np.array_equal(old.f64("gCV")[:nOld], g["oldCV"][:nOld])
        assert np.array_equal(old.f64("gCC")[:nOld * 3], g["oldCC"][:nOld * 3])
        assert np.array_equal(new.f64("gCV")[:nNew], g["newCV"][:nNew])
        assert np.array_equal(new.f64("gCC")[:nNew * 3], g["newCC"][:nNew * 3])
        assert np.array_equal(old.f64("cC")[:nOld * NP * 3], g["cC_old"][:nOld * NP * 3])
        old.close()
        new.close()


def test_transfer_entry_points_refuse_to_run_without_a_device():
    import torch
    from nebulasem_b200 import capi, host
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    a = host.Solver.open_case(os.path.join(GOLD, "3d_o2", "stage0"))
    b = host.Solver.open_case(os.path.join(GOLD, "3d_o2", "stage1"))
    g = load("3d_o2", 1)
    with pytest.raises(capi.NsemError, match="no CPU fallback"):
        b.adopt_refined_state(a, g["refineMap"], g["coarseMap"], g["cellMap"])
    with pytest.raises(capi.NsemError, match="no CPU fallback"):
        b.restart_state()
    a.close()
    b.close()


# ---------------------------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------------------------
def _set_real_nodes(s, g, which):
    n = g["nOld"] * g["NP"]
    rho, U, T, p = [x.copy() for x in s.state()]
    rho[:n] = g[which + "_rho"].reshape(-1)[:n]
    U[:n] = g[which + "_U"].reshape(-1, 3)[:n]
    T[:n] = g[which + "_T"].reshape(-1)[:n]
    p[:n] = g[which + "_p"].reshape(-1)[:n]
    s.set_state(rho, U, T, p)
    s.upload()


@pytest.mark.gpu
@pytest.mark.parametrize("name,k", PASSES)
def test_device_transfer_is_bit_identical_to_the_reference(name, k):
    """nsem_refine_state (device to device, the state never leaves HBM) reproduces the fields the reference wrote after its regrid,
    bit for bit: copy, interpolation + mass fix, projection all keep the reference's floating-point operation order."""
    from nebulasem_b200 import host
    g = load(name, k)
    old = host.Solver.open_case(os.path.join(GOLD, name, f"stage{k - 1}"))
    new = host.Solver.open_case(os.path.join(GOLD, name, f"stage{k}"))
    old.attach(0)
    new.attach(0)
    _set_real_nodes(old, g, "pre")
    n0 = new.launch_count
    new.adopt_refined_state(old, g["refineMap"], g["coarseMap"], g["cellMap"], restart=False)
    assert new.launch_count - n0 >= 2                                 # copy + split (+ merge) kernels ran
    new.download()
    n = g["nNew"] * g["NP"]
    worst = 0.0
    for (f, comps), dev in zip(FIELDS, new.state()):
        post = g["post_" + f].reshape(-1, comps)[:n]
        dev = np.asarray(dev).reshape(-1, comps)[:n]
        scale = np.abs(post).max()
        worst = max(worst, np.abs(dev - post).max() / scale)
        assert np.array_equal(dev, post), (f, np.abs(dev - post).max() / scale)
    old.close()
    new.close()


@pytest.mark.gpu
def test_device_transfer_then_restart_runs_on_the_regridded_mesh():
    """End to end on the device: 20 steps on the coarse grid, regrid (transfer + restart branch of the set-up), 20 steps on the
    non-conforming grid == the same 20 steps started from the reference's transferred field FILES (host set-up, start branch replaced by
    the restart branch is the only difference: p is rebuilt from rho, the ghost cells from the boundary conditions)."""
    from nebulasem_b200 import host
    name, k = "3d_o2", 1
    g = load(name, k)
    old = host.Solver.open_case(os.path.join(GOLD, name, "stage0"))
    old.attach(0)
    _set_real_nodes(old, g, "pre")
    new = host.Solver.open_case(os.path.join(GOLD, name, "stage1"))
    new.attach(0)
    new.adopt_refined_state(old, g["refineMap"], g["coarseMap"], g["cellMap"], restart=True)
    # the same state uploaded from the host: real nodes from the reference's files, ghost cells and p from the device's restart pass
    new.download()
    ref = [x.copy() for x in new.state()]
    n = g["nNew"] * g["NP"]
    for (f, comps), dev in zip(FIELDS[:3], ref[:3]):
        assert np.array_equal(np.asarray(dev).reshape(-1, comps)[:n], g["post_" + f].reshape(-1, comps)[:n]), f
    twin = host.Solver.open_case(os.path.join(GOLD, name, "stage1"))
    twin.attach(0)
    twin.set_state(*ref)
    twin.upload()
    new.step(20)
    twin.step(20)
    new.download()
    twin.download()
    for a, b in zip(new.state(), twin.state()):
        assert np.isfinite(a).all()
        assert np.array_equal(a, b)
    # mass is what the transfer left (conservative fluxes on the non-conforming grid)
    rho = new.state()[0][:n]
    cV = new.f64("cV")[:n]
    m0 = float((g["post_rho"].reshape(-1)[:n] * cV).sum())
    assert abs(float((rho * cV).sum()) - m0) <= 1e-12 * abs(m0)
    for s in (old, new, twin):
        s.close()


@pytest.mark.gpu
def test_restart_state_rebuilds_ghost_cells_and_pressure():
    """nsem_restart_state (euler.cpp:150-162 on the device): a run continued from real-node values only -- ghost cells zeroed, p discarded --
    is bit-identical to the uninterrupted run."""
    from nebulasem_b200 import host
    a = host.Solver.synthetic("bubble3d", 4, 3, 3, 4)
    b = host.Solver.synthetic("bubble3d", 4, 3, 3, 4)
    a.attach(0)
    b.attach(0)
    a.step(5)
    a.download()
    rho, U, T, p = [x.copy() for x in a.state()]
    n = a.gBCSfield
    rho[n:] = 0.0
    U[n:] = 0.0
    T[n:] = 0.0
    p[:] = 0.0
    b.set_state(rho, U, T, p)
    b.upload()
    b.restart_state()
    a.step(5)
    b.step(5)
    a.download()
    b.download()
    for x, y in zip(a.state(), b.state()):
        assert np.array_equal(x, y)
    a.close()
    b.close()
