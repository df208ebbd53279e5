"""GPU parity: CUDA path (through the C ABI) vs the numpy oracle on the same seeded cases.  -m gpu"""
import ast
import glob
import os

import numpy as np
import pytest

from tests.helpers import conserved_errors, device_from_oracle, make_oracle, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-11   # north_star: relative L2 per conserved variable


@pytest.mark.parametrize("name,kw,nsteps", [
    ("bubble2d", dict(n=6, order=4), 20),
    ("bubble3d", dict(n=3, order=4), 10),
    ("bubble3d", dict(n=3, order=2), 10),
    ("vortex", dict(n=6, order=3), 20),
    ("hill3d", dict(nx=6, ny=2, nz=4, order=3), 10),
])
def test_steps_match_oracle(tmp_cases, name, kw, nsteps):
    orc = make_oracle(tmp_cases, name, nsteps, exact=False, **kw)
    ctx = device_from_oracle(orc)
    if name in ("bubble3d", "hill3d"):
        # straight-edged 3-D meshes (affine boxes and the terrain-following hill): metrics evaluated on the fly
        assert "on the fly" in ctx.kernel_info, ctx.kernel_info
    # state round trip first: download must return exactly what was uploaded
    rho, U, T, p = ctx.download_state()
    assert np.array_equal(rho[: orc.gB], orc.rho[: orc.gB])
    assert np.array_equal(U[: orc.gB], orc.U[: orc.gB])
    ctx.step(nsteps)
    orc.run(nsteps)
    rho, U, T, p = ctx.download_state()
    err = conserved_errors(orc, rho, U, T)
    print(name, kw, nsteps, err)
    assert np.isfinite(rho).all() and np.isfinite(U).all() and np.isfinite(T).all()
    assert err["rho"] <= TOL
    assert err["rhoTheta"] <= TOL
    assert err["rhoU_scaled"] <= TOL
    ctx.close()


_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


@pytest.mark.parametrize("path", _GOLD, ids=[os.path.basename(q)[:-4] for q in _GOLD])
def test_device_matches_the_reference_binary_dumps(tmp_cases, path):
    """The CUDA path against the dumps of the UNMODIFIED reference binary (tests/golden/*.npz, made by make_golden.py), at their full
    step counts -- among them 100 steps of examples/atmo/srtb (BASELINE configs[0], north_star: <= 1e-11 after 100 steps) --
    with no oracle run in between."""
    g = np.load(path)
    case, kw, nsteps = str(g["case"]), ast.literal_eval(str(g["kwargs"])), int(g["nsteps"])
    orc = make_oracle(tmp_cases, case, nsteps, exact=False, **kw)
    ctx = device_from_oracle(orc)
    ctx.step(nsteps)
    rho, U, T, p = ctx.download_state()
    ctx.close()
    nb, T0 = orc.gB, orc.p.T0
    c0 = np.sqrt(orc.gamma * orc.R * T0)
    err = dict(rho=rel_l2(rho[:nb], g["rho"]),
               rhoTheta=rel_l2(rho[:nb] * (T[:nb] + T0), g["rho"] * (g["T"] + T0)),
               rhoU_scaled=rel_l2(rho[:nb, None] * U[:nb], g["rho"][:, None] * g["U"], scale=np.linalg.norm(g["rho"]) * c0),
               rhoU_self=rel_l2(rho[:nb, None] * U[:nb], g["rho"][:, None] * g["U"]))
    spread = float(g["spread_rhoU_self"])
    print(os.path.basename(path), nsteps, err, "reference -O2 vs -O3 rhoU_self spread:", spread)
    assert err["rho"] <= TOL and err["rhoTheta"] <= TOL and err["rhoU_scaled"] <= TOL, err
    # SURVEY finding 6 / 8(c), the second half of the momentum criterion: self-relative rho*U within 1e-11, or -- where the reference's own
    # -O2 and -O3 builds differ by more than that on this very case (the fixture holds their measured spread) -- within 3 x that spread
    assert err["rhoU_self"] <= max(TOL, 3.0 * spread), (err, spread)


@pytest.mark.parametrize("name,kw", [("bubble3d", dict(n=3, order=3)), ("hill3d", dict(nx=6, ny=2, nz=4, order=3)), ("bubble2d", dict(n=5, order=4))])
def test_gradient_operator_matches_oracle(tmp_cases, name, kw):
    """Operator-level parity: gradf<strong>(U), gradf<strong>(theta) + fillBCs(r, fIndex) as the fused sweep A leaves
    them (nsem_download_gradients) against the oracle's gradf (field.h:3328-3362, 2731-2769), ghost nodes included."""
    orc = make_oracle(tmp_cases, name, 4, exact=False, **kw)
    orc.run(3)                                                      # a state with velocity
    ctx = device_from_oracle(orc)
    GU = orc.gradf(orc.U, "U")                                      # [node, a, b] = d_a U_b
    GT = orc.gradf(orc.T + orc.p.T0, "T")
    ctx.step(1)
    dU, dT = ctx.download_gradients()
    ctx.close()
    rm = [0, 4, 8, 1, 5, 2, 3, 7, 6]                                # Tensor AoS order <- row-major a*3+b
    GU = GU.reshape(-1, 9)[:, rm]
    nb = orc.gB
    touched = np.zeros(GU.shape[0], bool)                           # real nodes + the ghost nodes a boundary face touches
    touched[:nb] = True
    touched[np.abs(dU).sum(axis=1) + np.abs(dT).sum(axis=1) > 0] = True
    sU, sT = np.abs(GU[:nb]).max(), np.abs(GT[:nb]).max()
    assert sU > 0 and sT > 0
    assert np.abs(dU[:nb] - GU[:nb]).max() <= 1e-11 * sU
    assert np.abs(dT[:nb] - GT[:nb]).max() <= 1e-10 * sT        # theta carries the 300 K offset: eps * 300 * |D| / h on both sides
    assert np.abs(dU[touched] - GU[touched]).max() <= 1e-11 * sU
    assert np.abs(dT[touched] - GT[touched]).max() <= 1e-10 * sT


def test_non_trilinear_metrics_run_the_stored_metric_kernels(tmp_cases):
    """A mesh whose Jinv no trilinear map reproduces (curved elements) must be detected by nsem_upload_mesh and run the
    stored-metric instantiation of the v4 kernels -- and still match the oracle fed with the same metrics."""
    from oracle.dg import T9_FLAT
    nsteps = 8
    orc = make_oracle(tmp_cases, "bubble3d", nsteps, exact=False, n=3, order=4)
    g = orc.g
    rng = np.random.default_rng(7)
    idx = rng.choice(g.Jinv33.shape[0], size=40, replace=False)
    g.Jinv33[idx] *= 1 + 1e-3 * rng.standard_normal((40, 1, 1))
    g.Jinv[:] = g.Jinv33.reshape(-1, 9)[:, T9_FLAT]
    orc.Jin = g.Jinv33 * g.cV[: orc.gB, None, None]
    ctx = device_from_oracle(orc)
    assert "stored metrics" in ctx.kernel_info, ctx.kernel_info
    ctx.step(nsteps)
    orc.run(nsteps)
    rho, U, T, p = ctx.download_state()
    err = conserved_errors(orc, rho, U, T)
    print(err)
    assert err["rho"] <= TOL and err["rhoTheta"] <= TOL and err["rhoU_scaled"] <= TOL
    ctx.close()


@pytest.mark.parametrize("name,kw,syn,nsteps", [
    ("bubble3d", dict(n=3, order=4), ("bubble3d", 3, 3, 3, 4), 10),
    ("bubble2d", dict(n=6, order=4), ("bubble2d", 6, 1, 6, 4), 10),
    ("vortex", dict(n=6, order=3), ("vortex", 6, 6, 1, 3), 10),
])
def test_host_solver_path_matches_oracle(tmp_cases, name, kw, syn, nsteps):
    """C++ host (mesh + geometry + set-up) -> C ABI -> CUDA, from a case directory and from the in-memory generator."""
    from nebulasem_b200 import host
    orc = make_oracle(tmp_cases, name, nsteps, exact=False, **kw)
    orc.run(nsteps)
    for s in (host.Solver.open_case(orc.case_dir), host.Solver.synthetic(*syn)):
        s.attach(0)
        s.step(nsteps)
        s.download()
        rho, U, T, p = s.state()
        err = conserved_errors(orc, rho, U, T)
        print(name, err)
        assert err["rho"] <= TOL and err["rhoTheta"] <= TOL and err["rhoU_scaled"] <= TOL
        s.close()


@pytest.mark.parametrize("fixture,nmortar,sweeps", [("srtb_amr", 56, "v1"), ("srtb3d_amr", 160, "v4"), ("srtb3d_amr", 160, "v1")])
def test_non_conforming_mesh_matches_oracle_and_reference_dump(tmp_path, monkeypatch, fixture, nmortar, sweeps):
    """Mortar (2:1 AMR) faces, BASELINE configs[4]: the regridded rising bubble (2-D order 4: 196 cells, 56 mortar
    sub-facets; 3-D order 2: 372 cells, 160 sub-facets; tests/golden/<fixture>, made by make_amr_golden.py from the
    reference's own regrid).  Both entries into the CUDA path -- the oracle's arrays through the C ABI, and the C++ host
    (non-conforming topology, merged sides, psiRef/psiCor) from the case directory -- against the oracle AND the
    reference binary's dump after 20 steps; mass conserved to rounding by the scatter/gather pair."""
    import shutil

    from nebulasem_b200 import host
    from oracle import case as ocase
    # 3-D cubic orders run the persistent v4 sweeps with the FM_MORTAR branch; NSEM_MORTAR_V1=1 keeps the plain-load sweeps (the 2-D case
    # always runs those)
    monkeypatch.setenv("NSEM_MORTAR_V1", "1" if sweeps == "v1" else "0")
    src = os.path.join(os.path.dirname(__file__), "golden", fixture)
    d = str(tmp_path / fixture)
    shutil.copytree(src, d)
    gold = np.load(os.path.join(src, "expected.npz"))
    nsteps = int(gold["nsteps"])
    orc = ocase.load_case(d, exact_order=False)
    assert len(orc.mortar_faces) == nmortar
    nb = orc.gB
    mass0 = float((orc.rho[:nb] * orc.g.cV[:nb]).sum())
    ctx = device_from_oracle(orc)
    assert "mortar" in ctx.kernel_info and ctx.kernel_info.startswith(sweeps), ctx.kernel_info
    s = host.Solver.open_case(d)
    s.attach(0)
    ctx.step(nsteps)
    s.step(nsteps)
    s.download()
    orc.run(nsteps)
    c0 = np.sqrt(orc.gamma * orc.R * orc.p.T0)
    for tag, (rho, U, T) in (("c-abi", ctx.download_state()[:3]), ("host", s.state()[:3])):
        err = conserved_errors(orc, rho, U, T)
        print("MORTAR_PARITY", fixture, sweeps, tag, err, "vs reference dump: rho %.2e" % (np.linalg.norm(rho[:nb] - gold["rho"]) / np.linalg.norm(gold["rho"])))
        assert np.isfinite(rho).all() and np.isfinite(U).all() and np.isfinite(T).all()
        assert err["rho"] <= TOL and err["rhoTheta"] <= TOL and err["rhoU_scaled"] <= TOL
        # the reference binary's own dump
        assert np.linalg.norm(rho[:nb] - gold["rho"]) <= TOL * np.linalg.norm(gold["rho"])
        th, th_ref = rho[:nb] * (T[:nb] + orc.p.T0), gold["rho"] * (gold["T"] + orc.p.T0)
        assert np.linalg.norm(th - th_ref) <= TOL * np.linalg.norm(th_ref)
        mom, mom_ref = rho[:nb, None] * U[:nb], gold["rho"][:, None] * gold["U"]
        assert np.linalg.norm(mom - mom_ref) <= TOL * np.linalg.norm(gold["rho"]) * c0
        mass = float((rho[:nb] * orc.g.cV[:nb]).sum())
        assert abs(mass - mass0) <= 1e-13 * abs(mass0)
    ctx.close()
    s.close()


def test_pipelined_transfers_match_blocking_ones():
    """nsem_upload_state_async / nsem_download_state_async (copies on their own streams, one staging buffer per direction, the download of a
    batch overlapping the upload of the next) return bit for bit what the blocking entry points return, batch after batch."""
    from nebulasem_b200 import host
    a = host.Solver.synthetic("bubble3d", 4, 3, 3, 4)
    b = host.Solver.synthetic("bubble3d", 4, 3, 3, 4)
    a.attach(0)
    b.attach(0)
    a.upload(); a.step(3); a.download()
    ref = [x.copy() for x in a.state()]
    for _ in range(4):                      # every batch restarts from the same host input
        b.upload_async()
        b.step(3)
        b.download_async()
    b.sync()
    out = b.state_out()
    for name, r, o in zip(("rho", "U", "T", "p"), ref, out):
        assert np.array_equal(r, o), name
    # and the blocking path still works after the pipelined one on the same context
    b.upload(); b.step(3); b.download()
    for name, r, o in zip(("rho", "U", "T", "p"), ref, b.state()):
        assert np.array_equal(r, o), name
    a.close()
    b.close()


@pytest.mark.parametrize("decomp", ["METIS", "XYZ", "METIS-periodic"])
def test_two_partitions_equal_one_partition(decomp):
    """One METIS/XYZ partition per GPU with the NCCL face-trace halo == the single-partition run.  METIS-periodic: the doubly
    periodic isentropic vortex -- the owner cells of paired CYCLIC faces are contracted before METIS, so the periodic copy stays a local read."""
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "mp_gpu_check.py"), decomp]
    env = dict(os.environ)
    if decomp == "METIS-periodic":
        cmd[-1] = "METIS"
        env["MP_CHECK_KIND"] = "vortex"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0 and "MP_CHECK_OK" in out.stdout


def test_diagnostics_match_oracle(tmp_cases):
    """nsem_diagnostics (Courant, mass, energy, volume; euler.cpp:261-283) against the oracle's sums."""
    from nebulasem_b200 import host
    nsteps = 8
    orc = make_oracle(tmp_cases, "hill3d", nsteps, exact=False, nx=6, ny=2, nz=4, order=3)
    orc.run(nsteps)
    s = host.Solver.open_case(orc.case_dir)
    s.attach(0)
    s.step(nsteps)
    d = s.diagnostics()
    mass, energy, volume = orc.diagnostics_sums(orc.T + orc.p.T0)
    nb = orc.gB
    co = np.sqrt((orc.U[:nb] ** 2).sum(axis=1)) * orc.p.dt / orc.g.cV[:nb] ** (1.0 / 3)
    assert abs(d["mass"] - mass) <= 1e-12 * abs(mass)
    assert abs(d["energy"] - energy) <= 1e-12 * abs(energy)
    assert abs(d["volume"] - volume) <= 1e-12 * abs(volume)
    assert abs(d["courant_max"] - co.max()) <= 1e-12 * co.max()
    assert abs(d["courant_avg"] - co.mean()) <= 1e-12 * co.mean()
    assert abs(d["mass_loss"]) < 1e-12          # the scheme conserves mass to rounding (reference prints ~1e-15)
    s.close()


@pytest.mark.parametrize("order", [2, 3, 4, 5, 6, 7])
def test_isentropic_vortex_orders_2_to_7(tmp_cases, order):
    """BASELINE configs[2]: examples/isentropic (CYCLIC, no viscosity, no buoyancy), orders 2-7, 15 steps vs the oracle.
    Velocities are O(1) here, so rho*U is held to 1e-11 SELF-relative as well."""
    nsteps = 15
    orc = make_oracle(tmp_cases, "vortex", nsteps, exact=False, n=5, order=order)
    ctx = device_from_oracle(orc)
    ctx.step(nsteps)
    orc.run(nsteps)
    rho, U, T, p = ctx.download_state()
    err = conserved_errors(orc, rho, U, T)
    print(order, err)
    assert err["rho"] <= TOL and err["rhoTheta"] <= TOL and err["rhoU_self"] <= TOL
    ctx.close()


@pytest.mark.parametrize("order", [1, 2, 3, 5, 6, 7])
def test_bubble3d_other_orders(tmp_cases, order):
    """3-D orders other than 4 (v4 kernels with their per-order thread/shared-memory configuration)."""
    nsteps = 6
    orc = make_oracle(tmp_cases, "bubble3d", nsteps, exact=False, n=2, order=order)
    ctx = device_from_oracle(orc)
    ctx.step(nsteps)
    orc.run(nsteps)
    rho, U, T, p = ctx.download_state()
    err = conserved_errors(orc, rho, U, T)
    print(order, err)
    assert err["rho"] <= TOL and err["rhoTheta"] <= TOL and err["rhoU_scaled"] <= TOL
    ctx.close()


def test_kernel_generations_agree_and_conserve_mass_at_scale():
    """Size-independent properties on a mesh far beyond what the oracle can run (48^3 order 4 = 13.8 M nodes):
    the kernel generations (plain loads / bulk-async staged / warp-per-element / persistent pipelined with metrics on the
    fly or stored) agree to rounding, the result does
    not depend on the element schedule, and mass is conserved to 1e-13 (the reference prints ~1e-15 losses)."""
    import subprocess
    import sys
    code = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from nebulasem_b200 import host
s = host.Solver.synthetic("bubble3d", 48, 48, 48, 4)
s.attach(0)
d0 = s.diagnostics()
s.step(10)
d = s.diagnostics()
d["mass_loss"] = (d0["mass"] - d["mass"]) / d0["mass"]      # same device reduction order before and after
d["volume_loss"] = (d0["volume"] - d["volume"]) / d0["volume"]
s.download()
rho, U, T, p = s.state()
nb = s.gBCSfield
np.save(sys.argv[1], np.concatenate([rho[:nb, None], U[:nb], T[:nb, None]], axis=1))
print("KERNELS", s.kernel_info)
print("MASS_LOSS", d["mass_loss"], "VOLUME_LOSS", d["volume_loss"])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    outs = {}
    with tempfile.TemporaryDirectory() as td:
        for tag, env in (("v1", {"NSEM_KERNELS": "v1"}), ("v2", {"NSEM_KERNELS": "v2"}),
                         ("v2m", {"NSEM_KERNELS": "v2", "NSEM_SCHEDULE": "morton"}), ("v4", {}), ("v4m", {"NSEM_SCHEDULE": "morton"}),
                         ("v4s", {"NSEM_METRICS": "stored"}),
                         ("v4r", {"NSEM_RUN": "32"})):
            f = os.path.join(td, tag + ".npy")
            r = subprocess.run([sys.executable, "-c", code, f], env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
            assert r.returncode == 0, r.stderr[-2000:]
            want = {"v1": "v1", "v2": "v2", "v2m": "v2", "v4": "on the fly", "v4m": "on the fly", "v4s": "stored metrics", "v4r": "on the fly"}[tag]
            assert want in r.stdout.split("KERNELS")[1].splitlines()[0], (tag, r.stdout)
            loss = float(r.stdout.split("MASS_LOSS")[1].split()[0])
            assert abs(loss) <= 1e-13, (tag, loss)
            outs[tag] = np.load(f)
    ref = outs["v1"]
    assert np.isfinite(ref).all()
    assert np.array_equal(outs["v2"], outs["v2m"])            # the schedule never changes the result
    assert np.array_equal(outs["v4"], outs["v4m"])
    assert np.array_equal(outs["v4"], outs["v4r"])            # nor do the runs of consecutive elements one CTA handles (RunIter)
    for tag in ("v2", "v4", "v4s"):
        a = outs[tag]
        assert np.linalg.norm(a[:, 0] - ref[:, 0]) / np.linalg.norm(ref[:, 0]) <= 1e-13
        assert np.linalg.norm(a[:, 0] * (a[:, 4] + 300) - ref[:, 0] * (ref[:, 4] + 300)) / np.linalg.norm(ref[:, 0] * (ref[:, 4] + 300)) <= 1e-13
        assert np.linalg.norm(a[:, 0:1] * a[:, 1:4] - ref[:, 0:1] * ref[:, 1:4]) / (np.linalg.norm(ref[:, 0]) * 347.0) <= 1e-13
