"""The drop-in `euler ./controls` binary (nebulasem_b200/csrc/host/euler_main.cpp), one process per partition.

CPU part (NSEM_DRYRUN: set-up only, no GPU): N processes decompose the same case, read their share of the field files,
dump per rank into grid<r>/ and rank 0 merges -- the merged dump must be bit-equal to the single-process dump.
GPU part: the binary against the oracle on one GPU, and 2 processes == 1 process on two GPUs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from nebulasem_b200 import build
from oracle import case as ocase
from oracle import cases as ocases
from oracle import refio
from tests.helpers import conserved_errors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_euler(case_dir, world, dry=None, timeout=600):
    procs = []
    for r in range(world):
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"):
            env.pop(k, None)
        if dry is not None:
            env["NSEM_DRYRUN"] = str(dry)
        if world > 1:
            env.update(NSEM_RANK=str(r), NSEM_WORLD=str(world))
        procs.append(subprocess.Popen([build.EULER_BIN, "./controls"], cwd=case_dir, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=timeout)
        assert p.returncode == 0, (out[-1500:], err[-1500:])


def make_case(tmp_path, name):
    a = str(tmp_path / (name + "_1"))
    if name.endswith("_amr") or "/" in name:
        a = str(tmp_path / (name.replace("/", "_") + "_1"))
        shutil.copytree(os.path.join(ROOT, "tests", "golden", name), a)
        for f in ("expected.npz", "geom_sha256.json"):
            if os.path.exists(os.path.join(a, f)):
                os.remove(os.path.join(a, f))
    else:
        ocases.CASES[name](n=4, order=2).write(a, 5)
    return a


@pytest.mark.parametrize("name,world", [("bubble3d", 2), ("bubble3d", 3), ("srtb3d_amr", 2), ("vortex", 2), ("vortex", 3),
                                        ("sphere/hydro-sphere", 2), ("sphere/hydro-sphere", 5), ("sphere/acoustic-sphere-regridded", 3)])
def test_partitioned_setup_merges_to_the_single_process_fields(tmp_path, name, world):
    """... and on the cubed sphere: projection with the whole grid's cube extremes, radial gravity and reference state per part, rho without
    a file, 2:1 faces kept inside a part -- the merged set-up dump equals the one-process dump bit for bit."""
    a = make_case(tmp_path, name)
    b = str(tmp_path / (name.replace("/", "_") + "_n"))
    shutil.copytree(a, b)
    run_euler(a, 1, dry=7)
    run_euler(b, world, dry=7)
    for f in ("rho7", "U7", "T7", "p7"):
        va, vb = refio.read_field_values(os.path.join(a, f)), refio.read_field_values(os.path.join(b, f))
        assert va.shape == vb.shape and np.array_equal(va, vb), f
        fa, fb = refio.read_field(os.path.join(a, f)), refio.read_field(os.path.join(b, f))
        assert [(x.patch, x.kind) for x in fa.bcs] == [(x.patch, x.kind) for x in fb.bcs]       # no interMesh_* patch in the merged file
    assert sorted(d for d in os.listdir(b) if d.startswith("grid") and os.path.isdir(os.path.join(b, d))) == [f"grid{r}" for r in range(world)]


@pytest.mark.gpu
def test_euler_binary_matches_oracle(tmp_path):
    """`euler ./controls` on one GPU: 5 steps of the 4^3 order-2 bubble, dump 1 against the oracle."""
    a = make_case(tmp_path, "bubble3d")
    orc = ocase.load_case(a, exact_order=False)
    run_euler(a, 1)
    orc.run(5)
    rho, U, T = (refio.read_field_values(os.path.join(a, f + "1")) for f in ("rho", "U", "T"))
    assert rho.shape[0] == orc.gB                       # dumps hold the real nodes
    err = conserved_errors(orc, rho[:, 0], U, T[:, 0])
    print(err)
    assert err["rho"] <= 1e-11 and err["rhoTheta"] <= 1e-11 and err["rhoU_scaled"] <= 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bubble3d", "srtb3d_amr"])
def test_two_processes_equal_one_process(tmp_path, name):
    """Two `euler` processes (METIS halves, NCCL halo, id through the case directory) write the same merged dump as one
    process, bit for bit -- also on the non-conforming mesh, whose mortar faces the decomposition keeps whole."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    a = make_case(tmp_path, name)
    b = str(tmp_path / (name + "_n"))
    shutil.copytree(a, b)
    run_euler(a, 1)
    run_euler(b, 2)
    idx = 1
    for f in ("rho", "U", "T", "p"):
        va, vb = refio.read_field_values(os.path.join(a, f + str(idx))), refio.read_field_values(os.path.join(b, f + str(idx)))
        assert va.shape == vb.shape and np.array_equal(va, vb), f


@pytest.mark.gpu
def test_euler_binary_runs_the_amr_cycle(tmp_path):
    """`euler ./controls` with amr_step (BASELINE configs[4] in small): regrid before step 1 and after every dump, in memory, the state
    transferred on the device; every dump matches the grid of its regrid, mass stays what it was (1e-12)."""
    a = str(tmp_path / "bubble3d_amr_cycle")
    ocases.CASES["bubble3d"](n=4, order=2).write(a, 15)
    ctl = open(os.path.join(a, "controls")).read()
    ctl = ctl.replace("write_interval 15", "write_interval 5\n    amr_step 1")
    ctl += "refinement\n{\n    direction 0 0 0\n    field T\n    field_min 0.1\n    field_max 0.4\n    max_level 2\n    limit 100000\n}\n"
    open(os.path.join(a, "controls"), "w").write(ctl)
    run_euler(a, 1)
    from nebulasem_b200 import host
    masses = []
    for dump, grid in ((1, 0), (2, 1), (3, 2)):          # dump k was computed on the grid written at regrid k-1
        assert os.path.exists(os.path.join(a, f"grid_{grid}.bin"))        # write_format BINARY (the default): the regridded grids are .bin
        rho = refio.read_field_values(os.path.join(a, f"rho{dump}"))[:, 0]
        T = refio.read_field_values(os.path.join(a, f"T{dump}"))[:, 0]
        case = str(tmp_path / f"check{dump}")
        os.makedirs(case)
        shutil.copy(os.path.join(a, f"grid_{grid}.bin"), os.path.join(case, "grid_0.bin"))
        for f in ("rho", "U", "T", "p"):
            shutil.copy(os.path.join(a, f"{f}{dump}.bin"), os.path.join(case, f"{f}0.bin"))
        open(os.path.join(case, "controls"), "w").write(ctl.replace("amr_step 1", ""))
        s = host.Solver.open_case(case)
        assert rho.shape[0] == s.gBCSfield and np.isfinite(rho).all() and np.isfinite(T).all()
        masses.append(float((rho * s.f64("cV")[:s.gBCSfield]).sum()))
        if dump == 1:
            assert s.nBCS == 372                        # the reference's initial regrid of this case refines the same 44 cells
        s.close()
    assert max(abs(m - masses[0]) for m in masses) <= 1e-12 * abs(masses[0]), masses


def test_amr_controls_are_accepted_on_several_partitions(tmp_path):
    """amr_step with more than one process: round 1 refused it; now the partitions are set up (the regrids run through the whole-domain
    solver every rank keeps, run_case -- the -m gpu test_amr_run_matches_the_reference_run[2] runs the cycle).  CPU part: the set-up of both
    ranks succeeds with and without NSEM_IGNORE_AMR_STEP, and a partition carries no forest of its own."""
    a = str(tmp_path / "amr_two_ranks")
    ocases.CASES["bubble3d"](n=4, order=2).write(a, 5)
    ctl = open(os.path.join(a, "controls")).read().replace("end_step", "amr_step 1\n    end_step", 1)
    open(os.path.join(a, "controls"), "w").write(ctl)
    for extra in ({}, {"NSEM_IGNORE_AMR_STEP": "1"}):
        env = dict(os.environ, NSEM_WORLD="2", NSEM_DRYRUN="1", **extra)
        procs = [subprocess.Popen([build.EULER_BIN, "./controls"], cwd=a, env=dict(env, NSEM_RANK=str(r)), stdout=subprocess.PIPE,
                                  stderr=subprocess.PIPE, text=True) for r in range(2)]
        for p in procs:
            o, e = p.communicate(timeout=120)
            assert p.returncode == 0, e[-500:]


def test_restart_reads_the_dump_start_step_names_and_the_newest_older_grid(tmp_path):
    """start_step counts time steps: the run resumes from dump start_step / write_interval (AmrIteration, iteration.h:102) on the newest
    grid <mesh>_<k>, k <= dump (findLastRefinedGrid, field.cpp:79-91) -- a fixed-mesh case only has grid_0."""
    gold = os.path.join(ROOT, "tests", "golden", "vtk", "bubble3d_n2_o2")            # controls + grid_0.txt + the reference's dump 1
    a = str(tmp_path / "restart")
    shutil.copytree(gold, a)
    ctl = open(os.path.join(a, "controls")).read()
    assert "start_step 0" in ctl and "write_interval 4" in ctl
    open(os.path.join(a, "controls"), "w").write(ctl.replace("start_step 0", "start_step 4").replace("end_step 4", "end_step 8"))
    run_euler(a, 1, dry=9)
    for f in ("U", "T"):
        assert np.array_equal(refio.read_field_values(os.path.join(a, f + "9")), refio.read_field_values(os.path.join(a, f + "1"))), f
    # the first cycle of a run recomputes rho from p and T even on a restart (ait.start(), euler.cpp:136-149), and the dumped p is the one the
    # last step started from (euler.cpp:211): rho comes back to ~1e-10, not bitwise -- the reference's own restart behaviour
    r9, r1 = refio.read_field_values(os.path.join(a, "rho9")), refio.read_field_values(os.path.join(a, "rho1"))
    assert 0 < np.abs(r9 - r1).max() <= 1e-8 * np.abs(r1).max()
    # a dump that does not exist is an error, not a silent fresh start
    open(os.path.join(a, "controls"), "w").write(ctl.replace("start_step 0", "start_step 8").replace("end_step 4", "end_step 12"))
    out = subprocess.run([build.EULER_BIN, "./controls"], cwd=a, env=dict(os.environ, NSEM_DRYRUN="9"), capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "rho2" in out.stderr, out.stderr[-300:]


@pytest.mark.gpu
@pytest.mark.parametrize("case,world", [("srtb-amr", 1), ("srtb-amr", 2), ("acoustic-sphere-amr-dg", 1), ("acoustic-sphere-amr-dg", 2)])
def test_amr_run_matches_the_reference_run(tmp_path, case, world):
    """BASELINE configs[4]: examples/atmo/srtb-amr exactly as it ships (amr_step 1), write_interval 50, 100 steps, through the drop-in
    `euler ./controls` on one GPU against the dump the UNMODIFIED reference binary wrote of the same run (tests/golden/amr_run/, made by
    make_amr_run_golden.py).  Both do what iteration.h:94-147 + euler.cpp:57-287 prescribe: start-branch set-up on the coarse grid, regrid
    (same 32 cells tagged: 100 -> 196), restart branch on the regridded grid, 50 steps, dump, regrid by the indicator of the dumped state
    (38 refined, 60 coarsened: 265 cells), 50 steps, dump.  The in-memory forest numbers cells and local frames differently from the
    reference's facet splitting, so cells are matched by centroid and nodes by position; then rho, rho*theta <= 1e-11 and rho*U <= 1e-11 of
    ||rho|| c0 (north_star), on a mesh with non-conforming faces, after two regrids and their field transfers."""
    from nebulasem_b200 import host
    # acoustic-sphere-amr-dg: the same on a cubed sphere (examples/atmo/acoustic-sphere-amr-dg at order 2, 20 steps: 600 -> 672 -> 612 cells;
    # 2-D refinement that never splits a cell's radial axis, tags weighed with the un-projected mesh's volumes as Prepare::refineMesh does)
    src = os.path.join(ROOT, "tests", "golden", "amr_run", case)
    a = str(tmp_path / case)
    shutil.copytree(src, a)
    exp = np.load(os.path.join(a, "expected.npz"))
    os.remove(os.path.join(a, "expected.npz"))
    if world > 1:
        # SURVEY 8(f)2 on several partitions: one process per GPU, the regrids through the whole-domain solver every rank keeps (run_case:
        # states summed over the ranks, the one-partition regrid, parts cut from the emitted grid), merged dumps against the same reference run
        import torch
        if torch.cuda.device_count() < world:
            pytest.skip(f"needs {world} GPUs")
    run_euler(a, world, timeout=900)
    nsteps, interval = int(exp["nsteps"]), int(exp["interval"])
    scale = np.abs(exp["node_xyz"]).max()
    T0, c0 = 300.0, np.sqrt(1004.67 / 715.5 * (1004.67 - 715.5) * 300.0)
    rel = lambda x, y, sc=None: float(np.linalg.norm((x - y).ravel()) / (np.linalg.norm(y.ravel()) if sc is None else sc))
    # half way: dump 1 = 50 steps on the grid of the initial regrid, against the reference run stopped there (its full run overwrites the
    # files of dump 1 with the fields transferred to the next grid)
    h = str(tmp_path / "half")
    os.makedirs(h)
    shutil.copy(os.path.join(a, "grid_0.bin"), os.path.join(h, "grid_0.bin"))
    for f in ("rho", "U", "T", "p"):
        shutil.copy(os.path.join(a, f"{f}1.bin"), os.path.join(h, f"{f}0.bin"))
    import re
    open(os.path.join(h, "controls"), "w").write(re.sub(r"(?m)^\s*amr_step\s+\d+\s*\n", "", open(os.path.join(a, "controls")).read()))
    s = host.Solver.open_case(h)                                  # for the geometry of that grid only: the set-up's start branch would
    n = s.gBCSfield                                               # recompute rho from the dumped p (which step 50 formed with theta of step 49)
    rho, U, T = (refio.read_field_values(os.path.join(a, f"{f}1")) for f in ("rho", "U", "T"))
    rho, T = rho[:, 0], T[:, 0]
    kx = lambda x: np.round(x / scale * 1e7).astype(np.int64)
    # DG nodes on element faces are duplicated: a node is identified by its position AND the centroid of its cell
    mine = np.concatenate([kx(np.repeat(s.f64("gCC")[:3 * s.nBCS].reshape(-1, 3), s.NP, axis=0)), kx(s.f64("cC").reshape(-1, 3)[:n])], axis=1)
    s.close()
    assert n == exp["half_node_xyz"].shape[0], (n, exp["half_node_xyz"].shape)
    ref = np.concatenate([kx(np.repeat(exp["grid0_CC"], int(exp["NP"]), axis=0)), kx(exp["half_node_xyz"])], axis=1)
    om, orf = np.lexsort(mine.T[::-1]), np.lexsort(ref.T[::-1])
    assert np.array_equal(mine[om], ref[orf]), "the initial regrids differ"
    r_m, U_m, T_m, r_r, U_r, T_r = rho[:n][om], U[:n][om], T[:n][om], exp["half_rho"][orf], exp["half_U"][orf], exp["half_T"][orf]
    err = dict(rho=rel(r_m, r_r), rhoTheta=rel(r_m * (T_m + T0), r_r * (T_r + T0)),
               rhoU_scaled=rel(r_m[:, None] * U_m, r_r[:, None] * U_r, np.linalg.norm(r_r) * c0))
    print("after 50 steps on the first regridded grid:", err)
    assert err["rho"] <= 1e-11 and err["rhoTheta"] <= 1e-11 and err["rhoU_scaled"] <= 1e-11, err
    s = host.Solver.open_case(a, nsteps // interval)             # grid of the last regrid + the last dump
    nb, NP = s.nBCS, s.NP
    assert nb == exp["grid1_CC"].shape[0], (nb, exp["grid1_CC"].shape)
    n = s.gBCSfield
    rho, U, T = (refio.read_field_values(os.path.join(a, f"{f}{nsteps // interval}")) for f in ("rho", "U", "T"))
    rho, T = rho[:, 0], T[:, 0]
    xyz = s.f64("cC").reshape(-1, 3)[:n]
    cc = np.repeat(s.f64("gCC")[:3 * nb].reshape(nb, 3), NP, axis=0)
    s.close()
    key = lambda c, x: np.round(np.concatenate([c, x], axis=1) / scale * 1e7).astype(np.int64)
    mine = key(cc, xyz)
    ref = key(np.repeat(exp["grid1_CC"], int(exp["NP"]), axis=0), exp["node_xyz"])
    om, orf = np.lexsort(mine.T[::-1]), np.lexsort(ref.T[::-1])
    assert np.array_equal(mine[om], ref[orf]), "the two runs do not end on the same grid"
    r_m, U_m, T_m = rho[:n][om], U[:n][om], T[:n][om]
    r_r, U_r, T_r = exp["rho"][orf], exp["U"][orf], exp["T"][orf]
    err = dict(rho=rel(r_m, r_r), rhoTheta=rel(r_m * (T_m + T0), r_r * (T_r + T0)),
               rhoU_scaled=rel(r_m[:, None] * U_m, r_r[:, None] * U_r, np.linalg.norm(r_r) * c0))
    print("AMR run vs the reference run:", err)
    d = np.abs(r_m - r_r)
    worst = np.argsort(d)[-5:]
    print("rho: nodes with |diff| > 1e-12:", int((d > 1e-12).sum()), "of", d.size, "worst:", [(float(d[i]), exp["node_xyz"][orf][i].round(2).tolist(), float(exp["node_cV"][orf][i])) for i in worst])
    dm = float((r_m * exp["node_cV"][orf]).sum() - (r_r * exp["node_cV"][orf]).sum())
    print("mass difference", dm, "of", float((r_r * exp["node_cV"][orf]).sum()), "mean diff", float((r_m - r_r).mean()), "T diff max", float(np.abs(T_m - T_r).max()))
    assert err["rho"] <= 1e-11 and err["rhoTheta"] <= 1e-11 and err["rhoU_scaled"] <= 1e-11, err
