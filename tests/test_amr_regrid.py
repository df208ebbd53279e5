"""Adaptive regrid in memory (nebulasem_b200/csrc/host/amr.cpp): own octree regrid + tagging against the reference's regrids.

The regrid is not a restatement of MeshObject::refineMesh (cell order and local frames differ), so grids are compared as SETS of cells
(centroid, volume), facet and mortar-face counts, and fields through cell/node coordinates:
  * tests/golden/refine_field/<case>/stage{0,1,2}: two regrids each of a 2-D and a 3-D case driven through the reference's refineMesh;
  * tests/golden/srtb3d_amr, srtb_amr: the reference's own initial regrid (its tagging) of examples/atmo/srtb-3d (4^3) and srtb-amr.
"""
import os
import shutil

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
RF = os.path.join(GOLD, "refine_field")
CASES = [("3d_o2", (0, 0, 0)), ("2d_o4", (0, 0, 1))]


@pytest.fixture(autouse=True)
def _amr_env(monkeypatch):
    monkeypatch.setenv("NSEM_AMR", "1")          # solvers created in these tests keep an AMR forest


def cell_signature(s):
    n = s.nBCS
    a = np.concatenate([s.f64("gCC")[:3 * n].reshape(n, 3), s.f64("gCV")[:n, None]], axis=1)
    return a[np.lexsort(np.round(a, 6).T[::-1])]


def same_cells(a, b):
    assert a.nBCS == b.nBCS and a.nFacets == b.nFacets
    assert int((a.u32("faceMortar") > 0).sum()) == int((b.u32("faceMortar") > 0).sum())
    sa, sb = cell_signature(a), cell_signature(b)
    assert np.abs(sa - sb).max() <= 1e-12 * np.abs(sb).max()


def match_cells(mine, ref, ids):
    """cells of `mine` at the centroids of cells `ids` of `ref`"""
    cm = mine.f64("gCC")[:3 * mine.nBCS].reshape(-1, 3)
    cr = ref.f64("gCC").reshape(-1, 3)
    out = []
    for i in ids:
        d = np.abs(cm - cr[i]).sum(axis=1)
        j = int(np.argmin(d))
        assert d[j] <= 1e-6
        out.append(j)
    return out


def regrid_like_the_reference(name, direction, k, s, ref_prev):
    """apply pass k of the golden case to solver s (whose cells are those of ref_prev in another order)"""
    from nebulasem_b200 import host
    g = np.load(os.path.join(RF, f"{name}_pass{k}.npz"))
    r = np.zeros(s.nBCS, np.uint8)
    c = np.zeros(s.nBCS, np.uint8)
    r[match_cells(s, ref_prev, g["rCells"])] = 1
    c[match_cells(s, ref_prev, np.nonzero(g["cCells"])[0])] = 1
    s.regrid(r, c)
    ref = host.Solver.open_case(os.path.join(RF, name, f"stage{k}"))
    return g, ref


@pytest.mark.parametrize("name,direction", CASES)
def test_regrid_produces_the_cells_of_the_reference_regrid(name, direction):
    from nebulasem_b200 import host
    s = host.Solver.open_case(os.path.join(RF, name, "stage0"))
    s.enable_amr(direction=direction)
    prev = host.Solver.open_case(os.path.join(RF, name, "stage0"))
    for k in (1, 2):
        g, ref = regrid_like_the_reference(name, direction, k, s, prev)
        same_cells(s, ref)
        # the maps describe the regrid the way MeshObject::refineMesh does: every new cell is copied, split from or merged into
        rm, cmap, km = s.u32("refineMap").astype(np.int64), s.u32("cellMap").astype(np.int64), s.u32("coarseMap").astype(np.int64)
        assert len(rm) == len(g["refineMap"]) and len(km) == len(g["coarseMap"])
        covered = np.zeros(s.nBCS, bool)
        n_old = prev.nBCS
        covered[cmap[:n_old][cmap[:n_old] != (1 << 31)]] = True
        i = 0
        while i < len(rm):
            covered[cmap[rm[i + 2:i + 2 + rm[i]]]] = True
            i += rm[i] + 2
        i = 0
        while i < len(km):
            covered[cmap[km[i + 1]]] = True
            i += km[i] + 2
        assert covered.all()
        lv = s.cell_levels()
        assert lv.max() == k and lv.min() == 0
        prev.close()
        prev = ref
    s.close()
    prev.close()


def test_tagging_reproduces_the_reference_initial_regrid_3d():
    """refinement{field T, field_min 0.1, field_max 0.4, max_level 2} on the 4^3 bubble: the reference's own initial regrid
    (tests/golden/srtb3d_amr, made by make_amr_golden.py) refines the same 44 cells."""
    from nebulasem_b200 import host
    s = host.Solver.synthetic("bubble3d", 4, 4, 4, 2)
    s.enable_amr(direction=(0, 0, 0), field="T", field_min=0.1, field_max=0.4, max_level=2, buffer_zone=2)
    s.regrid()
    ref = host.Solver.open_case(os.path.join(GOLD, "srtb3d_amr"))
    assert s.nBCS == 372
    same_cells(s, ref)
    s.close()
    ref.close()


FIELD_TXT = """size {comps}
internal 1
{{
    {init}
}}
boundary 3
{{
    top {{
        type {bc}
    }}
    bottom {{
        type {bc}
    }}
    sides {{
        type {bc}
    }}
}}
"""


def test_tagging_reproduces_the_reference_initial_regrid_2d(tmp_path):
    """examples/atmo/srtb-amr (x-y plane, refinement{direction 0 0 1, field T, 0.4 / 0.5, max_level 2}): the reference's initial regrid
    (tests/golden/srtb_amr) refines 32 of the 100 cells; the case is rebuilt here from the uniform grid and the example's initial fields."""
    from nebulasem_b200 import host
    d = tmp_path / "srtb_amr_start"
    shutil.copytree(os.path.join(RF, "2d_o4", "stage0"), d)
    for f in ("rho0.bin", "U0.bin", "T0.bin", "p0.bin"):
        os.remove(d / f)
    (d / "rho0.txt").write_text(FIELD_TXT.format(comps=1, init="uniform 0", bc="NEUMANN"))
    (d / "p0.txt").write_text(FIELD_TXT.format(comps=1, init="uniform 0", bc="NEUMANN"))
    (d / "U0.txt").write_text(FIELD_TXT.format(comps=3, init="uniform 0 0 0", bc="SYMMETRY"))
    (d / "T0.txt").write_text(FIELD_TXT.format(comps=1, init="cosine 0 0.5    500 350 50   250 250 1000", bc="NEUMANN"))
    s = host.Solver.open_case(str(d))
    s.enable_amr(direction=(0, 0, 1), field="T", field_min=0.4, field_max=0.5, max_level=2, buffer_zone=2)
    s.regrid()
    ref = host.Solver.open_case(os.path.join(GOLD, "srtb_amr"))
    assert s.nBCS == 196
    same_cells(s, ref)
    s.close()
    ref.close()


@pytest.mark.parametrize("case,direction,max_level", [("bubble3d", (0, 0, 0), 2), ("bubble2d", (0, 1, 0), 3)])
def test_amr_cycles_keep_the_grid_valid(case, direction, max_level):
    """Eight regrids that chase a bubble jumping around the domain (refinement and coarsening mixed, up to three levels): every grid loads
    through the topology code, the volume is that of the domain, neighbours across a face differ by at most one level, and the faces
    where they differ are exactly the mortar faces."""
    from nebulasem_b200 import host
    rng = np.random.default_rng(7)
    two_d = case == "bubble2d"
    s = host.Solver.synthetic(case, 4, 1 if two_d else 4, 4, 2)
    s.enable_amr(direction=direction, field="T", field_min=0.15, field_max=0.4, max_level=max_level, buffer_zone=1)
    vol0 = s.f64("gCV")[:s.nBCS].sum()
    seen_levels, counts = set(), []
    for cycle in range(8):
        x = s.f64("cC").reshape(-1, 3)
        ctr = np.array([200 + 600 * rng.random(), 50 if two_d else 200 + 600 * rng.random(), 200 + 600 * rng.random()])
        r = np.linalg.norm((x - ctr) * (np.array([1, 0, 1]) if two_d else 1), axis=1) / 250
        s.set_state(T=np.where(r < 1, 0.25 * (1 + np.cos(np.pi * r)), 0.0))
        s.regrid()
        lv = s.cell_levels()
        fo, fn, fm = s.u32("faceOwner"), s.u32("faceNeigh"), s.u32("faceMortar")
        inner = fn < s.nBCS
        jump = np.abs(lv[fo[inner]] - lv[fn[inner]])
        assert jump.max() <= 1 and lv.max() <= max_level
        assert int((jump == 1).sum()) == int((fm > 0).sum())
        assert abs(s.f64("gCV")[:s.nBCS].sum() - vol0) <= 1e-12 * vol0
        seen_levels.update(lv.tolist())
        counts.append(s.nBCS)
    assert 2 in seen_levels and len(set(counts)) > 3          # deeper levels were reached, cells were both added and removed
    s.close()


def test_amr_run_restarts_from_a_dump(tmp_path):
    """The forest is saved next to the grid of every regrid (<mesh>_<dump>.forest): a solver opened from that dump continues the AMR
    run exactly like the one that wrote it (same levels, same cells in the same order after the next regrid, coarsening included)."""
    from nebulasem_b200 import host
    d = tmp_path / "restart"
    shutil.copytree(os.path.join(RF, "3d_o2", "stage0"), d)

    def blob(s, ctr):
        x = s.f64("cC").reshape(-1, 3)
        r = np.linalg.norm(x - np.asarray(ctr, dtype=float), axis=1) / 250
        s.set_state(T=np.where(r < 1, 0.25 * (1 + np.cos(np.pi * r)), 0.0))

    a = host.Solver.open_case(str(d))
    a.enable_amr(direction=(0, 0, 0), field="T", field_min=0.15, field_max=0.4, max_level=2, buffer_zone=1)
    blob(a, (400, 400, 400))
    a.regrid()
    blob(a, (400, 400, 400))
    a.write(1)
    a.write_amr_grid(1)
    b = host.Solver.open_case(str(d), 1)                       # grid_1.txt + grid_1.forest + the fields of dump 1
    b.enable_amr(direction=(0, 0, 0), field="T", field_min=0.15, field_max=0.4, max_level=2, buffer_zone=1)
    assert b.nBCS == a.nBCS and np.array_equal(b.cell_levels(), a.cell_levels())
    assert np.array_equal(b.f64("gCC"), a.f64("gCC")) and np.array_equal(b.state()[2][:b.gBCSfield], a.state()[2][:a.gBCSfield])
    for s in (a, b):
        blob(s, (650, 600, 550))                                # the bubble has moved: the old families merge, new cells split
        s.regrid()
    assert a.nBCS == b.nBCS and np.array_equal(a.f64("gCC"), b.f64("gCC")) and np.array_equal(a.cell_levels(), b.cell_levels())
    assert len(a.u32("coarseMap")) > 0 and np.array_equal(a.u32("coarseMap"), b.u32("coarseMap"))
    assert np.array_equal(a.u32("refineMap"), b.u32("refineMap"))
    a.close()
    b.close()


def test_regrid_refuses_what_it_cannot_do(monkeypatch):
    from nebulasem_b200 import capi, host
    s = host.Solver.open_case(os.path.join(GOLD, "srtb3d_amr"))        # already non-conforming: cannot seed the forest
    with pytest.raises(capi.NsemError, match="no AMR forest"):
        s.regrid()
    s.close()
    s = host.Solver.synthetic("vortex", 4, 4, 1, 2)                     # CYCLIC patches pair their faces by position in the patch:
    r = np.zeros(16, np.uint8)                                          # a corner cell without its periodic images
    r[0] = 1
    with pytest.raises(capi.NsemError, match="CYCLIC"):
        s.regrid(r, np.zeros(16, np.uint8))
    s.enable_amr(direction=(0, 0, 1), field="U", field_min=0.2, field_max=0.6, max_level=1, buffer_zone=2)
    s.regrid()                                                          # tagged by the indicator, pairs stay together
    s.close()
    monkeypatch.delenv("NSEM_AMR")
    s = host.Solver.synthetic("bubble3d", 3, 3, 3, 2)                  # no forest was asked for
    with pytest.raises(capi.NsemError, match="no AMR forest"):
        s.regrid(np.zeros(27, np.uint8), np.zeros(27, np.uint8))
    s.close()


def fields_by_coordinates(mine, ref, ref_fields):
    """reference node values re-ordered to the nodes of `mine`: cells matched by centroid, nodes inside a cell by coordinates"""
    NP, n = mine.NP, mine.nBCS
    cells = match_cells(ref, mine, range(n))                             # ref cell of every cell of mine
    xm = mine.f64("cC")[:n * NP * 3].reshape(n, NP, 3)
    xr = ref.f64("cC")[:n * NP * 3].reshape(n, NP, 3)
    idx = np.empty((n, NP), dtype=np.int64)
    for c in range(n):
        d = np.abs(xm[c][:, None, :] - xr[cells[c]][None, :, :]).sum(axis=2)
        j = d.argmin(axis=1)
        assert d[np.arange(NP), j].max() <= 1e-6 and len(set(j)) == NP
        idx[c] = cells[c] * NP + j
    idx = idx.reshape(-1)
    return [np.asarray(f).reshape(ref.nBCS * NP, -1)[idx] for f in ref_fields]


@pytest.mark.parametrize("name,direction", CASES)
def test_regrid_maps_and_geometry_transfer_to_the_reference_fields(name, direction):
    """The maps and the geometry of the own regrid, fed to the (reference-pinned) numpy transfer oracle, give the fields the reference
    wrote after ITS regrid, node by node through coordinates: what nsem_refine_state receives from EulerSolver::regridded is right."""
    from nebulasem_b200 import host
    from oracle import amr
    s = host.Solver.open_case(os.path.join(RF, name, "stage0"))
    s.enable_amr(direction=direction)
    prev = host.Solver.open_case(os.path.join(RF, name, "stage0"))
    g1 = np.load(os.path.join(RF, f"{name}_pass1.npz"))
    npts = tuple(int(x) for x in g1["dims"][:3])
    NP = s.NP
    fields = [g1["pre_rho"].reshape(-1, 1), g1["pre_U"].reshape(-1, 3), g1["pre_T"].reshape(-1, 1)]
    fields = [f[:s.nBCS * NP] for f in fields]
    for k in (1, 2):
        n_old = s.nBCS
        oldCV, oldCC, cC_old = s.f64("gCV")[:n_old], s.f64("gCC")[:3 * n_old], s.f64("cC")[:n_old * NP * 3]
        g, ref = regrid_like_the_reference(name, direction, k, s, prev)
        n_new = s.nBCS
        args = (npts, s.u32("refineMap"), s.u32("coarseMap"), s.u32("cellMap"), n_new, oldCV, oldCC, s.f64("gCC")[:3 * n_new],
                s.f64("gCV")[:n_new], cC_old, [g[f"psiRef{i}"] for i in range(6)], [g[f"psiCor{i}"] for i in range(6)],
                [g[f"wgl{i}"] for i in range(3)])
        fields = [amr.refine_field(f, *args) for f in fields]
        want = fields_by_coordinates(s, ref, [g["post_rho"], g["post_U"], g["post_T"]])
        for nm, mine, w in zip(("rho", "U", "T"), fields, want):
            assert np.abs(mine - w).max() <= 1e-13 * max(np.abs(w).max(), 1e-300), (k, nm)
        prev.close()
        prev = ref
    s.close()
    prev.close()


# ---------------------------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,direction", CASES)
def test_regrid_and_device_transfer_reproduce_the_reference_fields(name, direction):
    """Own regrid + nsem_refine_state == the fields the reference wrote after ITS regrid, compared node by node through coordinates
    (cell order and local frames differ, so the sums run in another order: 1e-13 relative instead of bit-identical)."""
    from nebulasem_b200 import host
    s = host.Solver.open_case(os.path.join(RF, name, "stage0"))
    s.enable_amr(direction=direction)
    s.attach(0)
    prev = host.Solver.open_case(os.path.join(RF, name, "stage0"))
    g1 = np.load(os.path.join(RF, f"{name}_pass1.npz"))
    n = s.nBCS * s.NP
    rho, U, T, p = [x.copy() for x in s.state()]                         # stage0 has the reference's cell order: set the pre-regrid fields
    rho[:n], U[:n], T[:n], p[:n] = g1["pre_rho"].reshape(-1)[:n], g1["pre_U"].reshape(-1, 3)[:n], g1["pre_T"].reshape(-1)[:n], g1["pre_p"].reshape(-1)[:n]
    s.set_state(rho, U, T, p)
    s.upload()
    for k in (1, 2):
        g, ref = regrid_like_the_reference(name, direction, k, s, prev)
        s.download()
        nn = s.nBCS * s.NP
        want = fields_by_coordinates(s, ref, [g["post_rho"], g["post_U"], g["post_T"]])
        for nm, comps, dev, w in zip(("rho", "U", "T"), (1, 3, 1), s.state()[:3], want):
            dev = np.asarray(dev).reshape(-1, comps)[:nn]
            assert np.abs(dev - w).max() <= 1e-13 * max(np.abs(w).max(), 1e-300), (k, nm)
        prev.close()
        prev = ref
    s.close()
    prev.close()


@pytest.mark.gpu
def test_amr_cycle_on_the_device_conserves_mass():
    """regrid by the indicator -> steps -> regrid -> steps on one B200, the state never leaving the device: finite, mass as at the start
    (the transfer restores every family's integral, the fluxes on the non-conforming grid are conservative), cells follow the bubble."""
    from nebulasem_b200 import host
    s = host.Solver.synthetic("bubble3d", 4, 4, 4, 2)
    s.enable_amr(direction=(0, 0, 0), field="T", field_min=0.1, field_max=0.4, max_level=2, buffer_zone=2)
    s.attach(0)

    def mass():
        s.download()
        n = s.gBCSfield
        return float((s.state()[0][:n] * s.f64("cV")[:n]).sum())

    m0 = mass()
    cells = [s.nBCS]
    for cycle in range(3):
        s.regrid()
        cells.append(s.nBCS)
        assert abs(mass() - m0) <= 1e-12 * abs(m0), cycle
        s.step(10)
        assert abs(mass() - m0) <= 1e-12 * abs(m0), cycle
    rho, U, T, p = s.state()
    assert all(np.isfinite(x).all() for x in (rho, U, T, p))
    assert cells[1] == 372 and s.cell_levels().max() >= 1
    assert "mortar" in s.kernel_info
    s.close()


def test_cyclic_patches_follow_a_regrid(tmp_path):
    """examples/isentropic (CYCLIC in x and y): cells that touch a periodic patch are refined together with the owners of the paired faces
    (field.cpp:826-858), and after the regrid face j of a patch lies opposite face j of its neighbor patch again (the pairing
    applyExplicitBCs relies on, field.h:2662-2664).  Explicit flags that split a pair are refused before the forest is touched."""
    from nebulasem_b200 import host
    d = str(tmp_path / "isentropic")
    shutil.copytree(os.path.join(GOLD, "examples", "isentropic"), d)
    s = host.Solver.open_case(d)
    s.enable_amr(direction=(0, 0, 1), field="T", field_min=0.15, field_max=0.4, max_level=2, buffer_zone=1)
    fc = lambda: s.f64("faceCenter").reshape(-1, 3)
    import re
    pairs = re.findall(r"(\w+)\s*\{\s*type\s+CYCLIC\s+neighbor\s+(\w+)", open(os.path.join(d, "U0.txt")).read())
    assert len(pairs) == 4                                       # 2 periodic directions, each stated from both sides
    n0 = s.nBCS
    # an explicit request that refines one side of a periodic pair only
    left = s.patch_faces(pairs[0][0])
    r = np.zeros(s.nBCS, np.uint8)
    r[s.u32("faceOwner")[left[0]]] = 1
    with pytest.raises(Exception, match="CYCLIC"):
        s.regrid(r, np.zeros(s.nBCS, np.uint8))
    assert s.nBCS == n0 and s.cell_levels().max() == 0           # nothing happened
    # a blob on the corner of the domain: the tagging refines it on all four periodic images
    x = s.f64("cC").reshape(-1, 3)
    lo, hi = x.min(axis=0), x.max(axis=0)
    rr = np.linalg.norm((x - np.array([lo[0], lo[1], 0.0]))[:, :2], axis=1) / (0.3 * (hi[0] - lo[0]))
    s.set_state(T=np.where(rr < 1, 0.25 * (1 + np.cos(np.pi * rr)), 0.0))
    for cycle in range(2):
        s.regrid()
        c = fc()
        for a, b in pairs:
            fa, fb = s.patch_faces(a), s.patch_faces(b)
            assert len(fa) == len(fb) > 0
            shift = c[fb] - c[fa]
            assert np.abs(shift - shift[0]).max() <= 1e-9 * np.abs(hi - lo).max(), (a, b)
            lv = s.cell_levels()
            own = s.u32("faceOwner")
            assert np.array_equal(lv[own[fa]], lv[own[fb]])
        x = s.f64("cC").reshape(-1, 3)
        rr = np.linalg.norm((x - np.array([lo[0], lo[1], 0.0]))[:, :2], axis=1) / (0.3 * (hi[0] - lo[0]))
        s.set_state(T=np.where(rr < 1, 0.25 * (1 + np.cos(np.pi * rr)), 0.0))
    lv = s.cell_levels()
    own = s.u32("faceOwner")
    assert s.nBCS > n0 and lv.max() == 2 and lv[own[s.patch_faces(pairs[0][0])]].max() >= 1     # the periodic patches were refined
    s.close()


def test_regrids_of_a_terrain_following_mesh_keep_its_volume():
    """3-D regrids (refinement and coarsening, two levels) of the non-affine hill mesh: the new vertices sit at the reference's corrected face
    centres and cell centroids, the children tile their parents -- cell volumes and the nodal quadrature volume stay what they were."""
    from nebulasem_b200 import host
    s = host.Solver.synthetic("hill3d", 6, 3, 4, 2)
    s.enable_amr(direction=(0, 0, 0), field="T", field_min=0.15, field_max=0.4, max_level=2, buffer_zone=1)
    v0, q0 = s.f64("gCV")[:s.nBCS].sum(), s.f64("cV")[:s.gBCSfield].sum()
    rng = np.random.default_rng(3)
    x = s.f64("cC").reshape(-1, 3)
    lo, hi = x.min(axis=0), x.max(axis=0)
    counts, levels = [], set()
    for cycle in range(4):
        x = s.f64("cC").reshape(-1, 3)
        ctr = lo + (hi - lo) * rng.random(3) * np.array([1, 1, 0.5])
        r = np.linalg.norm((x - ctr) / (0.3 * (hi - lo)), axis=1)
        s.set_state(T=np.where(r < 1, 0.25 * (1 + np.cos(np.pi * r)), 0.0))
        s.regrid()
        assert abs(s.f64("gCV")[:s.nBCS].sum() - v0) <= 1e-14 * v0 and abs(s.f64("cV")[:s.gBCSfield].sum() - q0) <= 1e-14 * q0, cycle
        counts.append(s.nBCS)
        levels.update(s.cell_levels().tolist())
    assert 2 in levels and min(counts[1:]) < max(counts)          # deeper levels were reached, cells were removed again
    s.close()


@pytest.mark.parametrize("case,n0,n1", [("advection-leveque-amr", 256, 412), ("advection-sphere-amr", 384, 726), ("transport-wave2d-amr", 128, 212)])
def test_initial_regrid_of_the_convection_examples_is_the_references(tmp_path, case, n0, n1):
    """The convection app's AMR cases (tests/golden/convection/*-amr, made by the reference's convection binary): the scalar is tagged from
    the rho slot where this solver keeps it; flat (LeVeque, max_level 2 + buffer_zone 2; wave2d) and on the cubed sphere (tags weighed with the
    un-projected load's volumes, radial axis never split).  Same cells as the reference's initial regrid, centroid by centroid."""
    import shutil

    from scipy.spatial import cKDTree

    from nebulasem_b200 import host
    d = str(tmp_path / case)
    shutil.copytree(os.path.join(GOLD, "convection", case), d)
    exp = np.load(os.path.join(d, "expected.npz"))
    s = host.Solver.open_case(d)
    assert s.nBCS == n0
    s.regrid()
    assert s.nBCS == n1 == exp["half_CC"].shape[0]
    cc = s.f64("gCC")[: 3 * s.nBCS].reshape(-1, 3)
    s.close()
    dist, idx = cKDTree(exp["half_CC"]).query(cc)
    assert dist.max() <= 1e-12 * max(1.0, np.abs(exp["half_CC"]).max()) and len(np.unique(idx)) == len(idx)
