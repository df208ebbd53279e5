"""N > 1 host logic on CPU: world_size-2 gloo run of the partition/halo-ordering check, plus in-process properties
of the decomposition."""
import os
import subprocess
import sys

import numpy as np
import pytest

from nebulasem_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("decomp", ["METIS", "XYZ"])
def test_gloo_two_ranks_halo_ordering(decomp):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29521" if decomp == "METIS" else "29522", os.path.join(ROOT, "tests", "mp_cpu_check.py"), decomp]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "MP_CPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("case", ["sphere/hydro-sphere", "sphere/acoustic-sphere-regridded", "srtb3d_amr"])
def test_gloo_two_ranks_halo_ordering_on_case_directories(case):
    """The same world-size-2 check on case directories cut as their controls say: the cubed sphere (panel edges, curved elements: shared
    nodes coincide to rounding), the reference's own regridded sphere (2:1 faces kept inside a part) and a flat regridded 3-D mesh."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29526", os.path.join(ROOT, "tests", "mp_cpu_check.py"), "METIS"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, MP_CPU_CASE=os.path.join(ROOT, "tests", "golden", case)))
    assert out.returncode == 0 and "MP_CPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("decomp,nparts,pxyz", [("METIS", 4, (1, 1, 1)), ("XYZ", 4, (2, 2, 1)), ("CELLID", 3, (1, 1, 1))])
def test_partitions_tile_the_mesh_and_keep_element_geometry(decomp, nparts, pxyz):
    g = host.Solver.synthetic("bubble3d", 4, 4, 4, 2)
    cC_g = g.f64("cC").reshape(-1, 3)[: g.gBCSfield].reshape(g.nBCS, -1)
    J_g = g.f64("Jinv")[: g.gBCSfield * 9].reshape(g.nBCS, -1)
    seen = np.zeros(g.nBCS, dtype=int)
    for r in range(nparts):
        p = host.Solver.synthetic_part("bubble3d", 4, 4, 4, 2, r, nparts, decomp, pxyz)
        cg = p.u32("cellGlobal")
        seen[cg] += 1
        assert np.array_equal(p.f64("cC").reshape(-1, 3)[: p.gBCSfield].reshape(p.nBCS, -1), cC_g[cg])
        assert np.array_equal(p.f64("Jinv")[: p.gBCSfield * 9].reshape(p.nBCS, -1), J_g[cg])
        for q in p.peers():
            assert len(p.patch_faces(f"interMesh_{r}_{q}")) > 0
    assert (seen == 1).all()


@pytest.mark.parametrize("fixture", ["srtb_amr", "srtb3d_amr"])
@pytest.mark.parametrize("method,nparts", [("METIS", 2), ("METIS", 3), ("METIS", 5), ("CELLID", 3)])
def test_decomposition_never_cuts_a_non_conforming_face(fixture, method, nparts):
    """Cells joined by 2:1 faces are contracted into one graph vertex before METIS (the other methods place such a cluster where its first
    cell goes), so no decomposition of a regridded mesh can separate the two sides of a mortar (field.cpp:1215-1220; the reference only
    weighs those edges 1000, which a tight balance on a few hundred cells can still cut).  partition_grid raises if a cut slips through."""
    from oracle import refio
    d = os.path.join(ROOT, "tests", "golden", fixture)
    g = refio.read_grid(os.path.join(d, "grid_0"))
    part, fmc = host.partition_grid(os.path.join(d, "grid_0"), len(g.cells), len(g.facets), nparts, method)
    assert np.count_nonzero(fmc) > 0
    sizes = np.bincount(part, minlength=nparts)
    assert sizes.sum() == len(g.cells) and (sizes > 0).all(), sizes
    # both cells of every flagged face in the same part
    owner = {}
    for c, faces in enumerate(g.cells):
        for f in faces:
            owner.setdefault(f, []).append(c)
    for f, cs in owner.items():
        if fmc[f] and len(cs) == 2:
            assert part[cs[0]] == part[cs[1]], f
