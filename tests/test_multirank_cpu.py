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
