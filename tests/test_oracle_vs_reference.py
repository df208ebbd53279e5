"""Live comparison of the numpy oracle with the compiled reference (oracle/_ref, built by oracle/build_ref.sh).
Skipped where the binary is absent."""
import numpy as np
import pytest

from oracle import run_ref
from tests.helpers import make_oracle

pytestmark = pytest.mark.skipif(not run_ref.have_ref("parity"), reason="oracle/_ref/parity/euler not built")


@pytest.mark.parametrize("name,kw,nsteps", [
    ("bubble2d", dict(n=5, order=3), 15),
    ("bubble3d", dict(n=2, order=3), 6),
    ("vortex", dict(n=5, order=4), 12),
    ("hill3d", dict(nx=5, ny=1, nz=3, order=2), 8),
])
def test_bit_identical_to_reference(tmp_cases, name, kw, nsteps):
    orc = make_oracle(tmp_cases, name, nsteps, exact=True, **kw)
    run_ref.run_euler(orc.case_dir, variant="parity")
    ref = run_ref.read_dump(orc.case_dir, 1)
    orc.run(nsteps)
    nb = orc.gB
    assert np.array_equal(orc.rho[:nb], ref["rho"])
    assert np.array_equal(orc.U[:nb], ref["U"])
    assert np.array_equal(orc.T[:nb], ref["T"])
    assert np.array_equal(orc.pp[:nb], ref["p"])


def test_restart_from_a_dump_is_the_setup_on_the_dumped_fields(tmp_cases, tmp_path):
    """`start_step 4` with `write_interval 4`: the reference resumes from dump 1 on grid_0 (findLastRefinedGrid) and its first cycle
    recomputes rho from the dumped p and T (ait.start(), euler.cpp:136-149).  Its dump 2 is bit-identical to the oracle set up on the
    fields of dump 1 and run for 4 steps -- the semantics EulerSolver::read_controls / load_mesh follow."""
    import os
    import shutil

    from oracle import case as ocase
    orc0 = make_oracle(tmp_cases, "bubble3d", 4, exact=True, n=2, order=2)
    d = orc0.case_dir
    run_ref.run_euler(d, variant="parity")                                   # dump 1
    ctl = open(os.path.join(d, "controls")).read()
    assert "start_step 0" in ctl and "end_step 4" in ctl and "write_interval 4" in ctl
    open(os.path.join(d, "controls"), "w").write(ctl.replace("start_step 0", "start_step 4").replace("end_step 4", "end_step 8"))
    run_ref.run_euler(d, variant="parity")                                   # dump 2, resumed from dump 1
    ref = run_ref.read_dump(d, 2)
    d2 = str(tmp_path / "from_dump")
    os.makedirs(d2)
    shutil.copy(os.path.join(d, "controls"), d2)
    shutil.copy(os.path.join(d, "grid_0.txt"), d2)
    for f in ("rho", "U", "T", "p"):
        shutil.copy(os.path.join(d, f + "1.bin"), os.path.join(d2, f + "0.bin"))
    orc = ocase.load_case(d2, exact_order=True)
    orc.run(4)
    nb = orc.gB
    assert np.array_equal(orc.rho[:nb], ref["rho"]) and np.array_equal(orc.U[:nb], ref["U"])
    assert np.array_equal(orc.T[:nb], ref["T"]) and np.array_equal(orc.pp[:nb], ref["p"])
