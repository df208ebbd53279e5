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
