"""Run the compiled reference (oracle/_ref/<variant>/euler) on a case directory -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import subprocess
import time

from . import refio

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_bin(name: str, variant: str = "parity") -> str:
    return os.path.join(HERE, "_ref", variant, name)


def have_ref(variant: str = "parity") -> bool:
    return os.path.exists(ref_bin("euler", variant))


def run_euler(case_dir: str, variant: str = "parity", threads: int | None = None, timeout: float | None = None):
    """Runs `euler ./controls` in case_dir; returns (wall seconds, stdout)."""
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    t0 = time.time()
    out = subprocess.run([ref_bin("euler", variant), "./controls"], cwd=case_dir, env=env, capture_output=True,
                         text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(f"reference euler failed in {case_dir}:\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
    return time.time() - t0, out.stdout


def read_dump(case_dir: str, index: int = 1) -> dict:
    """rho/U/T/p of dump <index> (= step / write_interval) as arrays over the real nodes."""
    out = {}
    for n in ("rho", "U", "T", "p"):
        v = refio.read_field_values(os.path.join(case_dir, f"{n}{index}"))
        out[n] = v[:, 0] if v.shape[1] == 1 else v
    return out
