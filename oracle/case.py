"""Load a NebulaSEM case directory into the numpy oracle -- TEST INFRASTRUCTURE ONLY.

Mirrors what the `euler` binary does before its time loop: Solver::Initialize (apps/utils/wrapper.cpp:11-58),
Mesh::LoadMesh (src/field/field.cpp:95-167), MeshField::read (src/field/field.h:1412-1586).
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import libm, refio
from .dg import Basis, Geometry
from .euler import BCSpec, EulerOracle, Params, vmag
from .mesh import MeshTopo, equal


def init_field(ff: refio.FieldFile, geo: Geometry, gravity) -> np.ndarray:
    """readInternal_ (field.h:1412-1531): analytic initialisers run over ALL nodes incl. ghost nodes."""
    gA = geo.gALL
    comps = ff.comps
    out = np.zeros((gA, comps))
    if ff.values is not None:
        out[: len(ff.values)] = ff.values
        return out[:, 0] if comps == 1 else out
    PI = 3.14159265358979323846264
    for ini in ff.initializers:
        kind = ini[0]
        if kind == "uniform":
            out += ini[1][None, :]
        elif kind in ("cosine", "cosine2", "linear", "gaussian"):
            _, value, pert, center, radius = ini
            if getattr(geo, "spherical", False):
                # centre given as (radius, latitude, longitude); great-circle distance over |radius| (field.h:1441-1444)
                sr = vmag(geo.cC)
                lat = libm.atan2_(geo.cC[:, 2], libm.sqrt_(geo.cC[:, 0] * geo.cC[:, 0] + geo.cC[:, 1] * geo.cC[:, 1]))
                lon = libm.atan2_(geo.cC[:, 1], geo.cC[:, 0])
                d = (center[0] + sr) / 2
                d = d * libm.acos_(math.sin(center[1]) * libm.sin_(lat) + math.cos(center[1]) * libm.cos_(lat) * libm.cos_(center[2] - lon))
                Rr = d / float(vmag(radius[None, :])[0])
            else:
                Rr = vmag((geo.cC - center[None, :]) / radius[None, :])
            if kind == "gaussian":
                v = libm.exp_(-Rr * Rr)
                v = np.array([0.0 if equal(float(x), 0.0) else x for x in v])
                out += value[None, :] + pert[None, :] * v[:, None]
            else:
                Rr = np.minimum(1.0, Rr)
                if kind == "linear":
                    out += value[None, :] + pert[None, :] * (1.0 - Rr)[:, None]
                else:
                    pw = 2.0 if kind == "cosine2" else 1.0
                    c = libm.pow_(1.0 + libm.cos_(Rr * PI), pw)
                    out += value[None, :] + (pert / 2)[None, :] * c[:, None]
        elif kind == "gaussian-outside":                 # field.h:1468-1485 (examples/atmo/ctbs)
            _, value, pert, center, r1, r2 = ini
            Rr = (vmag(geo.cC - center[None, :]) - r1) / r2
            v = libm.exp_(-Rr * Rr)
            v = np.array([1.0 if r <= 0 else (0.0 if equal(float(x), 0.0) else x) for x, r in zip(v, Rr)])
            out += value[None, :] + pert[None, :] * v[:, None]
        elif kind == "hydrostatic":
            _, p0, scale, expon = ini
            if getattr(geo, "spherical", False):
                gh = -(vmag(geo.cC) - geo.sphere_radius) * float(vmag(np.asarray(gravity, dtype=float)[None, :])[0])
            else:
                gh = geo.cC @ np.asarray(gravity, dtype=float)
            out += p0[None, :] * libm.pow_(1.0 + scale * gh, expon)[:, None]
        else:
            raise NotImplementedError(kind)
    return out[:, 0] if comps == 1 else out


def bind_bcs(ff: refio.FieldFile, topo: MeshTopo) -> list:
    out = []
    for bc in ff.bcs:
        faces = np.array(topo.boundaries.get(bc.patch, []), dtype=np.int64)
        nb = None
        if bc.neighbor:
            nb = np.array(topo.boundaries.get(bc.neighbor, []), dtype=np.int64)
        out.append(BCSpec(bc.kind, faces, np.asarray(bc.value, dtype=float), nb, bc.shape,
                          np.asarray(bc.tvalue, dtype=float), bc.tshape, bc.zMin))
    return out


def load_case(case_dir: str, exact_order: bool = True, step: int = 0) -> EulerOracle:
    blocks = refio.read_controls(os.path.join(case_dir, "controls"))
    gen = blocks["general"]
    mesh_name = gen.get("mesh", ["grid"])[0]
    nop = [int(gen.get(k, ["0"])[0]) for k in ("npx", "npy", "npz")]
    grid = refio.read_grid(os.path.join(case_dir, f"{mesh_name}_{step}"))
    params = Params.from_controls(blocks)
    topo = MeshTopo(grid)
    topo.spherical, topo.sphere_radius, topo.sphere_height = params.is_spherical, params.sphere_radius, params.sphere_height
    topo.load()
    geo = Geometry(topo, Basis(nop))
    fields, bcs = {}, {}
    for name in ("p", "U", "T", "rho"):
        base = os.path.join(case_dir, f"{name}{step}")
        if name == "rho" and not (os.path.exists(base + ".txt") or os.path.exists(base + ".bin")):
            # MeshField::read skips a file that is not there (field.h:1579-1585; examples/atmo/hydro-sphere ships no rho0): the start
            # branch forms rho from p and T on every entry and, with no conditions to apply, boundary cells keep what Solve leaves there
            fields[name], bcs[name] = np.zeros(geo.gALL), []
            continue
        ff = refio.read_field(base)
        fields[name] = init_field(ff, geo, params.gravity)
        bcs[name] = bind_bcs(ff, topo)
    orc = EulerOracle(geo, params, exact_order=exact_order)
    orc.setup(fields["rho"], fields["U"], fields["T"], fields["p"], bcs)
    return orc


def load_convection_case(case_dir: str, exact_order: bool = True, step: int = 0):
    """The same for `solver convection` (apps/convection): fields U and T, controls block convection{problem_init}."""
    from .convection import ConvectionOracle
    blocks = refio.read_controls(os.path.join(case_dir, "controls"))
    gen = blocks["general"]
    mesh_name = gen.get("mesh", ["grid"])[0]
    nop = [int(gen.get(k, ["0"])[0]) for k in ("npx", "npy", "npz")]
    grid = refio.read_grid(os.path.join(case_dir, f"{mesh_name}_{step}"))
    params = Params.from_controls(blocks)
    topo = MeshTopo(grid)
    topo.spherical, topo.sphere_radius, topo.sphere_height = params.is_spherical, params.sphere_radius, params.sphere_height
    topo.load()
    geo = Geometry(topo, Basis(nop))
    fields, bcs = {}, {}
    for name in ("U", "T"):
        ff = refio.read_field(os.path.join(case_dir, f"{name}{step}"))
        fields[name] = init_field(ff, geo, params.gravity)
        bcs[name] = bind_bcs(ff, topo)
    orc = ConvectionOracle(geo, params, exact_order=exact_order)
    orc.setup_convection(fields["T"], fields["U"], bcs, blocks.get("convection", {}).get("problem_init", ["NONE"])[0],
                         int(gen.get("end_step", ["1"])[0]))
    return orc
