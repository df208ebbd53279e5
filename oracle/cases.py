"""Synthetic euler cases (grid + controls + field files) for the parity tests -- TEST INFRASTRUCTURE ONLY.

The case DEFINITIONS below restate the reference's example set-ups (parameters only):
  rising thermal bubble 2-D   examples/atmo/srtb/{controls,bubble,T0.txt,U0.txt,p0.txt,rho0.txt}
  rising thermal bubble 3-D   examples/atmo/srtb-3d/...
  isentropic vortex           examples/isentropic/...
  hill (terrain-following)    examples/hills/hill block layout with the srtb-amr-hill euler controls
The grids are generated here (own structured generator; cell order x-major with z fastest and per-cell face
order z-,z+,y-,y+,x-,x+ like the reference's block mesher, src/mesh/hexMesh.cpp:227-342) and written in the
reference's grid grammar, so the SAME files feed the reference binary, the numpy oracle and the CUDA path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from .refio import Grid, write_grid_text


def box_grid(n, lo, hi, patches, vertex_map=None) -> Grid:
    """Structured hex grid of n=(nx,ny,nz) cells on [lo,hi]. patches: dict side -> patch name, sides in
    {'x-','x+','y-','y+','z-','z+'}; sides mapped to the name 'delete' are dropped by LoadMesh (2-D cases)."""
    nx, ny, nz = n
    vx, vy, vz = nx + 1, ny + 1, nz + 1
    xs = np.linspace(lo[0], hi[0], vx)
    ys = np.linspace(lo[1], hi[1], vy)
    zs = np.linspace(lo[2], hi[2], vz)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    V = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    if vertex_map is not None:
        V = vertex_map(V)

    def vid(i, j, k):
        return (i * vy + j) * vz + k

    facets = []
    fz = np.zeros((nx, ny, vz), dtype=np.int64)
    fy = np.zeros((nx, vy, nz), dtype=np.int64)
    fx = np.zeros((vx, ny, nz), dtype=np.int64)
    for i in range(nx):
        for j in range(ny):
            for k in range(vz):
                fz[i, j, k] = len(facets)
                facets.append([vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k)])
    for i in range(nx):
        for j in range(vy):
            for k in range(nz):
                fy[i, j, k] = len(facets)
                facets.append([vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j, k + 1), vid(i, j, k + 1)])
    for i in range(vx):
        for j in range(ny):
            for k in range(nz):
                fx[i, j, k] = len(facets)
                facets.append([vid(i, j, k), vid(i, j + 1, k), vid(i, j + 1, k + 1), vid(i, j, k + 1)])
    cells = []
    for i in range(nx):
        for j in range(ny):
            for k in range(nz):
                cells.append([int(fz[i, j, k]), int(fz[i, j, k + 1]), int(fy[i, j, k]), int(fy[i, j + 1, k]),
                              int(fx[i, j, k]), int(fx[i + 1, j, k])])
    sides = {"z-": fz[:, :, 0], "z+": fz[:, :, nz], "y-": fy[:, 0, :], "y+": fy[:, ny, :], "x-": fx[0], "x+": fx[nx]}
    bnd: dict = {}
    for side, name in patches.items():
        bnd.setdefault(name, [])
        bnd[name] += [int(f) for f in sides[side].reshape(-1)]
    return Grid(V, facets, cells, bnd)


@dataclass
class Case:
    name: str
    grid: Grid
    nop: tuple                      # polynomial degree per direction (npx,npy,npz)
    general: dict                   # extra general{} entries
    euler: dict                     # euler{} entries
    fields: dict = field(default_factory=dict)   # name -> text of <name>0.txt

    def write(self, d: str, end_step: int, write_interval: int | None = None) -> str:
        os.makedirs(d, exist_ok=True)
        write_grid_text(os.path.join(d, "grid_0.txt"), self.grid)
        g = dict(solver="euler", mesh="grid", state="TRANSIENT", start_step=0, end_step=end_step,
                 write_interval=write_interval or end_step, n_deferred=0, convection_scheme="RUSANOV",
                 nonortho_scheme="OVER_RELAXED", blend_factor=0.3, parallel_method="BLOCKED", method="PCG",
                 preconditioner="DIAG", tolerance=1e-5, max_iterations=6400, SOR_omega=1.7,
                 npx=self.nop[0], npy=self.nop[1], npz=self.nop[2])
        g.update(self.general)
        lines = ["general", "{"]
        for k, v in g.items():
            lines.append(f"    {k} {v}")
        lines.append("    probe 0 {}")
        lines += ["}", "prepare", "{", "    fields 4 { U T p rho }", "}", "euler", "{"]
        for k, v in self.euler.items():
            lines.append(f"    {k} {v}")
        lines.append("}")
        with open(os.path.join(d, "controls"), "w") as f:
            f.write("\n".join(lines) + "\n")
        for name, txt in self.fields.items():
            with open(os.path.join(d, f"{name}0.txt"), "w") as f:
                f.write(txt)
        return d


def _field(comps, internal, bcs):
    out = [f"size {comps}", "internal 1", "{", f"    {internal}", "}", f"boundary {len(bcs)}", "{"]
    for patch, body in bcs.items():
        out.append(f"    {patch} {{")
        for ln in body:
            out.append(f"        {ln}")
        out.append("    }")
    out.append("}")
    return "\n".join(out) + "\n"


def _all(patches, body):
    return {p: list(body) for p in patches}


def bubble2d(n=10, order=4, scheme="BDF1") -> Case:
    """examples/atmo/srtb: 1 km x 1 km x-z slab, one element thick in y, cosine theta bubble, nu = 1.5."""
    grid = box_grid((n, 1, n), (0, 0, 0), (1000, 100, 1000),
                    {"z+": "top", "z-": "bottom", "x-": "sides", "x+": "sides", "y-": "delete", "y+": "delete"})
    pt = ["top", "bottom", "sides"]
    return Case(
        "bubble2d", grid, (order, 0, order),
        dict(rho=1.177, viscosity=1.5, dt=0.005, time_scheme=scheme, gravity="0 0 -9.80606"),
        dict(velocity_UR=0.5, pressure_UR=0.8, t_UR=0.8, diffusion="YES", buoyancy="YES"),
        dict(T=_field(1, "cosine 0 0.5    500 50 350   250 1000 250", _all(pt, ["type NEUMANN"])),
             U=_field(3, "uniform 0 0 0", _all(pt, ["type SYMMETRY"])),
             p=_field(1, "uniform 0", _all(pt, ["type NEUMANN"])),
             rho=_field(1, "uniform 0", _all(pt, ["type NEUMANN"]))))


def bubble3d(n=4, order=4, scheme="AB1") -> Case:
    """examples/atmo/srtb-3d: 1 km cube, gravity along -y, dt 0.00125."""
    grid = box_grid((n, n, n), (0, 0, 0), (1000, 1000, 1000),
                    {"y+": "top", "y-": "bottom", "x-": "sides", "x+": "sides", "z-": "sides", "z+": "sides"})
    pt = ["top", "bottom", "sides"]
    return Case(
        "bubble3d", grid, (order, order, order),
        dict(rho=1.177, viscosity=1.5, dt=0.00125, time_scheme=scheme, gravity="0 -9.80606 0"),
        dict(velocity_UR=0.5, pressure_UR=0.8, t_UR=0.8, diffusion="YES", buoyancy="YES"),
        dict(T=_field(1, "cosine 0 0.5    500 350 500   250 250 250", _all(pt, ["type NEUMANN"])),
             U=_field(3, "uniform 0 0 0", _all(pt, ["type SYMMETRY"])),
             p=_field(1, "uniform 0", _all(pt, ["type NEUMANN"])),
             rho=_field(1, "uniform 0", _all(pt, ["type NEUMANN"]))))


def vortex(n=15, order=3) -> Case:
    """examples/isentropic: periodic x-y square [-5,5]^2, unit gas, beta = 5 vortex advected by (1,1,0)."""
    grid = box_grid((n, n, 1), (-5, -5, -0.5), (5, 5, 0.5),
                    {"x-": "inx", "x+": "outx", "y-": "iny", "y+": "outy", "z-": "delete", "z+": "delete"})
    cyc = {"inx": ["type CYCLIC", "neighbor outx"], "outx": ["type CYCLIC", "neighbor inx"],
           "iny": ["type CYCLIC", "neighbor outy"], "outy": ["type CYCLIC", "neighbor iny"]}
    return Case(
        "vortex", grid, (order, order, 0),
        dict(rho=1.0, T0=1.0, beta=1.0, P0=1.0, cp=3.5, cv=2.5, viscosity=0.0, dt=0.0005, time_scheme="BDF1",
             gravity="0 -9.80606 0"),
        dict(velocity_UR=0.5, pressure_UR=0.8, t_UR=0.8, diffusion="NO", buoyancy="NO", problem_init="ISENTROPIC_VORTEX"),
        dict(T=_field(1, "uniform 0", cyc), U=_field(3, "uniform 1 1 0", cyc), p=_field(1, "uniform 0", cyc),
             rho=_field(1, "uniform 0", cyc)))


def hill3d(nx=12, ny=2, nz=8, order=3) -> Case:
    """Terrain-following mountain-wave set-up: cosine hill of height 200 m centred at x = 1000 m on a
    3400 m x 1400 m x-z section extruded in y, uniform inflow (10,0,0), DIRICHLET inlet / NEUMANN outlet /
    SYMMETRY ground, top and sides, buoyancy and diffusion on.  Non-affine elements (per-node Jinv varies)."""
    H, xc, hw, Lz = 200.0, 1000.0, 400.0, 1400.0

    def terrain(V):
        V = V.copy()
        x = V[:, 0]
        h = np.where(np.abs(x - xc) < hw, H * np.cos(0.5 * np.pi * (x - xc) / hw) ** 2, 0.0)
        V[:, 2] = h + V[:, 2] * (Lz - h) / Lz
        return V

    grid = box_grid((nx, ny, nz), (0, 0, 0), (3400, 100.0 * ny, Lz),
                    {"x-": "inlet", "x+": "outlet", "z-": "WALLS", "z+": "top", "y-": "sides", "y+": "sides"}, terrain)
    sc = {"inlet": ["type NEUMANN"], "outlet": ["type NEUMANN"], "WALLS": ["type NEUMANN"], "top": ["type NEUMANN"],
          "sides": ["type NEUMANN"]}
    ub = {"inlet": ["type DIRICHLET", "value 10 0 0"], "outlet": ["type NEUMANN"], "WALLS": ["type SYMMETRY"],
          "top": ["type SYMMETRY"], "sides": ["type SYMMETRY"]}
    return Case(
        "hill3d", grid, (order, order, order),
        dict(rho=1.177, viscosity=1.5, dt=0.001, time_scheme="BDF1", gravity="0 0 -9.80606"),
        dict(velocity_UR=0.5, pressure_UR=0.8, t_UR=0.8, diffusion="YES", buoyancy="YES"),
        dict(T=_field(1, "cosine 0 0.5    1700 100 700   400 1000 300", sc), U=_field(3, "uniform 10 0 0", ub),
             p=_field(1, "uniform 0", sc), rho=_field(1, "uniform 0", sc)))


CASES = {"bubble2d": bubble2d, "bubble3d": bubble3d, "vortex": vortex, "hill3d": hill3d}
