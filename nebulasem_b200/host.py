"""ctypes binding of libnsem_host.so: the C++ host side of the drop-in (mesh + DG geometry + euler set-up + I/O).

`Solver.open_case(dir)` performs what the reference's `euler ./controls` does before its time loop;
`Solver.synthetic(kind, ...)` builds the same cases in memory (benchmarks).  `attach()` uploads everything to the
GPU through the C ABI of include/nsem_c.h; `step()` runs the CUDA time loop.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

HOST_LIB_PATH = os.path.join(os.environ.get("NSEM_LIBDIR") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib"), "libnsem_host.so")
HOST_EXPORTS = ["nsemh_error", "nsemh_close", "nsemh_open_case", "nsemh_open_case_part", "nsemh_synthetic", "nsemh_synthetic_part", "nsemh_patch_faces",
                "nsemh_peers", "nsemh_diagnostics", "nsemh_attach", "nsemh_step",
                "nsemh_upload", "nsemh_download", "nsemh_upload_async", "nsemh_download_async", "nsemh_adopt_refined_state", "nsemh_restart_state", "nsemh_regrid", "nsemh_enable_amr", "nsemh_cell_levels", "nsemh_write_amr_grid", "nsemh_write", "nsemh_write_vtk", "nsemh_run", "nsemh_sync", "nsemh_time",
                "nsemh_launch_count", "nsemh_kernel_info", "nsemh_halo_info", "nsemh_halo_wait_ms", "nsemh_set_schedule", "nsemh_dims", "nsemh_params", "nsemh_f64", "nsemh_u32",
                "nsemh_state_ptr", "nsemh_totals", "nsemh_partition_grid"]
_lib = None


def load_host_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(HOST_LIB_PATH):
        raise RuntimeError(f"{HOST_LIB_PATH} is missing: run `python -m nebulasem_b200.build`")
    capi.load_library()          # libnsem_cuda.so first (rpath $ORIGIN also finds it)
    lib = C.CDLL(HOST_LIB_PATH)
    vp = C.c_void_p
    lib.nsemh_error.argtypes = [vp]
    lib.nsemh_error.restype = C.c_char_p
    lib.nsemh_close.argtypes = [vp]
    lib.nsemh_close.restype = None
    lib.nsemh_open_case.argtypes = [C.c_char_p, C.c_int]
    lib.nsemh_open_case.restype = vp
    lib.nsemh_open_case_part.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.nsemh_open_case_part.restype = vp
    lib.nsemh_synthetic.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.nsemh_synthetic.restype = vp
    lib.nsemh_synthetic_part.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                         C.c_int, C.c_int]
    lib.nsemh_synthetic_part.restype = vp
    lib.nsemh_patch_faces.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint32), C.c_uint64]
    lib.nsemh_patch_faces.restype = C.c_uint64
    lib.nsemh_peers.argtypes = [vp, C.POINTER(C.c_int), C.c_int]
    lib.nsemh_attach.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    for n in ("nsemh_upload", "nsemh_download", "nsemh_upload_async", "nsemh_download_async", "nsemh_restart_state", "nsemh_run", "nsemh_sync"):
        getattr(lib, n).argtypes = [vp]
    lib.nsemh_step.argtypes = [vp, C.c_int]
    u32p = C.POINTER(C.c_uint32)
    u8p = C.POINTER(C.c_uint8)
    lib.nsemh_regrid.argtypes = [vp, u8p, u8p, C.c_uint32]
    lib.nsemh_enable_amr.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_char_p, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.nsemh_cell_levels.argtypes = [vp, C.POINTER(C.c_int32), C.c_uint32]
    lib.nsemh_adopt_refined_state.argtypes = [vp, vp, u32p, C.c_uint32, u32p, C.c_uint32, u32p, C.c_uint32, C.c_int]
    lib.nsemh_write.argtypes = [vp, C.c_int]
    lib.nsemh_write_amr_grid.argtypes = [vp, C.c_int]
    lib.nsemh_write_vtk.argtypes = [vp, C.c_int]
    lib.nsemh_time.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.nsemh_diagnostics.argtypes = [vp, C.POINTER(C.c_double)]
    lib.nsemh_launch_count.argtypes = [vp]
    lib.nsemh_launch_count.restype = C.c_uint64
    lib.nsemh_kernel_info.argtypes = [vp]
    lib.nsemh_kernel_info.restype = C.c_char_p
    lib.nsemh_halo_info.argtypes = [vp]
    lib.nsemh_halo_info.restype = C.c_char_p
    lib.nsemh_halo_wait_ms.argtypes = [vp, C.POINTER(C.c_double)]
    lib.nsemh_set_schedule.argtypes = [vp, C.POINTER(C.c_uint32), C.c_uint32]
    lib.nsemh_dims.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.nsemh_params.argtypes = [vp, C.POINTER(C.c_double)]
    lib.nsemh_f64.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint64)]
    lib.nsemh_f64.restype = C.POINTER(C.c_double)
    lib.nsemh_u32.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint64)]
    lib.nsemh_u32.restype = C.POINTER(C.c_uint32)
    lib.nsemh_state_ptr.argtypes = [vp, C.c_char_p]
    lib.nsemh_state_ptr.restype = C.POINTER(C.c_double)
    lib.nsemh_totals.argtypes = [vp, C.POINTER(C.c_double)]
    lib.nsemh_totals.restype = None
    lib.nsemh_partition_grid.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint32)]
    lib.nsemh_partition_grid.restype = C.c_int64
    _lib = lib
    return lib


def partition_grid(grid_noext: str, n_cells: int, n_faces: int, nparts: int, method: str = "METIS", pxyz=(1, 1, 1)):
    """Prepare::decomposeMesh's cell -> rank map for a grid file (METIS | XYZ | CELLID) and the grid's gFMC flags.
    Raises when the decomposition would cut a non-conforming face (field.cpp:1215-1220)."""
    lib = load_host_library()
    part = np.zeros(n_cells, dtype=np.uint32)
    fmc = np.zeros(n_faces, dtype=np.uint32)
    n = lib.nsemh_partition_grid(os.fspath(grid_noext).encode(), nparts, method.encode(), int(pxyz[0]), int(pxyz[1]), int(pxyz[2]),
                                 part.ctypes.data_as(C.POINTER(C.c_uint32)), fmc.ctypes.data_as(C.POINTER(C.c_uint32)))
    if n < 0:
        raise capi.NsemError(lib.nsemh_error(None).decode())
    assert n == n_cells
    return part, fmc


class Solver:
    """Host-side euler solver (nsemh::EulerSolver)."""

    def __init__(self, handle):
        self.lib = load_host_library()
        if not handle:
            raise capi.NsemError(self.lib.nsemh_error(None).decode())
        self.h = C.c_void_p(handle)
        self._refresh_dims()
        p = (C.c_double * 12)()
        self.lib.nsemh_params(self.h, p)
        self.params = dict(P0=p[0], T0=p[1], cp=p[2], cv=p[3], viscosity=p[4], Pr=p[5], gravity=(p[6], p[7], p[8]),
                           dt=p[9], buoyancy=bool(p[10]), diffusion=bool(p[11]))

    def _refresh_dims(self):
        d = (C.c_uint64 * 10)()
        self.lib.nsemh_dims(self.h, d)
        (self.NPX, self.NPY, self.NPZ, self.NP, self.NPF, self.nBCS, self.nCells, self.nFacets, self.gBCSfield,
         self.gALL) = [int(x) for x in d]

    @classmethod
    def open_case(cls, case_dir: str, step: int = 0, rank: int = 0, nranks: int = 1) -> "Solver":
        """The euler (or convection) app's set-up on a case directory; nranks > 1 keeps partition `rank` of the decomposition the controls name."""
        lib = load_host_library()
        if nranks == 1:
            return cls(lib.nsemh_open_case(os.fspath(case_dir).encode(), step))
        s = cls(lib.nsemh_open_case_part(os.fspath(case_dir).encode(), step, rank, nranks))
        s.rank, s.nranks = rank, nranks
        return s

    @classmethod
    def synthetic(cls, kind: str, nx: int, ny: int, nz: int, order: int) -> "Solver":
        lib = load_host_library()
        return cls(lib.nsemh_synthetic(kind.encode(), nx, ny, nz, order))

    @classmethod
    def synthetic_part(cls, kind: str, nx: int, ny: int, nz: int, order: int, rank: int, nranks: int, decomp: str = "METIS",
                       pxyz=(1, 1, 1)) -> "Solver":
        """Partition `rank` of `nranks` of the synthetic case (decomp: METIS | XYZ with pxyz | CELLID)."""
        lib = load_host_library()
        s = cls(lib.nsemh_synthetic_part(kind.encode(), nx, ny, nz, order, rank, nranks, decomp.encode(), *[int(x) for x in pxyz]))
        s.rank, s.nranks = rank, nranks
        return s

    def patch_faces(self, name: str) -> np.ndarray:
        n = self.lib.nsemh_patch_faces(self.h, name.encode(), None, 0)
        out = np.zeros(int(n), dtype=np.uint32)
        if n:
            self.lib.nsemh_patch_faces(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_uint32)), n)
        return out

    def peers(self):
        buf = (C.c_int * 64)()
        n = self.lib.nsemh_peers(self.h, buf, 64)
        return [int(buf[i]) for i in range(n)]

    def close(self):
        if getattr(self, "h", None):
            self.lib.nsemh_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise capi.NsemError(self.lib.nsemh_error(self.h).decode())

    def f64(self, name: str) -> np.ndarray:
        n = C.c_uint64()
        p = self.lib.nsemh_f64(self.h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def u32(self, name: str) -> np.ndarray:
        n = C.c_uint64()
        p = self.lib.nsemh_u32(self.h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def state(self):
        """(rho, U, T, p) host copies over all nodes (reference layout)."""
        return self.f64("rho"), self.f64("U").reshape(-1, 3), self.f64("T"), self.f64("p")

    def set_state(self, rho=None, U=None, T=None, p=None):
        for name, a, comps in (("rho", rho, 1), ("U", U, 3), ("T", T, 1), ("p", p, 1)):
            if a is None:
                continue
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
            assert a.size == self.gALL * comps
            C.memmove(self.lib.nsemh_state_ptr(self.h, name.encode()), a.ctypes.data, a.nbytes)

    def totals(self):
        t = (C.c_double * 3)()
        self.lib.nsemh_totals(self.h, t)
        return tuple(t)

    # ---- device --------------------------------------------------------------------------------------
    def attach(self, device: int = 0, rank: int = 0, nranks: int = 1, unique_id: bytes | None = None):
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._ck(self.lib.nsemh_attach(self.h, device, rank, nranks, uid))

    def upload(self):
        self._ck(self.lib.nsemh_upload(self.h))

    def step(self, n: int = 1):
        self._ck(self.lib.nsemh_step(self.h, int(n)))

    def download(self):
        self._ck(self.lib.nsemh_download(self.h))

    def adopt_refined_state(self, old: "Solver", refine_map, coarse_map, cell_map, restart: bool = True):
        """AMR regrid with the state resident on the device: this solver (regridded mesh, attached) takes the state of `old` (mesh before
        the regrid, attached to the same device) -- MeshField::refineField for rho, U, T, p (nsem_refine_state), then with `restart` the
        set-up's restart branch: p from rho, ghost cells (nsem_restart_state).  The maps are those of MeshObject::refineMesh."""
        maps = [np.ascontiguousarray(m, dtype=np.uint32) for m in (refine_map, coarse_map, cell_map)]
        args = []
        for m in maps:
            args += [m.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(m.size)]
        self._ck(self.lib.nsemh_adopt_refined_state(self.h, old.h, *args, 1 if restart else 0))

    # ---- AMR in memory (amr.cpp).  The solver must have been created with NSEM_AMR=1 in the environment or amr_step in its controls.
    def enable_amr(self, direction=(0.0, 0.0, 0.0), field="T", field_min=0.2, field_max=0.6, max_level=1, buffer_zone=2):
        """refinement{} parameters (Controls::enrollRefine); direction = the axis that is never split (2-D refinement), zero = 3-D."""
        self._ck(self.lib.nsemh_enable_amr(self.h, float(direction[0]), float(direction[1]), float(direction[2]), field.encode(),
                                           float(field_min), float(field_max), int(max_level), int(buffer_zone)))

    def regrid(self, refine=None, coarsen=None):
        """Regrid: this handle continues as the solver on the new mesh; when attached, the state is transferred on the device
        (nsem_refine_state + nsem_restart_state) and the old context is destroyed.  refine / coarsen: one flag per cell; both None =
        tag by the refinement indicator.  The maps of the regrid are u32("refineMap" | "coarseMap" | "cellMap")."""
        if refine is None and coarsen is None:
            self._ck(self.lib.nsemh_regrid(self.h, None, None, 0))
        else:
            r = np.ascontiguousarray(refine, dtype=np.uint8)
            c = np.ascontiguousarray(coarsen, dtype=np.uint8)
            assert r.size == self.nBCS and c.size == self.nBCS
            u8p = C.POINTER(C.c_uint8)
            self._ck(self.lib.nsemh_regrid(self.h, r.ctypes.data_as(u8p), c.ctypes.data_as(u8p), r.size))
        self._refresh_dims()

    def write_amr_grid(self, dump: int):
        """<mesh>_<dump>.txt (the current grid) + <mesh>_<dump>.forest in the case directory: with the field dump of the same index an AMR
        run restarts from there."""
        self._ck(self.lib.nsemh_write_amr_grid(self.h, int(dump)))

    def cell_levels(self) -> np.ndarray:
        out = np.zeros(self.nBCS, dtype=np.int32)
        self._ck(self.lib.nsemh_cell_levels(self.h, out.ctypes.data_as(C.POINTER(C.c_int32)), out.size))
        return out

    def restart_state(self):
        """The set-up's restart branch on the device (euler.cpp:150-162): p from rho, ghost cells of rho, p, U, T (nsem_restart_state)."""
        self._ck(self.lib.nsemh_restart_state(self.h))

    def upload_async(self):
        """Enqueue the upload of the host state (pipelined batches); see nsem_upload_state_async."""
        self._ck(self.lib.nsemh_upload_async(self.h))

    def download_async(self):
        """Enqueue the download of the device state into the out_* host arrays; complete after sync()."""
        self._ck(self.lib.nsemh_download_async(self.h))

    def state_out(self):
        """(rho, U, T, p) as the last download_async() + sync() left them."""
        return self.f64("out_rho"), self.f64("out_U").reshape(-1, 3), self.f64("out_T"), self.f64("out_p")

    def sync(self):
        self._ck(self.lib.nsemh_sync(self.h))

    def write(self, index: int):
        self._ck(self.lib.nsemh_write(self.h, int(index)))

    def write_vtk(self, index: int):
        """<mesh><index>.vtk from the state on the host, as `prepare ./controls -vtk -start index` writes it (Vtk::write_vtk, vtk.cpp:125-286)."""
        self._ck(self.lib.nsemh_write_vtk(self.h, int(index)))

    def run(self):
        self._ck(self.lib.nsemh_run(self.h))

    def time_steps(self, nsteps: int, per_kernel: bool = False):
        ms = C.c_double()
        pk = (C.c_double * 4)()
        self._ck(self.lib.nsemh_time(self.h, int(nsteps), C.byref(ms), pk if per_kernel else None))
        return ms.value, list(pk)

    def diagnostics(self) -> dict:
        """Courant max/min/avg, mass, energy, volume of the device state and the relative losses euler.cpp:278-281 prints."""
        o = (C.c_double * 9)()
        self._ck(self.lib.nsemh_diagnostics(self.h, o))
        return dict(courant_max=o[0], courant_min=o[1], courant_avg=o[2], mass=o[3], energy=o[4], volume=o[5],
                    mass_loss=(o[6] - o[3]) / o[6], energy_loss=(o[7] - o[4]) / o[7], volume_loss=(o[8] - o[5]) / o[8])

    def set_schedule(self, order):
        if order is None:
            self._ck(self.lib.nsemh_set_schedule(self.h, None, 0))
        else:
            o = np.ascontiguousarray(order, dtype=np.uint32)
            self._ck(self.lib.nsemh_set_schedule(self.h, o.ctypes.data_as(C.POINTER(C.c_uint32)), len(o)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.nsemh_launch_count(self.h))

    @property
    def kernel_info(self) -> str:
        return self.lib.nsemh_kernel_info(self.h).decode()

    def halo_wait_ms(self):
        """ms this rank's stream waited for its neighbours' halo flags since the last call: (after sweep A, after sweep B, state exchanges)."""
        out = (C.c_double * 3)()
        self._ck(self.lib.nsemh_halo_wait_ms(self.h, out))
        return [float(x) for x in out]

    @property
    def halo_info(self) -> str:
        """Transport of the face-trace halo: "peer memory ...", "nccl send/recv" or "none"."""
        return self.lib.nsemh_halo_info(self.h).decode()
