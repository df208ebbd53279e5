"""ctypes binding of include/nsem_c.h (libnsem_cuda.so).

Thin and literal: one Python method per C entry point, numpy arrays in the reference's layouts.  There is no
CPU fallback: if the shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NSEM_LIBDIR selects another build of the same libraries (kernel-variant experiments, profiles/variants_r1.md)
LIB_PATH = os.path.join(os.environ.get("NSEM_LIBDIR") or os.path.join(_HERE, "lib"), "libnsem_cuda.so")

BC_KINDS = {"NEUMANN": 1, "DIRICHLET": 2, "SYMMETRY": 3, "CYCLIC": 4, "GHOST": 5, "FIXED": 6, "ROBIN": 7,
            "CALC_DIRICHLET": 6, "UNLISTED": 8}
FIELDS = {"rho": 0, "p": 1, "U": 2, "T": 3}

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)


class NsemMesh(C.Structure):
    _fields_ = [("n_cells_real", C.c_uint32), ("n_cells_all", C.c_uint32), ("n_faces", C.c_uint32),
                ("cV", _dp), ("Jinv", _dp), ("fN", _dp), ("fI", _dp), ("face_normal", _dp),
                ("FO", _up), ("FN", _up), ("face_begin", _up), ("face_end", _up), ("all_faces", _up),
                ("face_id", _up), ("face_owner", _up), ("face_neigh", _up), ("face_mortar", _up),
                # read only on non-conforming meshes (face_mortar != 0 somewhere)
                ("cC", _dp), ("face_center", _dp), ("psi_ref", _dp * 6), ("psi_cor", _dp * 6)]


class NsemBC(C.Structure):
    _fields_ = [("field", C.c_int32), ("kind", C.c_int32), ("n_faces", C.c_uint32), ("faces", _up),
                ("peer_faces", _up), ("value", C.c_double * 3), ("shape", C.c_double), ("tvalue", C.c_double * 3),
                ("tshape", C.c_double), ("zMin", C.c_double), ("fixed", _dp)]


class NsemParams(C.Structure):
    _fields_ = [("P0", C.c_double), ("T0", C.c_double), ("cp", C.c_double), ("cv", C.c_double),
                ("viscosity", C.c_double), ("Pr", C.c_double), ("gravity", C.c_double * 3), ("dt", C.c_double),
                ("buoyancy", C.c_int32), ("diffusion", C.c_int32)]


class NsemHaloPeer(C.Structure):
    _fields_ = [("peer_rank", C.c_int32), ("n_faces", C.c_uint32), ("faces", _up)]


class NsemRegrid(C.Structure):
    """nsem_regrid of include/nsem_c.h: one regrid as MeshObject::refineMesh reports it."""
    _fields_ = [("n_cells_new", C.c_uint32),
                ("refine_map", C.POINTER(C.c_uint32)), ("n_refine_map", C.c_uint32),
                ("coarse_map", C.POINTER(C.c_uint32)), ("n_coarse_map", C.c_uint32),
                ("cell_map", C.POINTER(C.c_uint32)), ("n_cell_map", C.c_uint32),
                ("old_cV", C.POINTER(C.c_double)), ("old_cC", C.POINTER(C.c_double)),
                ("new_cV", C.POINTER(C.c_double)), ("new_cC", C.POINTER(C.c_double)),
                ("old_node_cC", C.POINTER(C.c_double)),
                ("psi_ref", C.POINTER(C.c_double) * 6), ("psi_cor", C.POINTER(C.c_double) * 6)]


EXPORTS = ["nsem_create", "nsem_destroy", "nsem_last_error", "nsem_get_unique_id", "nsem_set_order", "nsem_set_basis",
           "nsem_upload_mesh", "nsem_set_bcs", "nsem_set_halo", "nsem_set_params", "nsem_set_schedule",
           "nsem_pin_host", "nsem_upload_state", "nsem_download_state", "nsem_upload_state_async", "nsem_download_state_async", "nsem_upload_ref", "nsem_upload_geopotential", "nsem_euler_step", "nsem_exchange_state_halos", "nsem_diagnostics",
           "nsem_sync", "nsem_time_steps", "nsem_launch_count", "nsem_kernel_info", "nsem_refine_state", "nsem_restart_state", "nsem_download_gradients",
           "nsem_op_cds", "nsem_op_rusanov", "nsem_op_gradf_strong", "nsem_op_divf_weak", "nsem_op_apply_bcs", "nsem_op_halo", "nsem_halo_info", "nsem_halo_wait_ms", "nsem_upload_coords", "nsem_set_convection_scheme", "nsem_set_ab_order", "nsem_allreduce_host", "nsem_device", "nsem_set_sphere", "nsem_set_convection", "nsem_convection_step"]

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m nebulasem_b200.build` (needs nvcc); "
                           "nebulasem_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.nsem_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    lib.nsem_destroy.argtypes = [vp]
    lib.nsem_destroy.restype = None
    lib.nsem_last_error.argtypes = [vp]
    lib.nsem_last_error.restype = C.c_char_p
    lib.nsem_get_unique_id.argtypes = [vp]
    lib.nsem_set_order.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.nsem_set_basis.argtypes = [vp, C.POINTER(_dp), C.POINTER(_dp)]
    lib.nsem_upload_mesh.argtypes = [vp, C.POINTER(NsemMesh)]
    lib.nsem_set_bcs.argtypes = [vp, C.POINTER(NsemBC), C.c_uint32]
    lib.nsem_set_halo.argtypes = [vp, C.POINTER(NsemHaloPeer), C.c_uint32]
    lib.nsem_set_params.argtypes = [vp, C.POINTER(NsemParams)]
    lib.nsem_set_schedule.argtypes = [vp, _up, C.c_uint32]
    lib.nsem_pin_host.argtypes = [vp, vp, C.c_uint64]
    lib.nsem_upload_state.argtypes = [vp, _dp, _dp, _dp, _dp]
    lib.nsem_download_state.argtypes = [vp, _dp, _dp, _dp, _dp]
    lib.nsem_upload_state_async.argtypes = [vp, _dp, _dp, _dp, _dp]
    lib.nsem_download_state_async.argtypes = [vp, _dp, _dp, _dp, _dp]
    lib.nsem_download_gradients.argtypes = [vp, _dp, _dp]
    lib.nsem_op_cds.argtypes = [vp, _dp, _dp]
    lib.nsem_op_rusanov.argtypes = [vp, _dp]
    lib.nsem_op_gradf_strong.argtypes = [vp, _dp, _dp]
    lib.nsem_op_divf_weak.argtypes = [vp, _dp, _dp, _dp]
    lib.nsem_op_apply_bcs.argtypes = [vp, C.c_int, _dp]
    lib.nsem_op_halo.argtypes = [vp]
    lib.nsem_halo_info.argtypes = [vp]
    lib.nsem_upload_coords.argtypes = [vp, _dp]
    lib.nsem_set_convection.argtypes = [vp, C.c_int, C.c_double, C.c_long]
    lib.nsem_set_sphere.argtypes = [vp, C.c_double]
    lib.nsem_set_ab_order.argtypes = [vp, C.c_int]
    lib.nsem_set_convection_scheme.argtypes = [vp, C.c_int, C.c_double]
    lib.nsem_allreduce_host.argtypes = [vp, C.c_void_p, C.c_uint64, C.c_int]
    lib.nsem_device.argtypes = [vp]
    lib.nsem_convection_step.argtypes = [vp, C.c_int]
    lib.nsem_halo_info.restype = C.c_char_p
    lib.nsem_halo_wait_ms.argtypes = [vp, _dp]
    lib.nsem_refine_state.argtypes = [vp, C.POINTER(NsemRegrid), vp]
    lib.nsem_restart_state.argtypes = [vp]
    lib.nsem_upload_ref.argtypes = [vp, _dp, _dp, _dp]
    lib.nsem_upload_geopotential.argtypes = [vp, _dp]
    lib.nsem_euler_step.argtypes = [vp, C.c_int]
    lib.nsem_exchange_state_halos.argtypes = [vp]
    lib.nsem_diagnostics.argtypes = [vp, _dp]
    lib.nsem_sync.argtypes = [vp]
    lib.nsem_time_steps.argtypes = [vp, C.c_int, _dp, _dp]
    lib.nsem_launch_count.argtypes = [vp]
    lib.nsem_launch_count.restype = C.c_uint64
    lib.nsem_kernel_info.argtypes = [vp]
    lib.nsem_kernel_info.restype = C.c_char_p
    _lib = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _pd(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _pu(a):
    return a.ctypes.data_as(_up) if a is not None else None


class NsemError(RuntimeError):
    pass


class Context:
    """One nsem_ctx (= one mesh partition on one GPU)."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, unique_id: bytes | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        uid = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        rc = self.lib.nsem_create(device, rank, nranks, uid, C.byref(h))
        if rc != 0:
            raise NsemError(self.lib.nsem_last_error(None).decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.nsem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise NsemError(self.lib.nsem_last_error(self.h).decode())

    # ---- set-up --------------------------------------------------------------------------------
    def set_order(self, NPX, NPY, NPZ):
        self.NP = NPX * NPY * NPZ
        self._ck(self.lib.nsem_set_order(self.h, NPX, NPY, NPZ))

    def set_basis(self, dpsi, wgl):
        d = [_f64(x) for x in dpsi]
        w = [_f64(x) for x in wgl]
        dp = (_dp * 3)(*[_pd(x) for x in d])
        wp = (_dp * 3)(*[_pd(x) for x in w])
        self._ck(self.lib.nsem_set_basis(self.h, dp, wp))

    def upload_mesh(self, *, n_cells_real, n_cells_all, n_faces, cV, Jinv, fN, fI, face_normal, FO, FN, face_begin,
                    face_end, all_faces, face_id, face_owner, face_neigh, face_mortar, cC=None, face_center=None,
                    psi_ref=None, psi_cor=None):
        """cC, face_center, psi_ref[6], psi_cor[6] are only needed when face_mortar marks non-conforming faces."""
        arrs = dict(cV=_f64(cV), Jinv=_f64(Jinv), fN=_f64(fN), fI=_f64(fI), face_normal=_f64(face_normal), FO=_u32(FO),
                    FN=_u32(FN), face_begin=_u32(face_begin), face_end=_u32(face_end), all_faces=_u32(all_faces),
                    face_id=_u32(face_id), face_owner=_u32(face_owner), face_neigh=_u32(face_neigh),
                    face_mortar=_u32(face_mortar))
        m = NsemMesh()
        m.n_cells_real, m.n_cells_all, m.n_faces = int(n_cells_real), int(n_cells_all), int(n_faces)
        for k, a in arrs.items():
            setattr(m, k, _pd(a) if a.dtype == np.float64 else _pu(a))
        keep = []
        if cC is not None and face_center is not None:
            keep += [_f64(cC), _f64(face_center)]
            m.cC, m.face_center = _pd(keep[0]), _pd(keep[1])
        if psi_ref is not None and psi_cor is not None:
            pr = [_f64(x) for x in psi_ref]
            pc = [_f64(x) for x in psi_cor]
            keep += pr + pc
            m.psi_ref = (_dp * 6)(*[_pd(x) for x in pr])
            m.psi_cor = (_dp * 6)(*[_pd(x) for x in pc])
        self.n_ref_nodes = int(n_cells_all) * self.NP
        self._ck(self.lib.nsem_upload_mesh(self.h, C.byref(m)))

    def set_bcs(self, bcs):
        """bcs: list of dicts {field, kind, faces, peer_faces?, value?, shape?, tvalue?, tshape?, zMin?, fixed?}"""
        arr = (NsemBC * max(1, len(bcs)))()
        keep = []
        for i, b in enumerate(bcs):
            r = arr[i]
            r.field = FIELDS[b["field"]] if isinstance(b["field"], str) else int(b["field"])
            r.kind = BC_KINDS[b["kind"]] if isinstance(b["kind"], str) else int(b["kind"])
            faces = _u32(b["faces"])
            keep.append(faces)
            r.n_faces = len(faces)
            r.faces = _pu(faces)
            if b.get("peer_faces") is not None:
                pf = _u32(b["peer_faces"])
                keep.append(pf)
                r.peer_faces = _pu(pf)
            val = np.zeros(3)
            v = np.atleast_1d(np.asarray(b.get("value", 0.0), dtype=float))
            val[: len(v)] = v
            tval = np.zeros(3)
            tv = np.atleast_1d(np.asarray(b.get("tvalue", 0.0), dtype=float))
            tval[: len(tv)] = tv
            r.value = (C.c_double * 3)(*val)
            r.tvalue = (C.c_double * 3)(*tval)
            r.shape = float(b.get("shape", 0.0))
            r.tshape = float(b.get("tshape", 0.0))
            r.zMin = float(b.get("zMin", 0.0))
            if b.get("fixed") is not None:
                fx = _f64(b["fixed"])
                keep.append(fx)
                r.fixed = _pd(fx)
        self._ck(self.lib.nsem_set_bcs(self.h, arr, len(bcs)))

    def set_params(self, *, P0, T0, cp, cv, viscosity, Pr, gravity, dt, buoyancy, diffusion):
        p = NsemParams(P0, T0, cp, cv, viscosity, Pr, (C.c_double * 3)(*gravity), dt, int(bool(buoyancy)), int(bool(diffusion)))
        self._ck(self.lib.nsem_set_params(self.h, C.byref(p)))

    def set_schedule(self, order):
        if order is None:
            self._ck(self.lib.nsem_set_schedule(self.h, None, 0))
        else:
            o = _u32(order)
            self._ck(self.lib.nsem_set_schedule(self.h, _pu(o), len(o)))

    # ---- state -----------------------------------------------------------------------------------
    def upload_state(self, rho, U, T, p=None):
        rho, U, T = _f64(rho), _f64(U), _f64(T)
        assert rho.size == self.n_ref_nodes and U.size == 3 * self.n_ref_nodes and T.size == self.n_ref_nodes
        pa = _f64(p) if p is not None else None
        self._ck(self.lib.nsem_upload_state(self.h, _pd(rho), _pd(U), _pd(T), _pd(pa)))

    def download_state(self):
        n = self.n_ref_nodes
        rho, U, T, p = np.zeros(n), np.zeros((n, 3)), np.zeros(n), np.zeros(n)
        self._ck(self.lib.nsem_download_state(self.h, _pd(rho), _pd(U), _pd(T), _pd(p)))
        return rho, U, T, p

    def download_gradients(self):
        """(gradf<strong>(U) as [n, 9] in the reference's Tensor order, gradf<strong>(T) as [n, 3]) of the last step."""
        n = self.n_ref_nodes
        gU, gT = np.zeros((n, 9)), np.zeros((n, 3))
        self._ck(self.lib.nsem_download_gradients(self.h, _pd(gU), _pd(gT)))
        return gU, gT

    # ---- explicit scalar advection (apps/convection) ----
    def upload_coords(self, cC):
        a = _f64(cC)
        assert a.size == self.n_ref_nodes * 3
        self._ck(self.lib.nsem_upload_coords(self.h, _pd(a)))

    def set_convection_scheme(self, scheme: str = "RUSANOV", blend_factor: float = 0.2):
        self._ck(self.lib.nsem_set_convection_scheme(self.h, {"RUSANOV": 0, "CDS": 1, "UDS": 2, "BLENDED": 3}[scheme], float(blend_factor)))

    def set_ab_order(self, order: int):
        self._ck(self.lib.nsem_set_ab_order(self.h, int(order)))

    def set_sphere(self, radius: float):
        self._ck(self.lib.nsem_set_sphere(self.h, float(radius)))

    def set_convection(self, problem_init: str = "NONE", etime: float = 1.0, first_step: int = 1):
        kinds = {"NONE": 0, "LEVEQUE": 1, "LAURITZEN_0": 2, "LAURITZEN_1": 3}
        self._ck(self.lib.nsem_set_convection(self.h, kinds[problem_init], float(etime), int(first_step)))

    def convection_step(self, nsteps=1):
        self._ck(self.lib.nsem_convection_step(self.h, int(nsteps)))

    # ---- operator-level views: one reference operator at a time on the current state ----
    def op_gradf_strong(self):
        n = self.n_ref_nodes
        gU, gT = np.zeros((n, 9)), np.zeros((n, 3))
        self._ck(self.lib.nsem_op_gradf_strong(self.h, _pd(gU), _pd(gT)))
        return gU, gT

    def op_divf_weak(self):
        """(r_rho [n], r_U [n,3], r_T [n]): divf<weak> of the three equations before src/addTemporal/Solve."""
        n = self.n_ref_nodes
        r, rU, rT = np.zeros(n), np.zeros((n, 3)), np.zeros(n)
        self._ck(self.lib.nsem_op_divf_weak(self.h, _pd(r), _pd(rU), _pd(rT)))
        return r, rU, rT

    def op_rusanov(self, n_cells_real: int, npf: int):
        """rusanov mass flux . fN per (element, local face, slot), owner's frame: [n_cells_real, 6, NPF]."""
        out = np.zeros((n_cells_real, 6, npf))
        self._ck(self.lib.nsem_op_rusanov(self.h, _pd(out)))
        return out

    def op_cds(self, field, n_cells_real: int, npf: int):
        f = _f64(field)
        assert f.size == self.n_ref_nodes
        out = np.zeros((n_cells_real, 6, npf))
        self._ck(self.lib.nsem_op_cds(self.h, _pd(f), _pd(out)))
        return out

    def op_apply_bcs(self, field: str, values):
        v = np.array(values, dtype=np.float64, order="C", copy=True)
        assert v.size == self.n_ref_nodes * (3 if field == "U" else 1)
        self._ck(self.lib.nsem_op_apply_bcs(self.h, {"rho": 0, "U": 2, "T": 3}[field], _pd(v)))
        return v

    def upload_ref(self, rho_ref, p_ref, g=None, gh=None):
        a, b = _f64(rho_ref), _f64(p_ref)
        gg = _f64(g) if g is not None else None
        self._ck(self.lib.nsem_upload_ref(self.h, _pd(a), _pd(b), _pd(gg)))
        if gh is not None:
            hh = _f64(gh)
            self._ck(self.lib.nsem_upload_geopotential(self.h, _pd(hh)))

    # ---- hot path ----------------------------------------------------------------------------------
    def step(self, nsteps=1):
        self._ck(self.lib.nsem_euler_step(self.h, int(nsteps)))

    def sync(self):
        self._ck(self.lib.nsem_sync(self.h))

    def time_steps(self, nsteps, per_kernel=False):
        ms = C.c_double()
        pk = np.zeros(4)
        self._ck(self.lib.nsem_time_steps(self.h, int(nsteps), C.byref(ms), _pd(pk) if per_kernel else None))
        return ms.value, pk

    def diagnostics(self):
        out = np.zeros(6)
        self._ck(self.lib.nsem_diagnostics(self.h, _pd(out)))
        return out

    @property
    def launch_count(self):
        return int(self.lib.nsem_launch_count(self.h))

    @property
    def kernel_info(self) -> str:
        return self.lib.nsem_kernel_info(self.h).decode()
