"""Shared test plumbing: build a case, run the numpy oracle, mirror it on the device through the C ABI."""
from __future__ import annotations

import os

import numpy as np

from oracle import case as ocase
from oracle import cases as ocases
from oracle.euler import EulerOracle


def make_oracle(tmpdir, name, nsteps_hint=1, exact=False, **kw) -> EulerOracle:
    c = ocases.CASES[name](**kw)
    d = os.path.join(str(tmpdir), f"{name}_" + "_".join(f"{k}{v}" for k, v in sorted(kw.items())))
    c.write(d, max(1, nsteps_hint))
    orc = ocase.load_case(d, exact_order=exact)
    orc.case_dir = d
    return orc


def device_bcs(orc: EulerOracle):
    out = []
    for fld in ("rho", "p", "U", "T"):
        for bc in orc.bcs[fld]:
            kind = bc.kind
            d = dict(field=fld, kind=kind, faces=bc.faces, value=bc.value, shape=bc.shape,
                     tvalue=bc.tvalue if bc.tvalue is not None else 0.0, tshape=bc.tshape, zMin=bc.zMin)
            if kind == "CYCLIC":
                d["peer_faces"] = bc.neighbor_faces
            if kind == "CALC_DIRICHLET":
                d["fixed"] = bc.fixed
            out.append(d)
    return out


def device_from_oracle(orc: EulerOracle, device=0):
    """Create an nsem Context holding exactly the oracle's mesh, BCs, parameters and CURRENT state."""
    from nebulasem_b200 import capi
    g, b, t = orc.g, orc.g.basis, orc.g.topo
    ctx = capi.Context(device)
    ctx.set_order(b.NPX, b.NPY, b.NPZ)
    ctx.set_basis(b.dpsi, b.wgl)
    face_id = np.concatenate([np.asarray(x, dtype=np.uint32) for x in t.faceID])
    ctx.upload_mesh(n_cells_real=g.nBCS, n_cells_all=g.nCells, n_faces=g.nFacets, cV=g.cV, Jinv=g.Jinv, fN=g.fN, fI=g.fI,
                    face_normal=t.FNv, FO=g.FO, FN=g.FN, face_begin=g.faceIndices[0], face_end=g.faceIndices[1],
                    all_faces=g.allFaces, face_id=face_id, face_owner=t.FOC, face_neigh=t.FNC, face_mortar=t.FMC,
                    cC=g.cC, face_center=t.FC, psi_ref=b.psiRef, psi_cor=b.psiCor)
    ctx.set_bcs(device_bcs(orc))
    p = orc.p
    ctx.set_params(P0=p.P0, T0=p.T0, cp=p.cp, cv=p.cv, viscosity=p.viscosity, Pr=p.Pr, gravity=p.gravity, dt=p.dt,
                   buoyancy=p.buoyancy, diffusion=p.diffusion)
    ctx.upload_ref(orc.rho_ref, orc.p_ref, None)
    ctx.upload_state(orc.rho, orc.U, orc.T, orc.pp)
    return ctx


def rel_l2(a, b, scale=None):
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    den = np.linalg.norm(b) if scale is None else scale
    return float(np.linalg.norm(a - b) / max(den, 1e-300))


def conserved_errors(orc: EulerOracle, rho, U, T):
    """Relative L2 errors of (rho, rho*U, rho*theta) over the real nodes; rho*U also against the momentum scale
    ||rho|| * c0 with c0 = sqrt(gamma R T0) (SURVEY finding 6)."""
    nb = orc.gB
    T0 = orc.p.T0
    r_o, U_o, th_o = orc.rho[:nb], orc.U[:nb], orc.T[:nb] + T0
    r_d, U_d, th_d = rho[:nb], U[:nb], T[:nb] + T0
    c0 = np.sqrt(orc.gamma * orc.R * T0)
    return dict(rho=rel_l2(r_d, r_o), rhoU_self=rel_l2(r_d[:, None] * U_d, r_o[:, None] * U_o),
                rhoU_scaled=rel_l2(r_d[:, None] * U_d, r_o[:, None] * U_o, scale=np.linalg.norm(r_o) * c0 * np.sqrt(3.0) / np.sqrt(3.0)),
                rhoTheta=rel_l2(r_d * th_d, r_o * th_o))
