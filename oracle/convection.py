"""numpy restatement of the reference's explicit scalar advection app -- TEST INFRASTRUCTURE ONLY (the checker of
nsem_convection_step; nothing under nebulasem_b200/ imports it).

apps/convection/convection.cpp:18-150:  dT/dt + div(T U) = 0 on the dGSEM operators of the euler path:
    Fc = flxc(U) = U, lambdaMax = cds(mag(U)) / 2                                   (:103-105, 119-121)
    M = divf(Fc * T, false, &F, &T, &lambdaMax); addTemporal<1>(M, t_UR); Solve(M)    (:129-135)
with the optional analytic wind of LeVeque's deformation test, or Lauritzen's two on the sphere, re-evaluated at every step (:42-87, 114-121).  One reference step is one
forward-Euler stage for BDF1 / AB1 / RK1 (SURVEY finding 1), like the euler path."""
from __future__ import annotations

import numpy as np

from . import libm
from .euler import EulerOracle, vdot, vmag

PI = 3.14159265358979323846264          # Constants::PI


class ConvectionOracle(EulerOracle):
    def setup_convection(self, T, U, bcs: dict, problem_init: str = "NONE", end_step: int = 1):
        self.bcs["T"] = bcs.get("T", [])
        self.bcs["U"] = bcs.get("U", [])
        self.T, self.U = T.copy(), U.copy()
        # MeshField::read_ applies the BCs right after reading (field.h:1562-1565)
        self.apply_bcs("U", self.U)
        self.apply_bcs("T", self.T)
        self.problem_init = problem_init
        self.end_step = end_step
        self.scalar0 = float(np.sum((self.T * self.g.cV)[: self.gB]))

    def wind(self, time: float, etime: float):
        """init_wind_field, LEVEQUE branch (convection.cpp:74-82), on EVERY entry of U (ghost cells carry their owner's coordinates)."""
        x, y = self.g.cC[:, 0], self.g.cC[:, 1]
        period = etime
        ct = libm.cos_(np.array([PI * time / period]))[0]
        if self.problem_init.startswith("LAURITZEN"):
            # deformational flow on the sphere (convection.cpp:59-72; wind_field, tensor.h:615-621)
            if not self.p.is_spherical:
                raise NotImplementedError("the Lauritzen winds need is_spherical YES (the reference leaves u, v unset otherwise)")
            z = self.g.cC[:, 2]
            lat = libm.atan2_(z, libm.sqrt_(x * x + y * y))
            lon = libm.atan2_(y, x)
            RoT = self.p.sphere_radius / period
            lam = lon - 2.0 * PI * time / period
            if self.problem_init == "LAURITZEN_0":
                u = 10.0 * RoT * libm.pow_(libm.sin_(lam), 2.0) * libm.sin_(2.0 * lat) * ct + 2.0 * PI * RoT * libm.cos_(lat)
                v = 10.0 * RoT * libm.sin_(2.0 * lam) * libm.cos_(lat) * ct
            else:
                u = -5.0 * RoT * libm.pow_(libm.sin_(0.5 * lam), 2.0) * libm.sin_(2.0 * lat) * libm.pow_(libm.cos_(lat), 2.0) * ct \
                    + 2.0 * PI * RoT * libm.cos_(lat)
                v = 2.5 * RoT * libm.sin_(lam) * libm.pow_(libm.cos_(lat), 3.0) * ct
            return np.stack([-u * libm.sin_(lon) - v * libm.sin_(lat) * libm.cos_(lon),
                             +u * libm.cos_(lon) - v * libm.sin_(lat) * libm.sin_(lon),
                             +v * libm.cos_(lat)], axis=1)
        u = libm.pow_(libm.sin_(PI * x), 2.0) * libm.sin_(2 * PI * y) * ct
        v = -libm.pow_(libm.sin_(PI * y), 2.0) * libm.sin_(2 * PI * x) * ct
        return np.stack([u, v, np.zeros_like(u)], axis=1)

    def step(self):
        P = self.p
        i = self.step_count + 1                                      # Iteration::get_step() inside the loop
        if self.problem_init != "NONE":
            self.U = self.wind(i * P.dt, self.end_step * P.dt)
        lam = self.cds(vmag(self.U)) / 2
        fq = self.U * self.T[:, None]
        sch = P.convection_scheme
        if sch == "RUSANOV":
            r = self.divf(fq, self.T, lam)
        else:
            # F = flx(U) = dot(cds(U), fN) (field.h:3396-3399), the sign that picks the upwind side
            F = vdot(self.cds(self.U), self.fNv)
            r = self.divf(fq, self.T, lam, scheme=sch, flux=F, blend=P.blend_factor)
        ap0 = (-1.0 / P.dt) * self.g.cV
        order = int(P.time_scheme[2]) if P.time_scheme.startswith("AB") else 1
        if order > 1:
            # AB2..AB5 (ddt, field.h:3789-3806; history of addTemporal, :3885-3905): the field's first step fills the whole history with its
            # residual (initStore, nstored = 1), later ones shift it (updateStore); the order applied is min(scheme, nstored)
            hist = getattr(self, "_ab_hist", None)
            if hist is None:
                self._ab_hist, self._ab_stored = [r.copy() for _ in range(order)], 1
            else:
                self._ab_hist = [r.copy()] + hist[:-1]
                self._ab_stored += 1
            h, use = self._ab_hist, min(order, self._ab_stored)
            if use == 5:
                comb = (1901 * h[0] - 2774 * h[1] + 2616 * h[2] - 1274 * h[3] + 251 * h[4]) / 720.0
            elif use == 4:
                comb = (55 * h[0] - 59 * h[1] + 37 * h[2] - 9 * h[3]) / 24.0
            elif use == 3:
                comb = (23 * h[0] - 16 * h[1] + 5 * h[2]) / 12.0
            elif use == 2:
                comb = (3 * h[0] - h[1]) / 2.0
            else:
                comb = h[0]
            r = comb
        Su = (r + self.T * ap0) if P.time_scheme.startswith("BDF") else (self.T * ap0 + r)
        T = Su / ap0
        self.apply_bcs("T", T)
        self.T = T
        self.step_count += 1
