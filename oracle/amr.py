"""AMR field transfer -- numpy restatement of MeshField::refineField (src/field/field.h:1863-2015). TEST INFRASTRUCTURE ONLY.

Pinned: bit-identical to the reference on tests/golden/refine_field/*.npz (two regrids each of a 2-D order-4 and a 3-D order-2 case
driven through the reference's own refineMesh/refineField by oracle/tools/refinedump.cpp; tests/test_refine_field.py).
The floating-point operation order of the reference is kept: every output node accumulates its contributions in (child, input node)
order, products are ((P*fx)*fy)*fz, and the mass-fix sums run sequentially over (child, input node, output node).
"""
from __future__ import annotations

import numpy as np

MAX_INT = 1 << 31                          # Constants::MAX_INT (tensor.h:455), the "cell removed" mark of cellMap


def child_half(cC_cell: np.ndarray, ccc: np.ndarray, ccp: np.ndarray, npts) -> tuple:
    """Which half of the parent a child covers along each element axis (field.h:1899-1909 / 1949-1959): corner nodes of the cell
    whose node coordinates are in use (the child when coarsening, the parent when refining), centroids of child and parent."""
    NPX, NPY, NPZ = npts
    c = cC_cell.reshape(NPX, NPY, NPZ, 3)
    v0 = c[0, 0, 0]
    offs = []
    for v in (c[NPX - 1, 0, 0], c[0, NPY - 1, 0], c[0, 0, NPZ - 1]):
        e = v - v0
        a = ccc - v0
        b = ccp - v0
        da = a[0] * e[0] + a[1] * e[1] + a[2] * e[2]
        db = b[0] * e[0] + b[1] * e[1] + b[2] * e[2]
        offs.append(0 if da <= db else 1)
    return tuple(offs)


def _families(m: np.ndarray):
    i, m = 0, np.asarray(m, dtype=np.int64)
    while i < len(m):
        n = int(m[i])
        yield int(m[i + 1]), m[i + 2:i + 2 + n]
        i += n + 2


def _seq_sum(start: np.ndarray, terms: np.ndarray) -> np.ndarray:
    """start + terms[0] + terms[1] + ... strictly left to right (np.cumsum does not reassociate), per component."""
    return np.cumsum(np.concatenate([start[None, :], terms], axis=0), axis=0)[-1]


def refine_field(P: np.ndarray, npts, refineMap, coarseMap, cellMap, nCells: int, oldCV, oldCC, newCC, newCV, cC_old,
                 psiRef, psiCor, wgl) -> np.ndarray:
    """P: [old real cells*NP, comps] -> [nCells*NP, comps].  psiRef/psiCor: six [n*n] tables (DG::psiRef[d*2+half]), wgl: three."""
    NPX, NPY, NPZ = npts
    NP = NPX * NPY * NPZ
    P = np.asarray(P, dtype=np.float64)
    if P.ndim == 1:
        P = P[:, None]
    comps = P.shape[1]
    cellMap = np.asarray(cellMap, dtype=np.int64)
    oldCC = np.asarray(oldCC).reshape(-1, 3)
    newCC = np.asarray(newCC).reshape(-1, 3)
    cC_old = np.asarray(cC_old).reshape(-1, NP, 3)
    nOld = P.shape[0] // NP
    Pc = P.reshape(nOld, NP, comps)
    Pn = np.zeros((nCells, NP, comps))
    # copy (field.h:1878-1884)
    for i in range(nOld):
        if cellMap[i] != MAX_INT:
            Pn[cellMap[i]] = Pc[i]
    # node index -> (ii, jj, kk) and the quadrature weight of a node, ((w0*w1)*w2)/8 (field.h:1969, 1982)
    ii, jj, kk = np.meshgrid(np.arange(NPX), np.arange(NPY), np.arange(NPZ), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    wnode = ((wgl[0][ii] * wgl[1][jj]) * wgl[2][kk]) / 8

    def tables(tab, off):
        # F[in, out] factors per direction gathered for all (input node, output node) pairs
        fx = tab[0 * 2 + off[0]].reshape(NPX, NPX)[ii[:, None], ii[None, :]]
        fy = tab[1 * 2 + off[1]].reshape(NPY, NPY)[jj[:, None], jj[None, :]]
        fz = tab[2 * 2 + off[2]].reshape(NPZ, NPZ)[kk[:, None], kk[None, :]]
        return fx, fy, fz

    # coarsening: volume-weighted projection of the children (field.h:1887-1934)
    for first, kids in _families(coarseMap):
        cid = cellMap[first]
        ccp = newCC[cid]
        acc = np.zeros((NP, comps))
        vol = 0.0
        for id1 in kids:
            off = child_half(cC_old[id1], oldCC[id1], ccp, npts)
            fx, fy, fz = tables(psiCor, off)
            P0 = Pc[id1] * oldCV[id1]                                   # [in, comps]
            for q in range(NP):                                          # input nodes in order; outputs in parallel
                acc += ((P0[q][None, :] * fx[q][:, None]) * fy[q][:, None]) * fz[q][:, None]
            vol += oldCV[id1]
        Pn[cid] = acc / vol
    # refinement: interpolation onto the children + one factor per family that restores the parent's integral (field.h:1937-2000)
    for pid, kids in _families(refineMap):
        ccp = oldCC[pid]
        toto = np.zeros(comps)
        totn = np.zeros(comps)
        new_ids = cellMap[kids]
        for j, id1 in enumerate(new_ids):
            off = child_half(cC_old[pid], newCC[id1], ccp, npts)
            fx, fy, fz = tables(psiRef, off)
            P0 = Pc[pid]
            if j == 0:
                toto = _seq_sum(toto, (P0 * oldCV[pid]) * wnode[:, None])
            acc = np.zeros((NP, comps))
            terms = np.empty((NP, NP, comps))
            for q in range(NP):
                P1 = ((P0[q][None, :] * fx[q][:, None]) * fy[q][:, None]) * fz[q][:, None]       # [out, comps]
                acc += P1
                terms[q] = (P1 * newCV[id1]) * wnode[:, None]
            totn = _seq_sum(totn, terms.reshape(NP * NP, comps))
            Pn[id1] = acc
        mo = np.sqrt((toto * toto).sum()) if comps > 1 else abs(toto[0])
        mn = np.sqrt((totn * totn).sum()) if comps > 1 else abs(totn[0])
        factor = mo / mn if mn else 0.0
        for id1 in new_ids:
            Pn[id1] *= factor
    return Pn.reshape(nCells * NP, comps)
