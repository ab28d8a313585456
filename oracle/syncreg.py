"""TEST INFRASTRUCTURE ONLY (oracle): a numpy / pure-Python restatement of IAMR's SyncRegister (Source/SyncRegister.cpp), kept in the
reference's own data structure -- one register fab per face and per coarsened fine grid (the FabSets bndry[face]) -- so that it
shares nothing with the library's dense coarse-level representation (iamr_b200/csrc/amr.cu).  Only tests/ may import this file.

Unlike the AMReX-internal operators, this algorithm is IAMR's own and fully visible in the reference tree: every step below cites the
lines it follows.  It cannot be run against the reference here (SyncRegister.cpp needs AMReX to compile), so the parity claim is
"restated from the visible source", not "pinned by reference output".

Conventions: coarse node arrays are dense over the node domain, shape (nz + 1, ny + 1, nx + 1), index [k, j, i]; boxes are
(lo, hi) tuples of (i, j, k) cell indices; refinement ratio 2.  FabSet::plusFrom over a source whose fabs share nodes (the coarse
level's own grids, or a periodic source that contains a node and its image) is taken to count every physical node once in
CrseInit -- see DESIGN.md 4a -- and to ADD all overlapping coarsened-fine fabs and periodic images in FineAdd, which the edge /
corner weights of FineAdd are built for.
"""
import itertools

import numpy as np


def _shifts(n, per):
    out = []
    for s in itertools.product((-1, 0, 1), repeat=3):
        if all(s[d] == 0 or per[d] for d in range(3)):
            out.append(tuple(s[d] * n[d] for d in range(3)))
    return out


class SyncRegister:
    def __init__(self, nc, per, fine_boxes):
        """SyncRegister::SyncRegister (:20-47): grids = coarsened fine boxes; one nodal register fab per face and grid."""
        self.nc, self.per = tuple(nc), tuple(per)
        self.grids = [(tuple(l // 2 for l in lo), tuple((h + 1) // 2 - 1 for h in hi)) for lo, hi in fine_boxes]
        self.fabs = {}   # (dir, side, grid) -> [lo, hi, array[k, j, i]] on the node plane of that face
        for g, (lo, hi) in enumerate(self.grids):
            for d in range(3):
                for side in range(2):
                    nlo, nhi = list(lo), [h + 1 for h in hi]
                    nlo[d] = nhi[d] = lo[d] if side == 0 else hi[d] + 1
                    shape = tuple(nhi[q] - nlo[q] + 1 for q in (2, 1, 0))
                    self.fabs[(d, side, g)] = [tuple(nlo), tuple(nhi), np.zeros(shape)]

    @staticmethod
    def _nodes(lo, hi):
        return itertools.product(range(lo[2], hi[2] + 1), range(lo[1], hi[1] + 1), range(lo[0], hi[0] + 1))

    def crse_init(self, resid_crse, mult):
        """CrseInit (:288-300): setVal(0); Sync_resid_crse.mult(mult); bndry[face].plusFrom(Sync_resid_crse, periodicity)."""
        for lo, hi, a in self.fabs.values():
            a[...] = 0.0
            for k, j, i in self._nodes(lo, hi):
                a[k - lo[2], j - lo[1], i - lo[0]] += mult * resid_crse[k, j, i]

    def fine_add(self, fine_boxes, resid_fine, mult):
        """FineAdd (:351-607).  resid_fine[g]: node array of fine grid g (valid nodes, shape (nz + 1, ny + 1, nx + 1))."""
        nc, per = self.nc, self.per
        crse_fabs = []
        for g, ((flo, fhi), rf) in enumerate(zip(fine_boxes, resid_fine)):
            nlo, nhi = flo, tuple(h + 1 for h in fhi)            # fine node box
            fab = np.zeros(tuple(nhi[q] - nlo[q] + 3 for q in (2, 1, 0)))     # one ghost node, zero (Projection.cpp:374-376)
            fab[1:-1, 1:-1, 1:-1] = mult * rf                     # Sync_resid_fine.mult(mult) (:355)
            F = lambda i, j, k: fab[k - nlo[2] + 1, j - nlo[1] + 1, i - nlo[0] + 1]
            # :369-418 the twelve edges count half, the eight corners a further two thirds
            w = np.ones(fab.shape)
            for k, j, i in self._nodes(nlo, nhi):
                nb = (i in (nlo[0], nhi[0])) + (j in (nlo[1], nhi[1])) + (k in (nlo[2], nhi[2]))
                if nb >= 2:
                    w[k - nlo[2] + 1, j - nlo[1] + 1, i - nlo[0] + 1] *= 0.5
                if nb == 3:
                    w[k - nlo[2] + 1, j - nlo[1] + 1, i - nlo[0] + 1] *= 2.0 / 3.0
            fab *= w
            clo, chi = self.grids[g][0], tuple(h + 1 for h in self.grids[g][1])   # coarsened node box
            crse = np.zeros(tuple(chi[q] - clo[q] + 1 for q in (2, 1, 0)))
            r = 2
            denom = r / float(r ** 6)
            for d in range(3):                                   # :425-550
                dim1 = 0 if d != 0 else 1
                dim2 = (1 if d == 2 else 2) if d != 0 else 2
                for side in range(2):
                    plo, phi = list(clo), list(chi)
                    plo[d] = phi[d] = clo[d] if side == 0 else chi[d]
                    for kc, jc, ic in self._nodes(plo, phi):
                        idxc = (ic, jc, kc)
                        v = 0.0
                        for n in range(r):
                            for m in range(r):
                                coeff = (r - m) * (r - n) * denom
                                if n == 0:
                                    coeff *= 0.5
                                if m == 0:
                                    coeff *= 0.5
                                for s1, s2 in ((1, 1), (-1, 1), (1, -1), (-1, -1)):
                                    idx = [r * idxc[0], r * idxc[1], r * idxc[2]]
                                    idx[dim1] += s1 * m
                                    idx[dim2] += s2 * n
                                    v += coeff * F(*idx)
                        for q in range(3):                       # :505-534 doubled on non-periodic domain planes
                            if not per[q] and idxc[q] in (0, nc[q]):
                                v *= 2.0
                        crse[kc - clo[2], jc - clo[1], ic - clo[0]] += v      # :538-547 crsefab += cbndfab
            crse_fabs.append((clo, chi, crse))
        # :603-606 bndry[face].plusFrom(Sync_resid_crse, periodicity): ADD of every overlapping coarsened-fine fab and image
        sh = _shifts(nc, per)
        for lo, hi, a in self.fabs.values():
            for k, j, i in self._nodes(lo, hi):
                for clo, chi, crse in crse_fabs:
                    for s in sh:
                        q = (i + s[0], j + s[1], k + s[2])
                        if all(clo[t] <= q[t] <= chi[t] for t in range(3)):
                            a[k - lo[2], j - lo[1], i - lo[0]] += crse[q[2] - clo[2], q[1] - clo[1], q[0] - clo[0]]

    def init_rhs(self, phys_lo=(0, 0, 0), phys_hi=(0, 0, 0), maxcount=3 * 3 * 3 - 0.5):
        """InitRHS (:49-285) on a coarse level that is one grid over the domain; returns the node array.  maxcount: the reference's
        AMREX_D_TERM(SPACEDIM, *SPACEDIM, *SPACEDIM) - 0.5 (:265) = 26.5 in 3-D (never exceeded: the mask is all ones there)."""
        nc, per = self.nc, self.per
        rhs = np.zeros((nc[2] + 1, nc[1] + 1, nc[0] + 1))
        sh = _shifts(nc, per)
        inside = lambda q: all(0 <= q[t] <= nc[t] for t in range(3))
        for lo, hi, a in self.fabs.values():                       # :61-64 copyTo(rhs, periodicity)
            for k, j, i in self._nodes(lo, hi):
                for s in sh:
                    q = (i + s[0], j + s[1], k + s[2])
                    if inside(q):
                        rhs[q[2], q[1], q[0]] = a[k - lo[2], j - lo[1], i - lo[0]]
        for d in range(3):                                         # :66-128 zero on outflow planes
            if phys_lo[d] == 2:
                sl = [slice(None)] * 3; sl[2 - d] = 0; rhs[tuple(sl)] = 0.0
            if phys_hi[d] == 2:
                sl = [slice(None)] * 3; sl[2 - d] = nc[d]; rhs[tuple(sl)] = 0.0
        # :126-272 bndry_mask: 0 where all eight cells around the node lie under the fine grids
        def covered(ci, cj, ck):
            for s in sh:                                          # :169-186 periodic shifts of the mask cells
                q = (ci + s[0], cj + s[1], ck + s[2])
                for glo, ghi in self.grids:
                    if all(glo[t] <= q[t] <= ghi[t] for t in range(3)):
                        return 1.0
            return 0.0
        for lo, hi, a in self.fabs.values():
            for k, j, i in self._nodes(lo, hi):
                cnt = sum(covered(i - di, j - dj, k - dk) for dk in (0, 1) for dj in (0, 1) for di in (0, 1))
                for q, idx in enumerate((i, j, k)):               # :204-250 doubled on non-periodic domain planes
                    if not per[q] and idx in (0, nc[q]):
                        cnt *= 2.0
                mask = 0.0 if cnt > maxcount else 1.0             # :253-272
                rhs[k, j, i] *= mask                              # :274-284 bndry_mask.copyTo(tmp) without periodicity
        return rhs
