"""TEST DOUBLE: a backend with the interface of autobz_b200.backend.DeviceBackend whose arithmetic is the
CPU oracle.  It exists so that the HOST control flow (tolerance rescaling, symmetrisation, AutoPTR loop,
IAI engine, multi-rank sharding + allreduce) can be tested on a box without a GPU and compared with the
reference's known-answer tests.  It lives under tests/ and is never imported by the product."""
import ctypes as C

import numpy as np

import orc
from autobz_b200.backend import share_planes, symptr_nodes_lowdim


def _oseries(fs):
    c = np.asarray(fs.c, dtype=np.complex128)
    while c.ndim < 5:
        c = c[..., None]
    lo = tuple(fs.lo) + (0,) * (3 - fs.ndim)
    per = tuple(fs.period) + (1.0,) * (3 - fs.ndim)
    return orc.Series(c, lo, per)


class OracleRule:
    def __init__(self, series, ndim, npt, syms, rank=0, nranks=1):
        self.fs, self.so, self.ndim, self.npt = series, _oseries(series), ndim, int(npt)
        self._syms = syms
        self.nsyms = 1 if syms is None else len(syms)
        if ndim == 3:
            if syms is None:
                i3, i2, i1 = np.meshgrid(np.arange(npt), np.arange(npt), np.arange(npt), indexing="ij")
                idx = np.stack([i1.ravel(), i2.ravel(), i3.ravel()], axis=1)
                w = np.ones(idx.shape[0])
                lo, hi = (npt * rank) // nranks, (npt * (rank + 1)) // nranks
                sel = (idx[:, 2] >= lo) & (idx[:, 2] < hi)
            else:
                sy = np.array([np.rint(np.asarray(S)).astype(np.int32) for S in syms])
                ws, nirr = orc.symptr_rule(npt, sy)
                i1, i2, i3 = np.nonzero(ws)
                order = np.lexsort((i1, i2, i3))
                idx = np.stack([i1[order], i2[order], i3[order]], axis=1)
                w = ws[i1[order], i2[order], i3[order]].astype(float)
                # the product's dealing of the k3 planes of a symmetric rule (backend.make_rule: serpentine for nranks > 1)
                sel = np.isin(idx[:, 2], share_planes(npt, rank, -nranks if nranks > 1 else 1))
            self.nnodes_total = idx.shape[0]
            self.idx, self.w = idx[sel], w[sel]
            self._sel = sel
        else:
            idx, w = symptr_nodes_lowdim(npt, ndim, syms)
            self.nnodes_total = idx.shape[0]
            lo, hi = (idx.shape[0] * rank) // nranks, (idx.shape[0] * (rank + 1)) // nranks
            self.idx, self.w = idx[lo:hi], w[lo:hi]
            self._sel = slice(lo, hi)
        self.nnodes = self.idx.shape[0]
        self._H = None

    def __len__(self):
        return self.nnodes_total

    def _Hk(self):
        if self._H is None:
            self._H = orc.eval_points(self.so, self.idx / float(self.npt))
        return self._H

    def materialize(self):
        self._Hk()

    def copy_out(self):
        return self._Hk(), self.idx / float(self.npt), self.w

    def resolvent_sum(self, z, sigma, fkind):
        H = self._Hk()
        if self.nnodes == 0:
            return np.zeros(1 if fkind == 1 else np.atleast_1d(z).size, dtype=complex)
        if fkind == 1:
            return np.array([np.sum(self.w * np.trace(H, axis1=0, axis2=1))])
        y = orc.resolvent_trace_batch(H, z, sigma)
        return (self.w[:, None] * y).sum(axis=0)

    def resolvent_matrix_sum(self, z, sigma):
        H = np.moveaxis(self._Hk(), 2, 0)
        n = H.shape[1]
        zs = np.atleast_1d(z)
        out = np.zeros((zs.size, n, n), dtype=complex)
        for w, zz in enumerate(zs):
            sg = 0 if sigma is None else np.asarray(sigma).reshape(n, n, zs.size)[:, :, w]
            if self.nnodes:
                out[w] = np.tensordot(self.w, np.linalg.inv(zz * np.eye(n) - H - sg), axes=(0, 0))
        return out

    def eig_sum(self, kind, params):
        ev = orc.eigvals_batch(self._Hk())
        p0, p1 = params
        f = lambda x: np.where(x > 0, np.exp(-np.abs(x)) / (1 + np.exp(-np.abs(x))), 1 / (1 + np.exp(-np.abs(x))))
        if kind == 0:
            g = ev
        elif kind == 1:
            g = ev * f((ev - p0) / p1)
        elif kind == 2:
            g = f((ev - p0) / p1)
        else:
            g = np.exp(-((ev - p0) / p1) ** 2) / (p1 * np.sqrt(np.pi))
        return float(np.sum(self.w * g.sum(axis=1)))

    def ggr_data(self, ndim, copy=True):
        # the oracle's data pass runs over the rule's full node set; keep this rank's nodes
        if self.ndim == 3 and self._syms is not None:
            sy = np.array([np.rint(np.asarray(S)).astype(np.int32) for S in self._syms])
            ws, _ = orc.symptr_rule(self.npt, sy)
        elif self._syms is None:
            ws = None
        else:
            ws = np.zeros((self.npt,) * self.ndim + (1,) * (3 - self.ndim), dtype=np.int32, order="F")
            idx, w = symptr_nodes_lowdim(self.npt, self.ndim, self._syms)
            ws[tuple(idx[:, d] for d in range(self.ndim)) + (0,) * (3 - self.ndim)] = w.astype(np.int32)
        w, e, v = orc.ggr_data(self.so, self.ndim, self.npt, ws)
        sel = self._sel
        self._ggr = (w[sel], e[sel], v[sel])
        return (self._ggr[1], self._ggr[2]) if copy else (None, None)

    def ggr_sum(self, E):
        w, e, v = self._ggr
        if w.size == 0:
            return np.zeros(np.atleast_1d(E).size)
        return orc.ggr_sum(self.ndim, self.npt, E, w, e, v)

    def close(self):
        pass


class OracleNest:
    def __init__(self, series, ndim):
        self.so, self.ndim = _oseries(series), ndim
        self.L2, self.L1 = {}, {}
        self.n = series.norb

    def _contract(self, src, rows, M, lo, period, x):
        out = np.empty(rows, dtype=np.complex128)
        orc.lib().orc_contract(src.ctypes.data_as(orc.c_dp), C.c_long(rows), C.c_int(M), C.c_int(lo), C.c_double(period),
                               C.c_double(x), out.ctypes.data_as(orc.c_dp))
        return out

    def contract3(self, x3, slot2):
        s = self.so
        rows = s.n * s.n * s.M[0] * s.M[1]
        flat = np.ascontiguousarray(s.c.reshape(-1, order="F"))
        for x, sl in zip(x3, slot2):
            self.L2[int(sl)] = self._contract(flat, rows, s.M[2], s.lo[2], s.period[2], float(x))

    def contract2(self, x2, parent, slot1):
        s = self.so
        rows = s.n * s.n * s.M[0]
        root = np.ascontiguousarray(s.c.reshape(-1, order="F"))
        for i, (x, sl) in enumerate(zip(x2, slot1)):
            src = root if parent is None else self.L2[int(parent[i])]
            self.L1[int(sl)] = self._contract(src, rows, s.M[1], s.lo[1], s.period[1], float(x))

    def eval_h(self, x1, slot1):
        s = self.so
        nn = s.n * s.n
        root = np.ascontiguousarray(s.c.reshape(-1, order="F"))
        H = np.empty((s.n, s.n, len(x1)), dtype=np.complex128, order="F")
        for i, x in enumerate(x1):
            src = root if slot1 is None else self.L1[int(slot1[i])]
            H[:, :, i] = self._contract(src, nn, s.M[0], s.lo[0], s.period[0], float(x)).reshape(s.n, s.n, order="F")
        return H

    def eval_matrix(self, x1, slot1, z, sigma=None):
        """(z - H - Sigma)^-1 at the nodes -> [npts, n, n] (numpy / LAPACK inverse: the reference's `inv` on a Matrix)"""
        H = np.moveaxis(self.eval_h(x1, slot1), 2, 0)
        n = self.so.n
        A = complex(np.atleast_1d(z)[0]) * np.eye(n)[None] - H
        if sigma is not None:
            A = A - np.asarray(sigma, dtype=np.complex128).reshape(n, n)[None]
        return np.linalg.inv(A)

    def eval(self, x1, slot1, z, sigma, fkind):
        s = self.so
        nn = s.n * s.n
        root = np.ascontiguousarray(s.c.reshape(-1, order="F"))
        H = np.empty((s.n, s.n, len(x1)), dtype=np.complex128, order="F")
        for i, x in enumerate(x1):
            src = root if slot1 is None else self.L1[int(slot1[i])]
            H[:, :, i] = self._contract(src, nn, s.M[0], s.lo[0], s.period[0], float(x)).reshape(s.n, s.n, order="F")
        if fkind == 1:
            return np.trace(H, axis1=0, axis2=1).astype(np.complex128)
        sg = None if sigma is None else np.asarray(sigma).reshape(s.n, s.n, 1)
        return orc.resolvent_trace_batch(H, [z], sg)[:, 0]


class OracleBackend:
    launch_count = 0

    def make_rule(self, series, ndim, npt, syms, rank=0, nranks=1, allreduce=None):
        return OracleRule(series, ndim, npt, syms, rank, nranks)

    def make_nest(self, series, ndim, cap2, cap1):
        return OracleNest(series, ndim)
