"""Device backend: the seam between the host control flow (rules.py, iai.py, interfaces.py) and
libautobz_cuda.so.  The control flow is written against this small duck-typed interface so that
the multi-rank host logic can be exercised without a GPU by a test double (tests/ only); the product
always uses DeviceBackend, which raises when the CUDA library or a B200 is missing."""
import os

import numpy as np

from . import _lib

_default_ctx = {}


def default_context(device=None):
    """One abz_ctx per device for this process (device defaults to LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx or _default_ctx[device].h is None:
        _default_ctx[device] = _lib.Context(device)
    return _default_ctx[device]


def embed_syms(syms, ndim):
    """d x d lattice-basis symmetry matrices embedded in 3 x 3 (identity on the unused dimensions)."""
    out = []
    for S in syms:
        S = np.asarray(S)
        E = np.eye(3, dtype=np.int32)
        E[:ndim, :ndim] = np.rint(S).astype(np.int32)
        if not np.allclose(S, E[:ndim, :ndim]):
            raise ValueError("symmetries must be integer matrices in the lattice basis")
        out.append(E)
    return np.array(out, dtype=np.int32)


def symptr_nodes_lowdim(npt, ndim, syms):
    """AutoSymPTR.symptr_rule for 1-d / 2-d grids on the host (control plane; tiny): returns
    (idx [nnodes,3] int32 sorted by (i3,i2,i1), w [nnodes])."""
    if ndim == 1:
        grid = np.arange(npt)[:, None]
    else:
        i2, i1 = np.meshgrid(np.arange(npt), np.arange(npt), indexing="ij")
        grid = np.stack([i1.ravel(), i2.ravel()], axis=1)
    lin = grid[:, 0] + (npt * grid[:, 1] if ndim == 2 else 0)
    if syms is None:
        w = np.ones(grid.shape[0])
        keep = np.ones(grid.shape[0], dtype=bool)
    else:
        imgs = []
        for S in syms:
            S = np.rint(np.asarray(S)).astype(np.int64)[:ndim, :ndim]
            g = (grid @ S.T) % npt
            imgs.append(g[:, 0] + (npt * g[:, 1] if ndim == 2 else 0))
        imgs.append(lin)
        imgs = np.sort(np.stack(imgs), axis=0)
        keep = imgs[0] == lin
        distinct = 1 + (np.diff(imgs, axis=0) != 0).sum(axis=0)
        w = distinct.astype(np.float64)
    idx = np.zeros((int(keep.sum()), 3), dtype=np.int32)
    idx[:, :ndim] = grid[keep]
    return idx, w[keep]


class DeviceRule:
    """FourierPTR / FourierMonkhorstPack on the device (src/fourier.jl:127-130, 210-214), possibly one
    rank's shard of it.  `nnodes` = local node count; `nnodes_total` = length(rule) of the reference."""

    def __init__(self, backend, series, ndim, npt, syms, rank=0, nranks=1, allreduce=None):
        ctx = backend.ctx
        ds = series.device(ctx)
        self.backend, self.series, self.ndim, self.npt = backend, series, ndim, int(npt)
        self.nsyms = 1 if syms is None else len(syms)
        self.syms = syms
        if ndim == 3:
            if syms is None:
                lo = (self.npt * rank) // nranks
                hi = (self.npt * (rank + 1)) // nranks
                self.dev = _lib.DeviceRule(ctx, ds, self.npt, k3_lo=lo, k3_hi=hi)
                self.nnodes_total = self.npt ** 3
            else:
                # symptr_rule + CSR compaction on the device; the dense weights never visit the host
                sy = embed_syms(syms, np.asarray(syms[0]).shape[0])
                # with an allreduce at hand every rank computes the orbit weights of ITS k3 planes only and the ranks add up
                # their node counts; without one (single rank, or a caller that shards by hand) the library counts all planes
                local_only = nranks > 1 and allreduce is not None
                self.dev = _lib.DeviceRule(ctx, ds, self.npt, syms=sy, k3_lo=rank, k3_stride=(-nranks if nranks > 1 else 1), count_all=not local_only)
                if local_only:
                    self.nnodes_total = int(round(float(np.asarray(allreduce(np.array([float(len(self.dev))]))).reshape(-1)[0])))
                else:
                    self.nnodes_total = self.dev.nirr_total
        else:
            idx, w = symptr_nodes_lowdim(self.npt, ndim, syms)
            self.nnodes_total = idx.shape[0]
            lo = (idx.shape[0] * rank) // nranks
            hi = (idx.shape[0] * (rank + 1)) // nranks
            self.dev = _lib.DeviceRule(ctx, ds, self.npt, nodes=idx[lo:hi], weights=w[lo:hi])
        self.nnodes = len(self.dev)

    def __len__(self):
        return self.nnodes_total

    def materialize(self):
        self.dev.materialize()

    def resolvent_sum(self, z, sigma, fkind):
        """sum_i w_i f(H(k_i)) over the local nodes, one value per frequency (unscaled)."""
        return self.dev.resolvent_sum(z, sigma=sigma, scale=1.0, fkind=fkind)

    def resolvent_matrix_sum(self, z, sigma):
        """sum_i w_i (z - H(k_i) - Sigma)^-1 over the local nodes -> [nw, n, n] (unscaled)"""
        return self.dev.resolvent_matrix_sum(z, sigma=sigma, scale=1.0)

    def eig_sum(self, kind, params):
        return self.dev.eig_sum(kind, params, scale=1.0)

    def eig_sum_batch(self, kind, params):
        """all parameter sets over ONE diagonalisation of every H(k) (parameter sweep over a shared grid, src/interfaces.jl:199-243)"""
        return self.dev.eig_sum_batch(kind, params, scale=1.0)

    def copy_out(self):
        return self.dev.copy_out()

    def ggr_data(self, ndim, copy=True):
        """get_ggr_data (src/dos_ggr.jl:14-44) on the local nodes: (energies, velocities), cached on the device"""
        return self.dev.ggr_data(ndim, copy=copy)

    def ggr_sum(self, E):
        """sum_ggr (src/dos_ggr.jl:58-65) over the local nodes for every energy in E"""
        return self.dev.ggr_sum(E)

    def close(self):
        self.dev.close()


def share_planes(npt, rank, k3_stride):
    """The k3 planes `abz_rule_create_sym / _symptr(..., k3_lo=rank, k3_stride)` select (`share_plane`, csrc/abz_common.cuh):
    k3_stride > 0: rank, rank + k3_stride, ... (the reference's round-robin dealing, src/fourier.jl:246-255);
    k3_stride < 0: serpentine dealing among W = -k3_stride ranks: rank, 2W-1-rank, 2W+rank, 4W-1-rank, ..."""
    if k3_stride > 0:
        return list(range(rank, npt, k3_stride))
    w2 = -2 * k3_stride
    return [p for p in range(npt) if p % w2 in (rank, w2 - 1 - rank)]


class DeviceBackend:
    """iai_engine: "native" (default) runs IAI's adaptive control flow in the library's C++ host engine
    (abz_iai_solve, one call per solve); "python" drives abz_nest_* round by round from iai.NestedGK.
    iai_device_leaves: run each innermost 1-D adaptive integral entirely on the device (norb <= 6);
    iai_device_middles: in 3-d solves also each middle integral (one CTA pair per node of the outermost panels; norb <= 5).
    iai_speculate: look-ahead on the outermost integral (two bisections per device round; same decisions and numevals)."""

    def __init__(self, device=None, ctx=None, iai_engine="native", iai_device_leaves=True, iai_device_middles=True, iai_speculate=True):
        self.ctx = ctx if ctx is not None else default_context(device)
        self.iai_engine, self.iai_device_leaves = iai_engine, iai_device_leaves
        self.iai_device_middles = iai_device_middles      # 3-d solves: whole middle integrals on the device too (abz_iai.cuh)
        self.iai_speculate = iai_speculate
        self._wsym_cache = {}

    def symptr_rule(self, npt, syms):
        """AutoSymPTR.symptr_rule on the device, cached per (npt, group) — the reference recomputes it
        for every rule and names it the likely bottleneck (src/fourier.jl:270)."""
        sy = embed_syms(syms, np.asarray(syms[0]).shape[0])
        key = (int(npt), sy.tobytes())
        if key not in self._wsym_cache:
            if len(self._wsym_cache) > 8:
                self._wsym_cache.pop(next(iter(self._wsym_cache)))
            self._wsym_cache[key] = self.ctx.symptr_rule(int(npt), sy)
        return self._wsym_cache[key]

    def make_rule(self, series, ndim, npt, syms, rank=0, nranks=1, allreduce=None):
        return DeviceRule(self, series, ndim, npt, syms, rank, nranks, allreduce)

    def make_nest(self, series, ndim, cap2, cap1):
        return _lib.DeviceNest(self.ctx, series.device(self.ctx), ndim, cap2 if ndim == 3 else 0, cap1 if ndim >= 2 else 0)

    @property
    def launch_count(self):
        return self.ctx.launch_count
