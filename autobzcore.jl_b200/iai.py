"""IAI host engine: nested adaptive Gauss-Kronrod (7,15) with the reference's control flow, batched
level-synchronously for the device.

Restates do_solve(::FourierIntegrand, lims, ::NestedQuad) + init_nest (src/fourier.jl:432-510) on top of
QuadGK.jl's do_quadgk / adapt / refine / evalrule and DataStructures.jl's binary heap (the algorithms
behind IteratedIntegration.auxquadgk, call site src/algorithms.jl:236-237).  Every 1-D adaptive
integral is an independent state machine whose trajectory depends only on its own integrand values
and tolerance, so evaluating all live panels of all live integrals in one device batch per round
leaves every accept/refine decision — and hence EvalCounter's numevals (src/fourier.jl:516-523) —
unchanged with respect to the sequential recursion."""
import numpy as np

# QUADPACK qk15 abscissae (x <= 0 half, as QuadGK.kronrod(7) orders them), Kronrod and Gauss weights
GK_X = np.array([-0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
                 -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
                 -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
                 -0.207784955007898467600689403773245, 0.0])
GK_W = np.array([0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                 0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                 0.204432940075298892414161999234649, 0.209482141084727828012999174891714])
GK_GW = np.array([0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                  0.381830050505118944950369775488975, 0.417959183673469387755102040816327])

# offsets (1 + x) and (1 - x) in QuadGK.evalrule's evaluation order:
# (x2+,x2-),(x1+,x1-),(x4+,x4-),(x3+,x3-),(x6+,x6-),(x5+,x5-), centre, (x7+,x7-)
_ORDER = [1, 0, 3, 2, 5, 4]
_OFF = []
for _i in _ORDER:
    _OFF += [1.0 + GK_X[_i], 1.0 - GK_X[_i]]
_OFF += [1.0, 1.0 + GK_X[6], 1.0 - GK_X[6]]
_OFF = np.array(_OFF)
# note: a + (1 + x[8]) s with x[8] = 0 equals a + s exactly


def gk15_nodes(a, b):
    """nodes of evalrule on [a, b]; a, b scalars or arrays -> (..., 15)"""
    a = np.asarray(a, dtype=np.float64)
    s = 0.5 * (np.asarray(b, dtype=np.float64) - a)
    return a[..., None] + _OFF * s[..., None]


def gk15_combine(a, b, f, vdim=0):
    """QuadGK.evalrule: f (..., 15) values in gk15_nodes order -> (I, E), same operation order.
    vdim > 0: array-valued integrand, f is (..., 15, *value_shape) with vdim value axes; E = norm(Ik s - Ig s) (Frobenius, LinearAlgebra.norm)"""
    if vdim:
        f = np.moveaxis(np.asarray(f), -1 - vdim, -1)          # (..., *value_shape, 15)
    s = 0.5 * (np.asarray(b, dtype=np.float64) - np.asarray(a, dtype=np.float64))
    if vdim:
        s = s.reshape(s.shape + (1,) * vdim)
    fg = f[..., 0] + f[..., 1]
    fk = f[..., 2] + f[..., 3]
    Ig = fg * GK_GW[0]
    Ik = fg * GK_W[1] + fk * GK_W[0]
    fg = f[..., 4] + f[..., 5]
    fk = f[..., 6] + f[..., 7]
    Ig = Ig + fg * GK_GW[1]
    Ik = Ik + (fg * GK_W[3] + fk * GK_W[2])
    fg = f[..., 8] + f[..., 9]
    fk = f[..., 10] + f[..., 11]
    Ig = Ig + fg * GK_GW[2]
    Ik = Ik + (fg * GK_W[5] + fk * GK_W[4])
    f0 = f[..., 12]
    Ig = Ig + f0 * GK_GW[3]
    Ik = Ik + (f0 * GK_W[7] + (f[..., 13] + f[..., 14]) * GK_W[6])
    Iks = Ik * s
    Igs = Ig * s
    if vdim:
        D = Iks - Igs
        return Iks, np.sqrt(np.sum(D.real ** 2 + D.imag ** 2, axis=tuple(range(-vdim, 0))))     # LinearAlgebra.norm = sqrt(sum(abs2, D))
    return Iks, np.abs(Iks - Igs)


def kronrod(n):
    """Gauss-Kronrod rule of order n in QuadGK.kronrod's conventions: x (the n+1 nodes in [-1, 0], ascending, x[n] = 0),
    w (their Kronrod weights), gw (the Gauss weights of x[1], x[3], ...).  The Kronrod nodes are the Gauss-Legendre nodes plus the roots
    of the Stieltjes polynomial E_{n+1} (orthogonal to every polynomial of degree <= n under the signed weight P_n); E_{n+1} is found
    in the Legendre basis from those n+1 conditions, its roots are polished by Newton steps, the weights come from exactness on
    P_0 .. P_2n.  Reproduces QUADPACK's qk15 / qk21 / qk31 tables to 3e-16 (tests/test_host_logic.py); QuadGK.jl computes the same rule by
    Laurie's Jacobi-Kronrod matrix, so agreement with the reference for orders != 7 is at rounding level, not bitwise."""
    from numpy.polynomial import legendre as L
    if n < 2:
        raise ValueError("Gauss-Kronrod order must be at least 2")
    xq, wq = L.leggauss(2 * n + 3)
    eye = np.eye(2 * n + 2)
    P = np.stack([L.legval(xq, eye[j, : j + 1]) for j in range(n + 2)])
    G = np.einsum("q,q,jq,kq->kj", wq, P[n], P, P[: n + 1])
    idx = [j for j in range(n + 1) if (j - (n + 1)) % 2 == 0]                   # E_{n+1} has the parity of n + 1
    csub = np.linalg.lstsq(G[:, idx], -G[:, n + 1], rcond=None)[0]
    c = np.zeros(n + 2)
    c[n + 1] = 1.0
    c[idx] = csub
    xe = np.sort(L.legroots(c).real)
    dc = L.legder(c)
    for _ in range(4):
        xe = xe - L.legval(xe, c) / L.legval(xe, dc)
    xg, gwf = L.leggauss(n)
    xs = np.sort(np.concatenate([xe, xg]))
    xs = 0.5 * (xs - xs[::-1])
    xs[n] = 0.0
    V = np.stack([L.legval(xs, eye[k, : k + 1]) for k in range(2 * n + 1)])
    rhs = np.zeros(2 * n + 1)
    rhs[0] = 2.0
    w = np.linalg.solve(V, rhs)
    w = 0.5 * (w + w[::-1])
    return xs[: n + 1].copy(), w[: n + 1].copy(), gwf[: (n + 1) // 2].copy()


class GKRule:
    """QuadGK.evalrule for one Gauss-Kronrod order (AuxQuadGKJL(order=n), src/algorithms.jl:202-208): K = 2n + 1 nodes per panel in
    evalrule's evaluation order, and the (I, E) combination in its operation order.  Order 7 uses the QUADPACK table above (the
    constants QuadGK.jl caches for Float64), so it is bit-identical to gk15_nodes / gk15_combine."""
    _cache = {}

    def __new__(cls, n=7):
        n = int(n)
        if n not in cls._cache:
            r = super().__new__(cls)
            r.n, r.K = n, 2 * n + 1
            r.x, r.w, r.gw = (GK_X, GK_W, GK_GW) if n == 7 else kronrod(n)
            r.odd = (n % 2 == 1)                                   # n1 = 1 - (length(x) & 1)
            r.npairs = len(r.gw) - (1 if r.odd else 0)
            off = []
            for i in range(1, r.npairs + 1):                       # (x[2i]+, x[2i]-), (x[2i-1]+, x[2i-1]-), 1-based
                off += [1.0 + r.x[2 * i - 1], 1.0 - r.x[2 * i - 1], 1.0 + r.x[2 * i - 2], 1.0 - r.x[2 * i - 2]]
            off += [1.0]
            if r.odd:
                off += [1.0 + r.x[n - 1], 1.0 - r.x[n - 1]]
            r.off = np.array(off)
            assert r.off.size == r.K
            cls._cache[n] = r
        return cls._cache[n]

    def nodes(self, a, b):
        a = np.asarray(a, dtype=np.float64)
        s = 0.5 * (np.asarray(b, dtype=np.float64) - a)
        return a[..., None] + self.off * s[..., None]

    def combine(self, a, b, f, vdim=0):
        if self.n == 7:
            return gk15_combine(a, b, f, vdim)
        if vdim:
            f = np.moveaxis(np.asarray(f), -1 - vdim, -1)
        s = 0.5 * (np.asarray(b, dtype=np.float64) - np.asarray(a, dtype=np.float64))
        if vdim:
            s = s.reshape(s.shape + (1,) * vdim)
        w, gw = self.w, self.gw
        fg = f[..., 0] + f[..., 1]
        fk = f[..., 2] + f[..., 3]
        Ig = fg * gw[0]
        Ik = fg * w[1] + fk * w[0]
        p = 4
        for i in range(2, self.npairs + 1):
            fg = f[..., p] + f[..., p + 1]
            fk = f[..., p + 2] + f[..., p + 3]
            Ig = Ig + fg * gw[i - 1]
            Ik = Ik + (fg * w[2 * i - 1] + fk * w[2 * i - 2])
            p += 4
        if self.odd:
            f0 = f[..., p]
            Ig = Ig + f0 * gw[-1]
            Ik = Ik + (f0 * w[-1] + (f[..., p + 1] + f[..., p + 2]) * w[-2])
        else:
            Ik = Ik + f[..., p] * w[-1]
        Iks = Ik * s
        Igs = Ig * s
        if vdim:
            D = Iks - Igs
            return Iks, np.sqrt(np.sum(D.real ** 2 + D.imag ** 2, axis=tuple(range(-vdim, 0))))     # LinearAlgebra.norm = sqrt(sum(abs2, D))
        return Iks, np.abs(Iks - Igs)


# ---- DataStructures.jl binary heap with Base.Reverse on Segment.E (segments are (E, a, b, I)) ---------
def _lt_rev(x, y):
    return y[0] < x[0]


def _percolate_down(xs, i, x, n):
    while True:
        left = 2 * i
        if left > n:
            break
        r = left + 1
        j = left if (r > n or _lt_rev(xs[left - 1], xs[r - 1])) else r
        if not _lt_rev(xs[j - 1], x):
            break
        xs[i - 1] = xs[j - 1]
        i = j
    xs[i - 1] = x


def _percolate_up(xs, i, x):
    while True:
        j = i // 2
        if j < 1:
            break
        if not _lt_rev(x, xs[j - 1]):
            break
        xs[i - 1] = xs[j - 1]
        i = j
    xs[i - 1] = x


def heappop(xs):
    x = xs[0]
    y = xs.pop()
    if xs:
        _percolate_down(xs, 1, y, len(xs))
    return x


def heappush(xs, x):
    xs.append(x)
    _percolate_up(xs, len(xs), x)


class DomainError(FloatingPointError):
    """QuadGK throws DomainError when the integrand produces NaN/Inf."""


class _Pend:
    __slots__ = ("a", "b", "vals", "remaining", "tag", "shared")

    def __init__(self, a, b, tag, dtype, vshape=(), K=15):
        self.a, self.b, self.tag = a, b, tag
        self.vals = np.zeros((K,) + tuple(vshape), dtype=dtype)
        self.remaining = K
        self.shared = False


class _Integral:
    __slots__ = ("level", "lims", "atol", "slot", "parent", "heap", "I", "E", "numevals", "popped", "s1", "s2", "state", "outer", "init_remaining")

    def __init__(self, level, lims, atol, slot, parent, outer=()):
        self.level, self.lims, self.atol, self.slot, self.parent = level, lims, atol, slot, parent
        self.outer = outer          # coordinates already fixed by the outer integrals: (x_{level+2}, ..., x_ndim)
        self.heap, self.numevals, self.state = [], 0, 0
        self.s1 = self.s2 = self.popped = None


class NestedGK:
    """One nested adaptive integration of a Fourier integrand over iterated limits.

    nest: device arena (backend.make_nest) exposing contract3 / contract2 / eval / eval_h
    post(y) -> integrand values from the device's per-node output (e.g. DOS = -Im tr / pi)
    user(H, k) -> integrand values for a generic host integrand: H [n, n, npts] from the device (abz_nest_eval_h),
                  k [npts, ndim] the full points (FourierValue(limit_iterate(lims, state, x), H), src/fourier.jl:454)
    """

    def __init__(self, nest, ndim, lims, fkind, z, sigma, post, dtype, atol, rtol, maxevals, cap2=64, cap1=2048, user=None,
                 rank=0, nranks=1, allreduce=None, vshape=(), matrix=False, orders=None):
        # orders[level]: Gauss-Kronrod order of each level, level 0 = innermost variable (NestedQuad(algs...): alg = algs[dim],
        # src/algorithms.jl:462-463); None = GK(7,15) everywhere
        self.rules = [GKRule(7 if orders is None else orders[l]) for l in range(ndim)]
        # vshape: shape of an array-valued integrand (() = scalar); matrix=True: values are the device's (z - H - Sigma)^-1
        self.vshape, self.vdim, self.matrix = tuple(vshape), len(tuple(vshape)), matrix
        # multi-rank (ndim >= 2): the 15 nodes of every panel of the OUTERMOST integral are dealt round-robin to the ranks; when a
        # rank's own work is exhausted all ranks meet in one sum-allreduce of the outstanding outer panels' node values (zeros for
        # foreign nodes) and take the identical accept/refine decision - the scheme of the C++ engine (csrc/abz_iai_engine.hpp)
        self.rank, self.nranks, self.allreduce = int(rank), (int(nranks) if ndim >= 2 else 1), allreduce
        self.spawn_counter, self.shared_pends, self.exchanges = 0, [], 0
        self.nest, self.ndim, self.lims = nest, ndim, lims
        self.fkind, self.z, self.sigma, self.post, self.dtype = fkind, z, sigma, post, dtype
        self.user = user
        self.atol, self.rtol, self.maxevals = atol, rtol, maxevals
        self.free2 = list(range(cap2 - 1, -1, -1))
        self.free1 = list(range(cap1 - 1, -1, -1))
        self.q_eval, self.q_c3, self.q_c2 = [], [], []
        self.numevals = 0
        self.rounds = 0
        self.root_result = None

    # ---- slots
    def _alloc(self, level):
        free = self.free2 if level == 2 else self.free1
        if not free:
            raise MemoryError("IAI arena exhausted (too many live panels)")
        return free.pop()

    def _free(self, level, slot):
        (self.free2 if level == 2 else self.free1).append(slot)

    # ---- state machine
    def _start_segment(self, q, a, b, tag):
        rule = self.rules[q.level]
        pend = _Pend(a, b, tag, self.dtype, self.vshape, rule.K)
        if q.level == 0:
            self.q_eval.append((q, pend))
            return
        xs = rule.nodes(a, b)
        shared = self.nranks > 1 and q.level == self.ndim - 1
        if shared:
            pend.shared = True
            self.shared_pends.append((q, pend))
        for i in range(rule.K):
            if shared:
                mine = (self.spawn_counter % self.nranks) == self.rank
                self.spawn_counter += 1
                if not mine:
                    continue                                  # another rank owns this node
            x = float(xs[i])
            clims = q.lims.fix(x)
            csegs = tuple(clims.segments())
            length = csegs[-1] - csegs[0]                 # len = segs[end] - segs[1] (src/fourier.jl:476)
            slot = self._alloc(q.level)
            if q.level == 2:
                self.q_c3.append((x, slot))
            else:
                self.q_c2.append((x, q.slot, slot))
            catol = q.atol if not self.atol_given else (q.atol / length if length != 0.0 else float("inf"))
            child = _Integral(q.level - 1, clims, catol, slot, (q, pend, i), (x,) + q.outer)
            self._start_initial(child, csegs)

    def _start_initial(self, q, segs):
        """do_quadgk's first pass: evalrule on every initial segment (panel k carries tag -k)"""
        ns = len(segs) - 1
        q.init_remaining = ns
        q.heap = [None] * ns
        for k in range(ns):
            self._start_segment(q, segs[k], segs[k + 1], -k)

    def _finish(self, q):
        heap = q.heap
        Iv, Ev = heap[0][3], heap[0][0]
        for s in heap[1:]:
            Iv = Iv + s[3]
            Ev = Ev + s[0]
        if q.parent is None:
            self.root_result = (Iv, Ev)
            return
        pq, pend, i = q.parent
        self._free(pq.level, q.slot)
        pend.vals[i] = Iv
        if pend.shared:
            return                                            # combined in _exchange once every rank has delivered its nodes
        pend.remaining -= 1
        if pend.remaining == 0:
            Is, Es = self.rules[pq.level].combine(pend.a, pend.b, pend.vals, self.vdim)
            self._segment_done(pq, pend, Is[()] if not self.vdim else Is, float(Es))

    def _refine(self, q):
        s = heappop(q.heap)
        q.popped = s
        mid = (s[1] + s[2]) / 2
        q.s1 = q.s2 = None
        q.state = 1
        self._start_segment(q, s[1], mid, 1)
        self._start_segment(q, mid, s[2], 2)

    def _segment_done(self, q, pend, Is, Es):
        if not np.isfinite(Es):
            raise DomainError(f"integrand produced {Es} in the interval ({pend.a}, {pend.b})")
        seg = (Es, pend.a, pend.b, Is)
        if pend.tag <= 0:
            # I, E = left folds over the initial segments; heapify! only when a subdivision is needed (QuadGK do_quadgk)
            q.heap[-pend.tag] = seg
            q.init_remaining -= 1
            if q.init_remaining > 0:
                return
            q.I, q.E = q.heap[0][3], q.heap[0][0]
            for sg in q.heap[1:]:
                q.I = q.I + sg[3]
                q.E = q.E + sg[0]
            q.numevals = self.rules[q.level].K * len(q.heap)
            if q.numevals >= self.maxevals or q.E <= q.atol or q.E <= self.rtol * self._nrm(q.I):
                self._finish(q)
            else:
                for i in range(len(q.heap) // 2, 0, -1):
                    _percolate_down(q.heap, i, q.heap[i - 1], len(q.heap))
                self._refine(q)
            return
        if pend.tag == 1:
            q.s1 = seg
        else:
            q.s2 = seg
        if q.s1 is None or q.s2 is None:
            return
        s = q.popped
        q.I = (q.I - s[3]) + q.s1[3] + q.s2[3]
        q.E = (q.E - s[0]) + q.s1[0] + q.s2[0]
        q.numevals += 2 * self.rules[q.level].K
        heappush(q.heap, q.s1)
        heappush(q.heap, q.s2)
        if q.E > q.atol and q.E > self.rtol * self._nrm(q.I) and q.numevals < self.maxevals:
            self._refine(q)
        else:
            self._finish(q)

    def _nrm(self, v):
        return abs(v) if not self.vdim else float(np.sqrt(np.sum(np.real(v) ** 2 + np.imag(v) ** 2)))

    # ---- driver
    def run(self):
        self.atol_given = self.atol is not None
        atol = self.atol if self.atol_given else 0.0
        rtol = self.rtol
        if rtol is None:
            rtol = np.sqrt(np.finfo(float).eps) if atol == 0 else 0.0
        self.rtol = rtol
        root = _Integral(self.ndim - 1, self.lims, atol, None, None)
        self._start_initial(root, tuple(self.lims.segments()))
        while self.root_result is None:
            self.rounds += 1
            if self.q_c3:
                xs = np.array([t[0] for t in self.q_c3])
                sl = np.array([t[1] for t in self.q_c3], dtype=np.int64)
                self.q_c3 = []
                self.nest.contract3(xs, sl)
            if self.q_c2:
                xs = np.array([t[0] for t in self.q_c2])
                par = None if self.ndim == 2 else np.array([t[1] for t in self.q_c2], dtype=np.int64)
                sl = np.array([t[2] for t in self.q_c2], dtype=np.int64)
                self.q_c2 = []
                self.nest.contract2(xs, par, sl)
            batch = self.q_eval
            self.q_eval = []
            if not batch:
                if self.nranks > 1:
                    self._exchange()
                    continue
                raise RuntimeError("IAI engine stalled")
            nseg = len(batch)
            aa = np.array([p.a for _, p in batch])
            bb = np.array([p.b for _, p in batch])
            rule0 = self.rules[0]
            K0 = rule0.K
            xs = rule0.nodes(aa, bb)
            slots = None
            if self.ndim >= 2:
                slots = np.repeat(np.array([q.slot for q, _ in batch], dtype=np.int64), K0)
            if self.matrix:
                vals = self.nest.eval_matrix(xs.reshape(-1), slots, self.z, self.sigma).reshape((nseg, K0) + self.vshape)
            elif self.user is not None:
                H = self.nest.eval_h(xs.reshape(-1), slots)
                k = np.empty((K0 * nseg, self.ndim))
                k[:, 0] = xs.reshape(-1)
                if self.ndim > 1:
                    k[:, 1:] = np.repeat(np.array([q.outer for q, _ in batch], dtype=np.float64).reshape(nseg, self.ndim - 1), K0, axis=0)
                vals = np.asarray(self.user(H, k), dtype=self.dtype).reshape((nseg, K0) + self.vshape)
            else:
                y = self.nest.eval(xs.reshape(-1), slots, self.z, self.sigma, self.fkind)
                vals = self.post(y).reshape(nseg, K0)
            self.numevals += K0 * nseg
            Is, Es = rule0.combine(aa, bb, vals, self.vdim)
            for i in range(nseg):
                q, pend = batch[i]
                self._segment_done(q, pend, Is[i], float(Es[i]))
        if self.nranks > 1:          # evaluations of all ranks (EvalCounter semantics of the whole solve)
            self.numevals = int(round(float(np.asarray(self.allreduce(np.array([float(self.numevals)]))).reshape(-1)[0])))
        return self.root_result[0], self.root_result[1], self.numevals

    def _exchange(self):
        """all ranks: sum the outstanding outermost panels' node values, then every rank combines and decides identically"""
        sp, self.shared_pends = self.shared_pends, []
        if not sp:
            raise RuntimeError("IAI engine stalled")
        self.exchanges += 1
        buf = np.concatenate([np.asarray(p.vals, dtype=np.complex128).reshape(-1) for _, p in sp])
        top = self.rules[self.ndim - 1]
        buf = np.asarray(self.allreduce(buf), dtype=np.complex128).reshape((len(sp), top.K) + self.vshape)
        for (q, pend), vals in zip(sp, buf):
            pend.vals[...] = vals if np.iscomplexobj(pend.vals) else vals.real
            Is, Es = top.combine(pend.a, pend.b, pend.vals, self.vdim)
            self._segment_done(q, pend, Is[()] if not self.vdim else Is, float(Es))
