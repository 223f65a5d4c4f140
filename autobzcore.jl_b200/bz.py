"""Brillouin-zone domain, symmetries and iterated limits — host-side mirror of the reference's
src/brillouin.jl:1-307 (SymmetricBZ, load_bz for FBZ / InversionSymIBZ / CubicSymIBZ) and of the
IteratedIntegration.jl limit types it uses (CubicLimits, TetrahedralLimits).  Pure control plane:
nothing here is on the data path."""
import itertools

import numpy as np


class CubicLimits:
    """IteratedIntegration.CubicLimits(a, b): x_d in [a_d, b_d]."""

    def __init__(self, a, b):
        self.a = tuple(float(x) for x in np.atleast_1d(a))
        self.b = tuple(float(x) for x in np.atleast_1d(b))
        if len(self.a) != len(self.b):
            raise ValueError("endpoints must have the same length")

    @property
    def ndim(self):
        return len(self.a)

    def segments(self):
        """limit_iterate(lims): segments of the outermost (last) variable."""
        return (self.a[-1], self.b[-1])

    def fix(self, x):
        """fixandeliminate(lims, x)"""
        return CubicLimits(self.a[:-1], self.b[:-1])

    def interior_point(self):
        return tuple((a + b) / 2 for a, b in zip(self.a, self.b))


class TetrahedralLimits:
    """IteratedIntegration.TetrahedralLimits(a): 0 <= x_d <= a_d s, then s <- x_d / a_d
    (used by load_bz(CubicSymIBZ), src/brillouin.jl:301-307)."""

    def __init__(self, a, s=1.0):
        self.a = tuple(float(x) for x in np.atleast_1d(a))
        self.s = float(s)

    @property
    def ndim(self):
        return len(self.a)

    def segments(self):
        return (0.0, self.a[-1] * self.s)

    def fix(self, x):
        return TetrahedralLimits(self.a[:-1], float(x) / self.a[-1])

    def interior_point(self):
        pt, lims = [], self
        while lims.ndim > 0:
            a, b = lims.segments()
            x = (a + b) / 2
            pt.append(x)
            lims = lims.fix(x)
        return tuple(reversed(pt))


class SegmentedLimits:
    """Iterated limits with several breakpoints per variable, independent of the outer variables: x_d runs over the segments
    [s_d[0], s_d[1]], [s_d[1], s_d[2]], ... (a PuncturedInterval per level, src/fourier.jl:395,500; IteratedIntegration.CubicLimits is
    the one-segment case).  Every 1-D integral starts from all of its segments, as QuadGK does."""

    def __init__(self, *breaks):
        self.breaks = tuple(tuple(float(x) for x in b) for b in breaks)
        for b in self.breaks:
            if len(b) < 2 or any(y < x for x, y in zip(b, b[1:])):
                raise ValueError("every level needs at least two ascending breakpoints")

    @property
    def ndim(self):
        return len(self.breaks)

    def segments(self):
        return self.breaks[-1]

    def fix(self, x):
        return SegmentedLimits(*self.breaks[:-1])

    def interior_point(self):
        return tuple((b[0] + b[-1]) / 2 for b in self.breaks)


class PolyhedronLimits:
    """Iterated limits of a convex polyhedron (3-d) given by its vertices - the shape of the reference's IBZ limits
    (`Polyhedron3` / `Polygon2` of ext/SymmetryReduceBZExt.jl:33-58: `segments` = the distinct z (y) coordinates of the vertices,
    `fixandeliminate` = the polygon of the slice at z, then the x-interval of the slice at y).  The geometry of the IBZ itself
    (SymmetryReduceBZ.jl) is out of scope; this class accepts its OUTPUT, a vertex list in lattice coordinates."""

    def __init__(self, vertices, tol=None):
        from scipy.spatial import ConvexHull
        v = np.asarray(vertices, dtype=float)
        if v.ndim != 2 or v.shape[1] != 3 or v.shape[0] < 4:
            raise ValueError("vertices must be [nvert >= 4, 3]")
        hull = ConvexHull(v)
        self.vert = v[hull.vertices]
        edges = set()
        for simplex in hull.simplices:
            for a in range(3):
                i, j = sorted((int(simplex[a]), int(simplex[(a + 1) % 3])))
                edges.add((i, j))
        self._edges = [(v[i], v[j]) for i, j in sorted(edges)]
        self.tol = np.sqrt(np.finfo(float).eps) if tol is None else tol

    ndim = 3

    @staticmethod
    def _unique_sorted(vals, tol):
        out = []
        for x in sorted(float(t) for t in vals):
            if not out or abs(x - out[-1]) > tol * max(1.0, abs(x)):
                out.append(x)
        return tuple(out)

    def segments(self):
        return self._unique_sorted(self.vert[:, 2], self.tol)

    def fix(self, z):
        pts = []
        for p, q in self._edges:
            dz = q[2] - p[2]
            if abs(dz) <= 1e-300:
                if abs(p[2] - z) <= self.tol:
                    pts += [p[:2], q[:2]]
                continue
            t = (z - p[2]) / dz
            if -1e-12 <= t <= 1 + 1e-12:
                pts.append(p[:2] + min(max(t, 0.0), 1.0) * (q[:2] - p[:2]))
        return PolygonLimits(np.array(pts) if pts else np.zeros((0, 2)), self.tol)

    def interior_point(self):
        return tuple(self.vert.mean(axis=0))


class PolygonLimits:
    """Iterated limits of a convex polygon (the z-slice of a PolyhedronLimits; `Polygon2` of the reference's extension)"""

    def __init__(self, pts, tol=None):
        self.tol = np.sqrt(np.finfo(float).eps) if tol is None else tol
        p = np.asarray(pts, dtype=float).reshape(-1, 2)
        if p.shape[0] >= 3:
            from scipy.spatial import ConvexHull, QhullError
            try:
                p = p[ConvexHull(p).vertices]
            except QhullError:          # degenerate slice (a segment or a point)
                pass
        self.pts = p

    ndim = 2

    def segments(self):
        if self.pts.shape[0] == 0:
            return (0.0, 0.0)
        s = PolyhedronLimits._unique_sorted(self.pts[:, 1], self.tol)
        return s if len(s) >= 2 else (s[0], s[0])

    def fix(self, y):
        n = self.pts.shape[0]
        xs = []
        for i in range(n):
            p, q = self.pts[i], self.pts[(i + 1) % n]
            dy = q[1] - p[1]
            if abs(dy) <= 1e-300:
                if abs(p[1] - y) <= self.tol:
                    xs += [p[0], q[0]]
                continue
            t = (y - p[1]) / dy
            if -1e-12 <= t <= 1 + 1e-12:
                xs.append(p[0] + min(max(t, 0.0), 1.0) * (q[0] - p[0]))
        if not xs:
            return CubicLimits([0.0], [0.0])
        return CubicLimits([min(xs)], [max(xs)])

    def interior_point(self):
        return tuple(self.pts.mean(axis=0))


class SymmetricBZ:
    """SymmetricBZ(A, B, lims, syms) (src/brillouin.jl:33-41).  A, B hold the real / reciprocal basis
    vectors in their columns; lims and syms are in the lattice basis (fractional coordinates)."""

    def __init__(self, A, B, lims, syms):
        self.A = np.array(A, dtype=float)
        self.B = np.array(B, dtype=float)
        self.lims = lims
        self.syms = None if syms is None else [np.array(S) for S in syms]

    @property
    def ndim(self):
        return self.A.shape[0]

    @property
    def nsyms(self):
        return 1 if self.syms is None else len(self.syms)

    @property
    def is_full(self):
        return self.syms is None

    def __repr__(self):
        return f"{self.ndim}-dimensional Brillouin zone with {'trivial' if self.is_full else self.nsyms} symmetries"


def nsyms(bz):
    return bz.nsyms


def canonical_reciprocal_basis(A):
    """B = A'^-1 2 pi (src/brillouin.jl:9)"""
    A = np.array(A, dtype=float)
    return np.linalg.solve(A.T, 2 * np.pi * np.eye(A.shape[0]))


class FBZ:
    def __init__(self, ndim=None):
        self.ndim = ndim


class IBZ:
    """IBZ(n) (src/brillouin.jl:205-247): the polyhedral irreducible BZ.  In the reference the polyhedron and the point group come
    from the SymmetryReduceBZ.jl extension (ext/SymmetryReduceBZExt.jl: `calc_ibz` -> convex hull -> `load_limits`) and `load_bz`
    errors without it.  That package is not available here, so the caller supplies what it would compute: the IBZ's vertices in
    lattice coordinates (`polyhedron`, [nv, 3]) and the point-group operators in lattice coordinates (`syms`); `load_bz` then
    builds the same SymmetricBZ(A, B, polyhedral limits, syms).  Without them `load_bz(IBZ(), ...)` raises like the reference."""

    def __init__(self, ndim=None, polyhedron=None, syms=None):
        self.ndim, self.polyhedron, self.syms = ndim, polyhedron, syms


class InversionSymIBZ:
    def __init__(self, ndim=None):
        self.ndim = ndim


class CubicSymIBZ:
    def __init__(self, ndim=None):
        self.ndim = ndim


# ---- symmetry representations (src/brillouin.jl:44-114)
class AbstractSymRep:
    pass


class UnknownRep(AbstractSymRep):
    """fallback for values without a user-defined representation: the IBZ value is returned as is"""


class TrivialRep(AbstractSymRep):
    """values that do not transform under the group (numbers): IBZ value x nsyms"""


class FunctionRep(AbstractSymRep):
    """a user-supplied map (bz, x) -> x on the full BZ, e.g. sum_S S x S^H (the reference's user-defined SymRep + symmetrize_)"""

    def __init__(self, fn):
        self.fn = fn


def SymRep(f):
    """SymRep(f) (src/brillouin.jl:72-85): the representation of the integral of `f`.  Scalar-valued integrands are handled by
    `symmetrize` itself (TrivialRep); a matrix-valued integrand carries a FunctionRep if it was given `symmetrize=`."""
    if isinstance(f, AbstractSymRep):
        return f
    inner = getattr(f, "f", f)
    fn = getattr(inner, "symmetrize", None)
    return FunctionRep(fn) if callable(fn) else UnknownRep()


def symmetrize(f, bz, x):
    """symmetrize(f, bz, x) (src/brillouin.jl:87-107): map a value computed on the symmetry-reduced domain to the full BZ"""
    if bz.syms is None:
        return x
    if np.ndim(x) == 0:
        return bz.nsyms * x                       # TrivialRepType = Union{Number, 0-dim array}
    rep = SymRep(f)
    if isinstance(rep, TrivialRep):
        return bz.nsyms * x
    if isinstance(rep, FunctionRep):
        return rep.fn(bz, x)
    return x


def sign_flip_matrices(d):
    return [np.diag(np.array(s, dtype=np.int64)) for s in itertools.product((1, -1), repeat=d)]


def permutation_matrices(d):
    out = []
    for p in itertools.permutations(range(d)):
        m = np.zeros((d, d), dtype=np.int64)
        for i in range(d):
            m[i, p[i]] = 1
        out.append(m)
    return out


def cube_automorphisms(d):
    """S*P for sign flips S and permutations P (src/brillouin.jl:286): 2^d d! matrices incl. identity."""
    return [S @ P for P in permutation_matrices(d) for S in sign_flip_matrices(d)]


def load_bz(bz, A=None, B=None, atol=None):
    """load_bz(bz::AbstractBZ, A, [B]) (src/brillouin.jl:179-307)."""
    if A is None:
        if bz.ndim is None:
            raise ValueError("BZ dimension must be integer")
        A = np.eye(bz.ndim)
    A = np.array(A, dtype=float)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    d = A.shape[0]
    if bz.ndim is not None and bz.ndim != d:
        raise ValueError("dimension mismatch between the BZ type and the lattice")
    B = canonical_reciprocal_basis(A) if B is None else np.array(B, dtype=float)
    if B.shape != A.shape:
        raise ValueError(f"Bravais lattices {A} and {B} must have the same shape")
    tol = np.sqrt(np.finfo(float).eps) if atol is None else atol
    if np.linalg.norm(A.T @ B - 2 * np.pi * np.eye(d)) >= tol:
        raise ValueError(f"Real and reciprocal Bravais lattice bases non-orthogonal to tolerance {tol}")
    if isinstance(bz, FBZ):
        return SymmetricBZ(A, B, CubicLimits(np.zeros(d), np.ones(d)), None)
    ortho = np.allclose(A.T @ A, np.diag(np.diag(A.T @ A)))
    if isinstance(bz, InversionSymIBZ):
        if not ortho:
            import warnings
            warnings.warn("Non-orthogonal lattice vectors detected with InversionSymIBZ. Unexpected behavior may occur")
        return SymmetricBZ(A, B, CubicLimits(np.zeros(d), np.full(d, 0.5)), sign_flip_matrices(d))
    if isinstance(bz, CubicSymIBZ):
        if not ortho:
            import warnings
            warnings.warn("Non-orthogonal lattice vectors detected with CubicSymIBZ. Unexpected behavior may occur")
        return SymmetricBZ(A, B, TetrahedralLimits(np.full(d, 0.5)), cube_automorphisms(d))
    if isinstance(bz, IBZ):
        if bz.polyhedron is None or bz.syms is None:
            raise NotImplementedError("SymmetryReduceBZ extension not loaded: pass IBZ(polyhedron=vertices, syms=point group)")   # src/brillouin.jl:234-241
        verts = np.asarray(bz.polyhedron, dtype=float)
        if d != 3 or verts.ndim != 2 or verts.shape[1] != 3 or verts.shape[0] < 4:
            raise ValueError("IBZ(polyhedron=...) takes the [nv >= 4, 3] vertices of a convex polyhedron in lattice coordinates")
        return SymmetricBZ(A, B, PolyhedronLimits(verts), [np.asarray(S) for S in bz.syms])
    raise TypeError("unsupported BZ type")
