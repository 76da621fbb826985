"""Host-side mirror of the reference's problem/API types (src/types.jl): Status, Settings, QP.

Same names, argument meaning and validation as the Julia structs so that parity tests read like the
reference's own tests; arrays are numpy float64.  (Julia source for the same surface: julia/SSQPB200.jl.)
"""
import enum
import warnings
import numpy as np
from . import capi


class Status(enum.IntEnum):
    """@enum Status (src/types.jl:17-23); Int32 codes shared with the C ABI."""
    IN = 0
    DN = 1
    UP = 2
    OE = 3
    EO = 4


IN, DN, UP, OE, EO = Status.IN, Status.DN, Status.UP, Status.OE, Status.EO


class Settings:
    """struct Settings{Float64} (src/types.jl:390-408): maxIter=7777, tol=2^-26, tolG=2^-33, pivot, rule."""

    def __init__(self, maxIter=7777, tol=2.0 ** -26, tolG=2.0 ** -33, pivot="column", rule="Dantzig"):
        self.maxIter = int(maxIter)
        self.tol = float(tol)
        self.tolG = float(tolG)
        self.pivot = pivot
        self.rule = rule

    def to_c(self):
        rules = {"Dantzig": 0, "stpEdgeLP": 1, "maxImprovement": 2}
        if self.rule not in rules:
            raise ValueError("unknown rule %r" % (self.rule,))
        return capi.CSettings(self.maxIter, self.tol, self.tolG, rules[self.rule], 0 if self.pivot == "column" else 1)


class QP:
    """struct QP{Float64} + keyword constructor (src/types.jl:214-301).

        min (1/2) z'Vz + q'z   s.t.  Az = b,  Gz <= g,  d <= z <= u

    Defaults as in the reference: q=0, u=+Inf, d=0, G=[], g=[], A=ones(1,N), b=[1].  V is symmetrised,
    u<d pairs are swapped, and `mc` carries the validity code (-70 not PSD, -30 d==u, -20 no bounds)."""

    def __init__(self, V, q=None, u=None, d=None, G=None, g=None, A=None, b=None, _raw=None):
        if _raw is not None:
            (self.V, self.A, self.G, self.q, self.b, self.g, self.d, self.u, self.N, self.M, self.J, self.mc) = _raw
            return
        V = np.array(V, dtype=np.float64)
        N = V.shape[0]
        if V.shape != (N, N):
            raise ValueError("incompatible dimension: V")                      # DimensionMismatch, types.jl:242
        V = (V + V.T) / 2                                                      # types.jl:243
        q = np.zeros(N) if q is None else np.array(q, dtype=np.float64).ravel()
        u = np.full(N, np.inf) if u is None else np.array(u, dtype=np.float64).ravel()
        d = np.zeros(N) if d is None else np.array(d, dtype=np.float64).ravel()
        G = np.ones((0, N)) if G is None else np.array(G, dtype=np.float64).reshape(-1, N)
        g = np.ones(0) if g is None else np.array(g, dtype=np.float64).ravel()
        A = np.ones((1, N)) if A is None else np.array(A, dtype=np.float64).reshape(-1, N)
        b = np.ones(1) if b is None else np.array(b, dtype=np.float64).ravel()
        M, J = b.size, g.size
        mc = 1
        if np.linalg.eigvalsh(V)[0] < 0:                                      # eigmin(V) < 0, types.jl:246
            mc = -70
            warnings.warn("variance matrix is not positive-semidefinite")
        if A.shape != (M, N):
            raise ValueError("incompatible dimension: A")
        if G.shape != (J, N):
            raise ValueError("incompatible dimension: G")
        for name, v in (("q", q), ("d", d), ("u", u)):
            if v.size != N:
                raise ValueError("incompatible dimension: " + name)
        if np.any(d == u):                                                     # types.jl:275
            mc = -30
            warnings.warn("downside bound == upper bound detected")
        if not (J > 0 or np.any(np.isfinite(d)) or np.any(np.isfinite(u))):    # types.jl:281
            mc = -20
            warnings.warn("no inequalities and bounds")
        iu = u < d
        if iu.any():                                                           # types.jl:286-292
            warnings.warn("swap the elements where u < d, to make sure u > d")
            t = u[iu].copy()
            u[iu] = d[iu]
            d[iu] = t
        self.V, self.A, self.G, self.q, self.b, self.g, self.d, self.u = V, A, G, q, b, g, d, u
        self.N, self.M, self.J, self.mc = N, M, J, mc

    @classmethod
    def with_L(cls, P, q, L=0.0):
        """QP(P::QP, q, L): replace the linear term by -L*q (src/types.jl:303-319); shares V/A/G/b/g/d/u."""
        return cls(None, _raw=(P.V, P.A, P.G, -L * np.asarray(q, dtype=np.float64), P.b, P.g, P.d, P.u,
                               P.N, P.M, P.J, P.mc))

    @classmethod
    def with_mu(cls, P, mu, q):
        """QP(P::QP, mu, q): append the row q'z = mu to Az=b and zero the linear term (src/types.jl:321-339)."""
        q = np.asarray(q, dtype=np.float64)
        return cls(None, _raw=(P.V, np.vstack([P.A, q[None, :]]), P.G, np.zeros(P.N), np.append(P.b, mu), P.g,
                               P.d, P.u, P.N, P.M + 1, P.J, P.mc))


class LP:
    """Mirror of `struct LP` + keyword constructor (src/types.jl:84-182):  min c'x  s.t. Ax=b, Gx<=g, d<=x<=u.
    mc: 1 ok, -30 some d == u, -20 no inequalities and no bounds; u<d pairs are swapped (with the reference's warning)."""

    def __init__(self, c, A, b, d=None, u=None, G=None, g=None):
        import warnings
        self.c = np.array(c, dtype=np.float64).ravel()
        N = self.N = self.c.size
        self.A = np.array(A, dtype=np.float64).reshape(-1, N) if np.size(A) else np.zeros((0, N))
        self.b = np.array(b, dtype=np.float64).ravel()
        self.G = np.ones((0, N)) if G is None else (np.array(G, dtype=np.float64).reshape(-1, N) if np.size(G) else np.zeros((0, N)))
        self.g = np.ones(0) if g is None else np.array(g, dtype=np.float64).ravel()
        self.d = np.zeros(N) if d is None else np.array(d, dtype=np.float64).ravel()
        self.u = np.full(N, np.inf) if u is None else np.array(u, dtype=np.float64).ravel()
        self.M, self.J = self.b.size, self.g.size
        if self.A.shape != (self.M, N):
            raise ValueError("incompatible dimension: A")
        if self.G.shape != (self.J, N):
            raise ValueError("incompatible dimension: G")
        if self.d.size != N or self.u.size != N:
            raise ValueError("incompatible dimension: d / u")
        self.mc = 1
        if np.any(self.d == self.u):
            self.mc = -30
            warnings.warn("downside bound == upper bound detected")
        if not (self.J > 0 or np.any(np.isfinite(self.d)) or np.any(np.isfinite(self.u))):
            self.mc = -20
            warnings.warn("no inequalities and bounds")
        iu = self.u < self.d
        if iu.any():
            warnings.warn("swap the elements where u < d, to make sure u > d")
            t = self.u[iu].copy(); self.u[iu] = self.d[iu]; self.d[iu] = t
