"""numpy prototype of the device algorithm (Phase 2): explicit inverse of the reduced KKT matrix,
bordered rank-1 add / remove updates, fresh gradient each trip.  Used to validate decision parity
with the oracle before writing CUDA.  Not shipped / not imported by the product."""
import numpy as np
IN, DN, UP, OE, EO = 0, 1, 2, 3, 4


def isless(a, b):
    if a < b: return True
    if a == 0.0 and b == 0.0: return np.signbit(a) and not np.signbit(b)
    return False


class Sys:
    def __init__(s, V, C, N, M0):
        s.V, s.C, s.N, s.M0 = V, C, N, M0
        s.items = []
        s.Kinv = np.zeros((0, 0))
        s.min_abs_s = np.inf

    def col_of(s, it):
        N = s.N
        col = np.zeros(len(s.items))
        if it < N:
            for p, jt in enumerate(s.items):
                col[p] = s.V[jt, it] if jt < N else s.C[jt - N, it]
            return col, s.V[it, it]
        r = it - N
        for p, jt in enumerate(s.items):
            col[p] = s.C[r, jt] if jt < N else 0.0
        return col, 0.0

    def add(s, it):
        col, diag = s.col_of(it)
        n = len(s.items)
        if n == 0:
            sp = diag
            s.Kinv = np.array([[1.0 / sp]])
            s.items.append(it)
            return sp
        h = s.Kinv @ col
        sp = diag - col @ h
        s.min_abs_s = min(s.min_abs_s, abs(sp))
        Kn = np.zeros((n + 1, n + 1))
        Kn[:n, :n] = s.Kinv + np.outer(h, h) / sp
        Kn[:n, n] = -h / sp
        Kn[n, :n] = -h / sp
        Kn[n, n] = 1.0 / sp
        s.Kinv = Kn
        s.items.append(it)
        return sp

    def remove(s, it):
        j = s.items.index(it)
        c = s.Kinv[:, j].copy()
        piv = c[j]
        Kn = s.Kinv - np.outer(c, c) / piv
        n = len(s.items)
        last = n - 1
        if j != last:     # move last into j
            Kn[j, :] = Kn[last, :]
            Kn[:, j] = Kn[:, last]
            Kn[j, j] = Kn[last, last]
            s.items[j] = s.items[last]
        s.items.pop()
        s.Kinv = Kn[:last, :last].copy()
        return piv


def solve_phase2(V, A, G, q, b, g, d, u, S, x0, maxIter=7777, tol=2.0**-26, tolG=2.0**-33, log=None):
    N = V.shape[0]; M = A.shape[0]; J = G.shape[0]; M0 = M + J
    C = np.vstack([A, G]); bg = np.concatenate([b, g])
    S = S.copy(); z = x0.copy()
    sysm = None
    it = 0
    def rebuild():
        sm = Sys(V, C, N, M0)
        for k in range(N):
            if S[k] == IN: sm.add(k)
        for r in range(M): sm.add(N + r)
        for j in range(J):
            if S[N + j] == EO: sm.add(N + M + j)
        return sm
    while True:
        it += 1
        if it > maxIter: return z, S, -it
        F = np.flatnonzero(S[:N] == IN); K = len(F)
        if K == 0:
            p = V @ z + q
            S0 = S.copy(); t = True
            for k in range(N):
                if (p[k] >= -tol and S[k] == UP) or (p[k] <= tol and S[k] == DN):
                    S[k] = IN; t = False
            if t: return z, S, it
            ip = np.flatnonzero(S[:N] == IN)
            if len(ip) > 0 and np.max(np.abs(p[ip])) <= tol:
                S[ip] = S0[ip]; return z, S, it
            sysm = None
            continue
        if sysm is None: sysm = rebuild()
        items = sysm.items; n = len(items)
        gr = V @ z + q
        slack = bg - C @ z
        rhs = np.array([-gr[i] if i < N else slack[i - N] for i in items])
        sol = sysm.Kinv @ rhs
        pfull = np.zeros(N); lam = np.zeros(M0)
        for pp, i in enumerate(items):
            if i < N: pfull[i] = sol[pp]
            else: lam[i - N] = sol[pp]
        pn = np.max(np.abs(pfull[F]))
        if log is not None: log.append((K, n - K, pn))
        if pn > tolG:
            ev = []
            for k in F:
                t = pfull[k]
                if t > tol and u[k] < np.inf: ev.append(((u[k] - z[k]) / t, k, UP))
                elif t < -tol and d[k] > -np.inf: ev.append(((d[k] - z[k]) / t, k, DN))
            Cp = C @ pfull
            for j in range(J):
                if S[N + j] == OE and Cp[M + j] > tol:
                    ev.append((slack[M + j] / Cp[M + j], N + j, EO))
            L1 = 1.0
            if ev:
                best = ev[0]
                for e in ev[1:]:
                    if isless(e[0], best[0]): best = e
                L1 = best[0]
            if L1 < 1.0:
                z[F] += L1 * pfull[F]
                for (L, idx, To) in ev:
                    if L - L1 > tol: continue
                    S[idx] = To
                    if idx < N:
                        z[idx] = d[idx] if To == DN else u[idx]
                        sysm.remove(idx)
                    else:
                        sysm.add(N + M + (idx - N))
                if (S[:N] == IN).sum() == 0: sysm = None
                continue
            else:
                z[F] += pfull[F]
                gr = V @ z + q
        gam = gr + C.T @ lam
        best = None
        for k in range(N):
            if S[k] == UP and gam[k] > tolG: e = (-gam[k], k, IN)
            elif S[k] == DN and gam[k] < -tolG: e = (gam[k], k, IN)
            else: continue
            if best is None or isless(e[0], best[0]): best = e
        for j in range(J):
            if S[N + j] == EO and lam[M + j] < -tolG:
                e = (lam[M + j], N + j, OE)
                if best is None or isless(e[0], best[0]): best = e
        if best is not None:
            _, idx, To = best
            S[idx] = To
            if idx < N: sysm.add(idx)
            else: sysm.remove(N + M + (idx - N))
            continue
        for k in range(N):
            if S[k] == DN: z[k] = d[k]
            elif S[k] == UP: z[k] = u[k]
            else:
                if abs(z[k] - d[k]) < tol: z[k] = d[k]; S[k] = DN
                elif abs(z[k] - u[k]) < tol: z[k] = u[k]; S[k] = UP
        for j in range(J):
            S[N + j] = EO if abs(g[j] - z @ G[j]) < tol else OE
        return z, S, it, sysm.min_abs_s
