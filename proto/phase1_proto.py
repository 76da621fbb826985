"""numpy prototype of the device Phase-1 (initQP + cDantzigLP) with product-form invB updates,
unsorted basis rows and (value, variable-id) tie-breaks.  Finite d only (n=0 free variables)."""
import numpy as np
IN, DN, UP, OE, EO = 0, 1, 2, 3, 4


def init_qp(A, G, b, g, d, u, tol=2.0**-26, stats=None):
    M, N = A.shape; J = G.shape[0]; M0 = M + J; N0 = N + J; N1 = N0 + M0
    C = np.vstack([A, G]); bg = np.concatenate([b, g])
    cA = np.sqrt((C * C).sum(axis=0))
    # bounds for all N1 vars
    lo = np.concatenate([d, np.zeros(J + M0)]); hi = np.concatenate([u, np.full(J + M0, np.inf)])
    S1 = np.full(N1, DN, dtype=np.int32)
    Bv = np.arange(N0, N1)
    S1[Bv] = IN
    q0 = C @ d
    sig = np.where(bg >= q0, 1.0, -1.0)
    invB = np.diag(sig)
    qB = np.abs(q0 - bg)
    def column(k):
        if k < N: return C[:, k]
        e = np.zeros(M0)
        if k < N0: e[M + (k - N)] = 1.0
        else: e[k - N0] = sig[k - N0]
        return e
    loop = 0; Bland = False; pivots = 0
    while True:
        art = Bv >= N0
        pi = invB[art, :].sum(axis=0)            # invB' c_B
        dots = C.T @ pi                          # structural
        best = None
        def consider(k, h, ca):
            nonlocal best
            if h > tol:
                sc = h / ca
                if Bland:
                    if best is None: best = (sc, k)
                elif best is None or sc > best[0]:
                    best = (sc, k)
        for k in range(N):
            if S1[k] == IN: continue
            rc = -dots[k]
            consider(k, -rc if S1[k] == DN else rc, cA[k])
        for i in range(J):
            k = N + i
            if S1[k] == IN: continue
            rc = -pi[M + i]
            consider(k, -rc if S1[k] == DN else rc, 1.0)
        for i in range(M0):
            k = N0 + i
            if S1[k] == IN: continue
            rc = 1.0 - sig[i] * pi[i]
            consider(k, -rc if S1[k] == DN else rc, 1.0)
        if best is None: break
        loop += 1
        if loop > N1: Bland = True
        # NOTE: the reference sets Bland at the top of the loop body *before* selection for this loop.
        if Bland and loop == N1 + 1:
            # redo selection under Bland for this very loop (first candidate)
            best = None
            continue_sel = True
        k = best[1]
        p = invB @ column(k)
        kd = S1[k] == DN
        cand = []
        for j in range(M0):
            i = Bv[j]
            if kd:
                if p[j] > tol: cand.append(((qB[j] - lo[i]) / p[j], i, j, DN))
                elif p[j] < -tol: cand.append(((qB[j] - hi[i]) / p[j], i, j, UP))
            else:
                if p[j] > tol: cand.append(((qB[j] - hi[i]) / p[j], i, j, UP))
                elif p[j] < -tol: cand.append(((qB[j] - lo[i]) / p[j], i, j, DN))
        l = None
        fu = hi[k] < np.inf
        if kd:
            if not cand:
                if fu: l = -1
                else: raise RuntimeError("unbounded")
            else:
                gl, vi, row, Sl = min(cand, key=lambda e: (e[0], e[1]))
                if fu and gl >= hi[k] - lo[k]: l = -1
                elif (not fu) and np.isinf(gl): raise RuntimeError("unbounded")
                else: l = row
        else:
            if not cand: l = -2
            else:
                gl, vi, row, Sl = max(cand, key=lambda e: (e[0], -e[1]))
                if gl <= -(hi[k] - lo[k]): l = -2
                else: l = row
        if l == -1: S1[k] = UP
        elif l == -2: S1[k] = DN
        else:
            lv = Bv[l]
            Bv[l] = k
            S1[k] = IN; S1[lv] = Sl
            pr = p[l]
            rowl = invB[l, :] / pr
            invB = invB - np.outer(p, rowl)
            invB[l, :] = rowl
            pivots += 1
        # fresh q = invB (b - sum_nonbasic A1[:,k] x_k)
        r = bg.copy()
        for kk in range(N):
            if S1[kk] == IN: continue
            xv = lo[kk] if S1[kk] == DN else hi[kk]
            if xv != 0.0: r -= C[:, kk] * xv
        # slacks/artificials nonbasic sit at 0 (lower bound) -> no contribution
        qB = invB @ r
    x = np.where(S1 == UP, hi, lo)
    x[Bv] = qB
    if stats is not None: stats.update(loops=loop, pivots=pivots)
    S = S1[:N + J].copy()
    f = x[N0:].sum()
    if f > tol: return x[:N], S, 0
    for kk in range(N, N + J): S[kk] = OE if S[kk] == IN else EO
    return x[:N], S, 1
