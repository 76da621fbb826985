"""Prototype (numpy, sequential) of the in-place from-scratch build of the reduced-KKT inverse on PACKED lower-triangular
storage, exactly in the order the CUDA kernel uses (kinv_build_chol in csrc/ssqp_kernel.cuh):

    K = [V_FF AE'; AE 0]  (order n = nK + nW)  ->  H = K^-1 = [VQ TC; TC' -C]
    C = (AE V_FF^-1 AE')^-1, TC = V_FF^-1 AE' C, VQ = V_FF^-1 - TC AE V_FF^-1     (src/SSQP.jl:322-331)

Steps: L = chol(V_FF) in place; Li = L^-1 in place; Y' = AE Li' ; Cinv = Y' Y; L2 = chol(Cinv); L2i = L2^-1;
U = L2i Y'; P = U Li; VQ = Li'Li - P'P; TC' = L2i' P; -C = -L2i'L2i.   All on one packed array."""
import numpy as np


def tri(i):
    return i * (i + 1) // 2


def build(VFF, AE):
    nK, nW = VFF.shape[0], AE.shape[0]
    n = nK + nW
    H = np.zeros(tri(n))
    def g(i, j): return H[tri(i) + j]
    def s(i, j, v): H[tri(i) + j] = v
    # 1. fill V_FF
    for i in range(nK):
        for j in range(i + 1):
            s(i, j, VFF[i, j])
    # 2. right-looking Cholesky
    for j in range(nK):
        d = g(j, j)
        assert d > 0
        sd = np.sqrt(d)
        s(j, j, sd)
        col = np.zeros(nK)
        for i in range(j + 1, nK):
            col[i] = g(i, j) / sd
            s(i, j, col[i])
        for i in range(j + 1, nK):
            for k in range(j + 1, i + 1):
                s(i, k, g(i, k) - col[i] * col[k])
    # 3. Li = inv(L), row by row
    for i in range(nK):
        dii = g(i, i)
        old = np.array([g(i, k) for k in range(i)])
        for j in range(i):
            acc = 0.0
            for k in range(j, i):
                acc += old[k] * g(k, j)
            s(i, j, -acc / dii)
        s(i, i, 1.0 / dii)
    # 4. Y' rows: y_r = Li a_r   (rows nK + r, first nK entries), a_r staged
    for r in range(nW):
        a = AE[r].copy()
        for i in range(nK):
            acc = 0.0
            for k in range(i + 1):
                acc += g(i, k) * a[k]
            s(nK + r, i, acc)
    # 5. Cinv = Y' Y'^T  into the W-part (lower)
    for r in range(nW):
        for q in range(r + 1):
            acc = 0.0
            for k in range(nK):
                acc += g(nK + r, k) * g(nK + q, k)
            s(nK + r, nK + q, acc)
    # 6. Cholesky of Cinv + inverse, on the strided W-part
    def gw(r, q): return g(nK + r, nK + q)
    def sw(r, q, v): s(nK + r, nK + q, v)
    for j in range(nW):
        d = gw(j, j)
        assert d > 0
        sd = np.sqrt(d)
        sw(j, j, sd)
        col = np.zeros(nW)
        for i in range(j + 1, nW):
            col[i] = gw(i, j) / sd
            sw(i, j, col[i])
        for i in range(j + 1, nW):
            for k in range(j + 1, i + 1):
                sw(i, k, gw(i, k) - col[i] * col[k])
    for i in range(nW):
        dii = gw(i, i)
        old = np.array([gw(i, k) for k in range(i)])
        for j in range(i):
            acc = 0.0
            for k in range(j, i):
                acc += old[k] * gw(k, j)
            sw(i, j, -acc / dii)
        sw(i, i, 1.0 / dii)
    # 7. U = L2i Y'  in place, rows from last to first
    for r in range(nW - 1, -1, -1):
        for c in range(nK):
            acc = 0.0
            for q in range(r + 1):
                acc += gw(r, q) * g(nK + q, c)
            s(nK + r, c, acc)
    # 8. P = U Li  in place per row, ascending j:  P[r,j] = sum_{i>=j} U[r,i] Li[i,j]
    for r in range(nW):
        for j in range(nK):
            acc = 0.0
            for i in range(j, nK):
                acc += g(nK + r, i) * g(i, j)
            s(nK + r, j, acc)
    # 9. VQ = Li'Li - P'P  in place, rows ascending: (Li'Li)[i,j] = sum_{m>=i} Li[m,i] Li[m,j]  (j <= i)
    for i in range(nK):
        new = np.zeros(i + 1)
        for j in range(i + 1):
            acc = 0.0
            for m in range(i, nK):
                acc += g(m, i) * g(m, j)
            for r in range(nW):
                acc -= g(nK + r, i) * g(nK + r, j)
            new[j] = acc
        for j in range(i + 1):
            s(i, j, new[j])
    # 10. TC' = L2i' P  in place rows ascending: TC'[r,:] = sum_{q>=r} L2i[q,r] P[q,:]
    for r in range(nW):
        for c in range(nK):
            acc = 0.0
            for q in range(r, nW):
                acc += gw(q, r) * g(nK + q, c)
            s(nK + r, c, acc)
    # 11. -C = -L2i'L2i  in place rows ascending
    for i in range(nW):
        new = np.zeros(i + 1)
        for j in range(i + 1):
            acc = 0.0
            for m in range(i, nW):
                acc += gw(m, i) * gw(m, j)
            new[j] = -acc
        for j in range(i + 1):
            sw(i, j, new[j])
    full = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            full[i, j] = full[j, i] = g(i, j)
    return full


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    nK, nW = 13, 5
    B = rng.standard_normal((nK, nK)); V = B @ B.T + 0.5 * np.eye(nK)
    A = rng.standard_normal((nW, nK))
    Hm = build(V, A)
    Kmat = np.block([[V, A.T], [A, np.zeros((nW, nW))]])
    print("max |H K - I| =", np.abs(Hm @ Kmat - np.eye(nK + nW)).max())
    Hm0 = build(V, np.zeros((0, nK)))
    print("W = 0:", np.abs(Hm0 @ V - np.eye(nK)).max())
