"""Write a seeded sample of BASELINE config 4 (the bench workload) for bench/reference_threads.jl.
usage: python scripts/dump_config4.py <n_qps> <out.bin>     (QPs evenly spaced over the 8192-QP global batch, like bench.py's CPU sample)"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ssqp_b200 as S

n, path = int(sys.argv[1]), sys.argv[2]
total = 8192
w = S.workloads.config4(index=np.linspace(0, total - 1, n).astype(np.int64), total=total)
N, M, J = w["V"].shape[0], w["A"].shape[0], w["G"].shape[0]
with open(path, "wb") as f:
    np.array([N, M, J, n], dtype="<i8").tofile(f)
    for name in ("V", "A", "G"):
        np.asfortranarray(w[name], dtype="<f8").T.tofile(f)          # column-major
    for name in ("q", "b", "g", "d", "u"):
        np.ascontiguousarray(w[name], dtype="<f8").tofile(f)         # (nb, len) row-major == (len, nb) column-major
print("wrote", path, "N M J nb =", N, M, J, n)
