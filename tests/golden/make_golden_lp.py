"""Golden vectors for BASELINE config 5 at FULL size (LPs N=1000, M=20, J=180 sharing A and G): the oracle's SimplexLP
(src/Simplex.jl:831-1034 restated) on the first NLP LPs of the batch, one LP per core — about 15 minutes per LP.

Stored per LP: status, S (int8), simplex loop / pivot / flip counts, x, objective — for the oracle's LAPACK form (the
default: OpenBLAS dgetrf/dgetri/dgemm/dgemv, what Julia calls) or `--form scalar`.
Run:  python tests/golden/make_golden_lp.py [--nlp 8] [--form lapack]"""
import argparse
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def one(args):
    i, form = args
    import ssqp_b200 as S
    from oracle import ssqp_oracle as O
    assert O.use_lapack(form == "lapack") == form
    k = S.workloads.config5(index=np.array([i]))
    t = time.time()
    r = O.simplex_lp(k["c"][0], k["A"], k["G"], k["b"][0], k["g"][0], k["d"][0], k["u"][0])
    print("LP %d (%s): status %d loops %d in %.0f s" % (i, form, r["status"], r["stats"][0], time.time() - t), flush=True)
    return i, r["status"], r["S"].astype(np.int8), r["stats"], r["x"], float(k["c"][0] @ r["x"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlp", type=int, default=8)
    ap.add_argument("--form", default="lapack")
    ap.add_argument("--procs", type=int, default=8)
    a = ap.parse_args()
    with ProcessPoolExecutor(max_workers=a.procs) as ex:
        res = sorted(ex.map(one, [(i, a.form) for i in range(a.nlp)]))
    np.savez_compressed(os.path.join(HERE, "config5_full_%s_%d.npz" % (a.form, a.nlp)),
                        index=np.array([r[0] for r in res]), status=np.array([r[1] for r in res]),
                        S=np.stack([r[2] for r in res]), stats=np.stack([r[3] for r in res]),
                        x=np.stack([r[4] for r in res]), obj=np.array([r[5] for r in res]))
    print("wrote config5_full_%s_%d.npz" % (a.form, a.nlp))


if __name__ == "__main__":
    main()
