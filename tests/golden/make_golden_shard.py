"""Golden vectors for the WHOLE bench shard 0::8 of BASELINE config 4 (8 192 of the 65 536 QPs), from the CPU oracle.

The reference (Julia) cannot run in this environment; these are outputs of oracle/ssqp_oracle.cpp in BOTH of its forms:
  "lapack": dense algebra on scipy's OpenBLAS through the routines Julia's LinearAlgebra calls (dpotrf/dpotri/dgetrf/dgetri/
            dgemm/dgemv) — the closest available stand-in for the reference's own roundoff;
  "scalar": the plain loops.
Where the two forms agree, the result does not depend on LAPACK-level roundoff; where they do not (a Phase-1 ratio-test
tie decided by 1e-17), the reference's own trip count depends on its (unpinned) OpenBLAS build — the file records both.

Stored per QP (inputs are regenerated from the seeds by workloads.config4):
  status_{form} int64, S_{form} int8 (N+J), loops_{form} (Phase-1 simplex loops), obj_{form}, xinf_{form},
  proj_{form} (4 fixed random projections of x), and the free components of x for every 8th QP (xs_{form}, full vectors).
Run (about 25 min per form on 8 cores):  python tests/golden/make_golden_shard.py [--threads 8] [--forms lapack,scalar]
"""
import argparse
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ssqp_b200 as S                      # noqa: E402
from oracle import ssqp_oracle as O        # noqa: E402

TOTAL, SHARDS, N, J = 65536, 8, 500, 99
XS_EVERY = 8


def projections():
    return np.random.default_rng(20261018).standard_normal((4, N))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--forms", default="lapack,scalar")
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--limit", type=int, default=0, help="only the first LIMIT QPs of the shard (smoke run)")
    args = ap.parse_args()
    idx = np.arange(0, TOTAL, SHARDS, dtype=np.int64)
    if args.limit:
        idx = idx[:args.limit]
    nb = idx.size
    P = projections()
    out = {"index": idx}
    for form in args.forms.split(","):
        assert O.use_lapack(form == "lapack") == form
        status = np.zeros(nb, np.int64); Sg = np.zeros((nb, N + J), np.int8); loops = np.zeros(nb, np.int64)
        obj = np.zeros(nb); xinf = np.zeros(nb); proj = np.zeros((nb, 4)); xs = np.zeros((len(idx[::XS_EVERY]), N))
        t0 = time.time()
        for s in range(0, nb, args.chunk):
            sl = slice(s, min(nb, s + args.chunk))
            c = S.workloads.config4(index=idx[sl], total=TOTAL)
            r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], nthreads=args.threads, want_stats=True)
            x = r["x"]
            status[sl] = r["status"]; Sg[sl] = r["S"]; loops[sl] = r["stats"][:, 5]
            obj[sl] = 0.5 * np.einsum("ij,jk,ik->i", x, c["V"], x) + np.einsum("ij,ij->i", x, c["q"])
            xinf[sl] = np.abs(x).max(axis=1); proj[sl] = x @ P.T
            for t in range(sl.start, sl.stop):
                if t % XS_EVERY == 0:
                    xs[t // XS_EVERY] = x[t - sl.start]
            print("%s %d/%d  %.0f s" % (form, sl.stop, nb, time.time() - t0), flush=True)
        for k, v in (("status", status), ("S", Sg), ("loops", loops), ("obj", obj), ("xinf", xinf), ("proj", proj), ("xs", xs)):
            out[k + "_" + form] = v
        np.savez_compressed(os.path.join(HERE, "config4_shard0of8.partial.npz"), **out)      # (kept if the second form is interrupted)
    name = "config4_shard0of8.npz" if not args.limit else "config4_shard0of8_first%d.npz" % args.limit
    np.savez_compressed(os.path.join(HERE, name), **out)
    if os.path.exists(os.path.join(HERE, "config4_shard0of8.partial.npz")):
        os.remove(os.path.join(HERE, "config4_shard0of8.partial.npz"))
    print("wrote", name)


if __name__ == "__main__":
    main()
