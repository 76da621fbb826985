"""Phase-1 end points of the oracle's LAPACK form for a sample of the bench shard 0::8 (config 4): 64 QPs on which the LAPACK
form and the scalar form of the oracle disagree in the trip count (Phase-1 ties decided by OpenBLAS-level roundoff) plus 32
on which they agree.  The device, warm-started from these points, must reproduce the LAPACK form's trip counts exactly
(tests/test_gpu_shard.py).  The points are stored because another CPU's OpenBLAS kernels may break the ties differently.
Run after make_golden_shard.py:  python tests/golden/make_golden_phase1.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ssqp_b200 as S                      # noqa: E402
from oracle import ssqp_oracle as O        # noqa: E402

TOTAL, SHARDS = 65536, 8


def main():
    g = np.load(os.path.join(HERE, "config4_shard0of8.npz"))
    diff = np.flatnonzero(g["status_lapack"] != g["status_scalar"])
    same = np.flatnonzero(g["status_lapack"] == g["status_scalar"])
    rng = np.random.default_rng(7)
    pick = np.sort(np.concatenate([rng.choice(diff, min(64, diff.size), replace=False), rng.choice(same, 32, replace=False)]))
    idx = np.arange(0, TOTAL, SHARDS, dtype=np.int64)[pick]
    c = S.workloads.config4(index=idx, total=TOTAL)
    assert O.use_lapack(True) == "lapack"
    r1 = O.init_batch(c["A"], c["G"], c["b"], c["g"], c["d"], c["u"], nthreads=8)
    assert (r1["status"] == 1).all()
    assert np.array_equal(r1["stats"][:, 0].astype(np.int64), g["loops_lapack"][pick]), "Phase 1 is not the golden's"
    # sanity: the oracle's own Phase 2 from these points gives the golden's trip counts
    for t in range(0, pick.size, 8):
        r = O.solve_qp(c["V"], c["A"], c["G"], c["q"][t], c["b"][t], c["g"][t], c["d"][t], c["u"][t], S0=r1["S"][t].copy(), x0=r1["x"][t])
        assert r["status"] == g["status_lapack"][pick[t]], (pick[t], r["status"], g["status_lapack"][pick[t]])
    np.savez_compressed(os.path.join(HERE, "config4_shard0of8_phase1_lapack.npz"), pick=pick, x0=r1["x"], S0=r1["S"].astype(np.int8))
    print("wrote %d Phase-1 points (%d where the two oracle forms disagree)" % (pick.size, np.isin(pick, diff).sum()))


if __name__ == "__main__":
    main()
