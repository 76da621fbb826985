"""Generate the golden vectors under tests/golden/ from the CPU oracle (oracle/ssqp_oracle.cpp).

The reference (Julia) cannot run in this environment, so the goldens are OUTPUTS OF THE ORACLE on the seeded
workloads of statusswitchingqp.jl_b200/workloads.py (inputs are regenerated from the seeds, only outputs are
stored).  They pin (a) the oracle against accidental edits and (b) the CUDA path at fixed, committed values.
Run:  python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ssqp_b200 as S                      # noqa: E402
from oracle import ssqp_oracle as O        # noqa: E402

W = S.workloads


def cases():
    """name -> workload dict (small enough for the oracle to finish in seconds)."""
    yield "kat_3asset", W.kat_3asset()
    yield "config1_n300", W.config1()
    yield "config2_shared_16", W.config2(nb=16)
    yield "config2_perqp_6", W.config2(nb=6, shared_V=False)
    yield "config3_sweep_4", W.config3(nb=4)            # (mu from the minimum-variance return to 0.98 x the maximum one)
    yield "config4_sample_6", W.config4(index=np.linspace(0, 65535, 6).astype(int), total=65536)
    yield "config4_degenerate_qp280_of_296", W.config4(index=np.array([279, 280, 281]), total=296)


def mu_range():
    """The two constants of workloads.config3's target-return range: the return of the minimum-variance portfolio
    (solveQP without the return row, q = 0) and the largest feasible return (SimplexLP max E'x), both from the oracle."""
    N = 500
    c = W.config3(nb=2)
    r = O.solve_qp(c["V"], np.ones((1, N)), c["G"], np.zeros(N), np.ones(1), c["g"][0], np.zeros(N), np.full(N, 0.05))
    lp = O.simplex_lp(-c["E"], np.ones((1, N)), c["G"], np.ones(1), c["g"][0], np.zeros(N), np.full(N, 0.05))
    assert r["status"] > 0 and lp["status"] in (1, 2)
    print("CONFIG3_MU_MINVAR =", repr(float(c["E"] @ r["x"])))
    print("CONFIG3_MU_MAX =", repr(float(c["E"] @ lp["x"])))


def main():
    if "--mu-range" in sys.argv:
        return mu_range()
    for name, c in cases():
        r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=r["x"], S=r["S"].astype(np.int8), status=r["status"])
        print(name, "status", r["status"].tolist())


if __name__ == "__main__":
    main()
