"""Re-run the oracle on a subset of the bench shard and patch tests/golden/config4_shard0of8.npz in place.

Used once in round 2: the first full-shard comparison with the device exposed a bug in the oracle's `lstsq` (the purged-row
multiplier of KKTchk!, src/SSQP.jl:158 — column norms were not swapped with their columns in the pivoted QR).  lstsq only
runs on trips whose working set had a row purged by getRowsGJr, so only QPs that reach such a trip can change: this script
re-solves (both oracle forms) every QP named on the command line — the QPs whose trip count differed from the device's in
either form, plus the device's six `degen > 0` QPs — and rewrites their entries.
Run:  python tests/golden/patch_golden_shard.py idx.npy"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ssqp_b200 as S                      # noqa: E402
from oracle import ssqp_oracle as O        # noqa: E402
from make_golden_shard import TOTAL, SHARDS, N, XS_EVERY, projections   # noqa: E402


def main():
    pick = np.unique(np.load(sys.argv[1]).astype(np.int64))
    path = os.path.join(HERE, "config4_shard0of8.npz")
    g = dict(np.load(path))
    idx = g["index"][pick]
    c = S.workloads.config4(index=idx, total=TOTAL)
    P = projections()
    for form in ("lapack", "scalar"):
        assert O.use_lapack(form == "lapack") == form
        r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], nthreads=8, want_stats=True)
        x = r["x"]
        changed = np.flatnonzero(r["status"] != g["status_" + form][pick])
        print(form, "re-solved", pick.size, "QPs; trip count changed on", pick[changed].tolist(),
              "from", g["status_" + form][pick][changed].tolist(), "to", r["status"][changed].tolist())
        g["status_" + form][pick] = r["status"]
        g["S_" + form][pick] = r["S"]
        g["loops_" + form][pick] = r["stats"][:, 5]
        g["obj_" + form][pick] = 0.5 * np.einsum("ij,jk,ik->i", x, c["V"], x) + np.einsum("ij,ij->i", x, c["q"])
        g["xinf_" + form][pick] = np.abs(x).max(axis=1)
        g["proj_" + form][pick] = x @ P.T
        for t, i in enumerate(pick):
            if i % XS_EVERY == 0:
                g["xs_" + form][i // XS_EVERY] = x[t]
    np.savez_compressed(path, **g)
    print("patched", path)


if __name__ == "__main__":
    main()
