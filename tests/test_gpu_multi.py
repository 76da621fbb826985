"""Multi-GPU tests (need >= 2 visible CUDA devices; skipped otherwise): in-process sharding by QP index inside the
C ABI (one host thread + stream per device) must return exactly what one device returns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_device_context_matches_single_device():
    import ssqp_b200 as S
    if S.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    c = S.workloads.config4(index=np.arange(0, 65536, 2048), total=65536)        # 32 QPs
    one = S.Context([0])
    one.set_shared(c["V"], c["A"], c["G"])
    X1, S1, st1 = one.solve_batch(c["q"], c["b"], c["g"], c["d"], c["u"])
    two = S.Context([0, 1])
    two.set_shared(c["V"], c["A"], c["G"])
    X2, S2, st2 = two.solve_batch(c["q"], c["b"], c["g"], c["d"], c["u"])
    assert np.array_equal(st1, st2) and np.array_equal(S1, S2) and np.array_equal(X1, X2)
    stats = two.stats(32)
    assert (stats[:, 0] == st2).all()          # per-QP stats gathered back in QP order
    one.close(); two.close()
