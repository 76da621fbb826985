"""One launch of the 256-thread kernel flavour on BASELINE config 2 (4 096 QPs, N=100, shared V) — the ncu target for that flavour."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssqp_b200 as S
c = S.workloads.config2(nb=4096)
for _ in range(2):
    X, St, st = S.solveQP_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"])
print("config2: kernel %.2f ms -> %.0f QPs/s, optimal %d | %s" % (S.context().last_kernel_ms(), 4096 / S.context().last_kernel_ms() * 1e3, (st > 0).sum(), S.context().last_launch_config()))
