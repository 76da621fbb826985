"""The product's OWN multi-device path, timed: ONE process, one ssqp_create over all visible devices, one host-pointer
ssqp_solve_batch over the whole named batch (65 536 QPs of config 4) — what a Julia caller of the C ABI does.  The library
shards the batch by QP index over its devices (one host thread + stream per device, strided H2D gather / D2H scatter from
the caller's pinned buffers), no collective.  Compare with bench.py under torchrun (one process per GPU).
Run on a multi-GPU box:  python scripts/bench_multi_device.py [--devices 8] [--steps 3]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ssqp_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=65536)
a = ap.parse_args()
nd = a.devices or S.device_count()
c = S.workloads.config4(nb=a.batch)
ctx = S.Context(list(range(nd)))
ctx.set_shared(c["V"], c["A"], c["G"])
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
host = {k: pin(c[k]) for k in ("q", "b", "g", "d", "u")}
nb, N, J = a.batch, 500, 99
x_h = torch.empty((nb, N), dtype=torch.float64).pin_memory()
S_h = torch.empty((nb, N + J), dtype=torch.int32).pin_memory()
st_h = torch.empty((nb,), dtype=torch.int64).pin_memory()
import ctypes as C
vp = lambda t: C.c_void_p(t.data_ptr())
st = S.Settings().to_c()
def step():
    rc = ctx._L.ssqp_solve_batch(ctx._h, nb, None, vp(host["q"]), vp(host["b"]), vp(host["g"]), vp(host["d"]), vp(host["u"]),
                                 None, None, C.byref(st), None, vp(x_h), vp(S_h), vp(st_h))
    assert rc == 0, rc
step()
ts = []
for _ in range(a.steps):
    t = time.perf_counter(); step(); ts.append(time.perf_counter() - t)
sec = float(np.mean(ts))
print(json.dumps({"what": "one process, ssqp_create over %d devices, one ssqp_solve_batch (host pointers, pinned) over %d QPs" % (nd, nb),
                  "n_devices": nd, "qps_per_s_e2e": nb / sec, "sec_per_call": sec, "calls": a.steps, "kernel_ms_max_over_devices": ctx.last_kernel_ms(),
                  "qps_per_s_kernel": nb / (ctx.last_kernel_ms() * 1e-3), "optimal": int((st_h.numpy() > 0).sum()),
                  "h2d_bytes": int(nb * 8 * (3 * N + 1 + J)), "d2h_bytes": int(nb * (8 * N + 4 * (N + J) + 8))}))
ctx.close()
