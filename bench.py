#!/usr/bin/env python
"""bench.py — QPs solved/sec (FP64) for the batched status-switching QP hot path on B200.

Workload (BASELINE.json configs[3]): the named batch of 65,536 portfolio QPs N=500, M=1, J=99 sharing V/A/G, per-QP
q and g.  EVERY N solves that whole batch (`--batch`, default 65536): rank r solves the interleaved shard r::world of
it ("strong" scaling, no data-path collective), so N=1 is the named configuration on one GPU.

One step = one pass of the hot path over the rank's shard.
  value : whole-job QPs/s with the shard already resident in HBM (ssqp_solve_batch_device), CUDA-event
          timed on the launch stream, max over ranks; L2 is flushed between timed steps.
  e2e   : the same through the host-pointer C-ABI call ssqp_solve_batch with PINNED host buffers
          (H2D of q,b,g,d,u + solve + D2H of x,S,status inside the timed region).
  --impl reference : the CPU oracle (reference-form restatement of solveQP; Julia is not installed, so the
          reference itself cannot run — kind "port") in its LAPACK form (the dense algebra on OpenBLAS through the
          routines Julia's LinearAlgebra calls, one QP per thread, BLAS threads = 1) on all host threads, on a
          bounded 256-QP sample per step; `--cpu-form scalar` times the plain-loop form instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_, M_, J_ = 500, 1, 99
METRIC = "QPs solved/sec (FP64, N=500 batch)"
UNIT = "QPs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="QPs per step over ALL GPUs (the named batch)")
    ap.add_argument("--cpu-sample", type=int, default=256, help="QPs in the CPU baseline sample")
    ap.add_argument("--cpu-form", default="lapack", choices=["lapack", "scalar"], help="dense algebra of the CPU oracle")
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed end-to-end steps (at most --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": "configs[3]: portfolio QPs N=500 M=1 J=99 (M+J=100), shared V/A/G, per-QP q,g; "
                        "d=0,u=0.05; cold start (Phase-1 simplex + Phase-2 active set)",
            "N": N_, "M": M_, "J": J_, "qps_per_gpu": args.batch // world, "global_batch": args.batch,
            "named_batch": 65536,
            "sharding": "rank r solves the interleaved shard r::%d of the %d-QP batch (strong scaling: every N solves the "
                        "whole batch), no collective" % (world, args.batch),
            "l2": "flushed between timed steps (256 MiB write)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_ev = threading.Event()
        self.rows = []

    def run(self):
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([t.strip() for t in line.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}



def shard_indices(rank, world, total):
    """Interleaved sharding by QP index: rank r solves QPs r, r+world, ... (trip counts are heavy-tailed and vary
    smoothly with the QP index, so interleaving balances the ranks).  No data-path collective."""
    return np.arange(rank, total, world, dtype=np.int64)


def reduce_over_ranks(my_ms, my_e2e_ms, my_sums, world, device):
    """max over ranks of the two timings, sum over ranks of the counters (the only collectives of the job)."""
    import torch
    import torch.distributed as dist
    red = torch.tensor([my_ms, my_e2e_ms], dtype=torch.float64, device=device)
    sums = torch.tensor(list(my_sums), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return float(red[0]), float(red[1]), [float(v) for v in sums]


def dram_traffic(nb):
    """DRAM bytes per launch, extrapolated per QP from the committed ncu --set full capture (profiles/*_ncu_full.csv)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_solve_kernel_ncu_full.csv")))
    if not files:
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, grid_qps = 0.0, None
    for line in open(files[-1]):
        t = line.strip().split(",")
        if len(t) == 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(t[2]) * mult.get(t[1], 1.0)
        if line.startswith("#") and "--batch" in line:
            try:
                grid_qps = int(line.split("--batch")[1].split()[0])
            except Exception:
                pass
    if not tot or not grid_qps:
        return None
    return tot / grid_qps * nb


def ncu_profile_numbers():
    """Executed FP64 flops per QP from the committed ncu capture (profiles/*_solve_kernel_ncu_full.csv): the
    smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum counters, 2 flops per DFMA."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_solve_kernel_ncu_full.csv")))
    if not files:
        return {}
    vals, grid_qps = {}, None
    for line in open(files[-1]):
        t = line.strip().split(",")
        if len(t) == 3 and t[0].startswith("smsp__sass_thread_inst_executed_op_d"):
            try:
                vals[t[0]] = float(t[2])
            except ValueError:
                pass
        if line.startswith("#") and "--batch" in line:
            try:
                grid_qps = int(line.split("--batch")[1].split()[0])
            except Exception:
                pass
    if not grid_qps or not vals:
        return {}
    fl = 0.0
    for k, v in vals.items():
        fl += (2.0 if "dfma" in k else 1.0) * v
    return {"executed_gflop_per_qp": fl / grid_qps / 1e9, "source": os.path.basename(files[-1])}


def cpu_sample_indices(total, n):
    return np.unique(np.linspace(0, total - 1, n).astype(np.int64))


def run_cpu(total, n_sample, steps, warmup, form="lapack"):
    """Time the CPU oracle (OpenMP over the batch, one QP per thread) on an evenly spaced sample."""
    from oracle import ssqp_oracle as O
    import ssqp_b200 as S
    form = O.use_lapack(form == "lapack")
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to its ranks: ask the OS, not OpenMP)
    try:
        thr = len(os.sched_getaffinity(0))
    except AttributeError:
        thr = os.cpu_count() or 1
    n = n_sample or 2 * thr
    idx = cpu_sample_indices(total, n)
    c = S.workloads.config4(index=idx, total=total)
    times = []
    for it in range(warmup + steps):
        t = time.perf_counter()
        r = O.solve_batch(c["V"], c["A"], c["G"], c["q"], c["b"], c["g"], c["d"], c["u"], nthreads=thr, want_stats=True)
        dt = time.perf_counter() - t
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    fref = float(r["stats"][:, 2].sum())         # flops of the reference's own operation count (explicit inverses, inv(lu) per pivot)
    return {"value": len(idx) / sec, "unit": UNIT, "cores": int(r["threads"]), "kind": "port", "form": "port-" + form,
            "gflops_per_core": fref / sec / 1e9 / max(int(r["threads"]), 1),
            "sample": "%d QPs evenly spaced over the named %d-QP batch, %.1f s per pass, %d timed passes; C++ restatement of "
                      "the reference in reference form (Julia unavailable), dense algebra: %s; %.1f GFLOP/s per core of the "
                      "reference's own operation count"
                      % (len(idx), total, sec, len(times),
                         "OpenBLAS dpotrf/dpotri/dgetrf/dgetri/dgemm/dgemv (scipy's build, 1 BLAS thread per QP)" if form == "lapack" else "plain loops",
                         fref / sec / 1e9 / max(int(r["threads"]), 1)),
            "sec_per_pass": sec, "n": int(len(idx)), "ok": int((r["status"] > 0).sum())}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    total = args.batch                          # QPs solved per step by the whole job: the named batch at every N
    # Strong scaling over the NAMED batch (BASELINE config 4: 65 536 QPs): rank r solves the interleaved shard r::world
    # of it, so N = 1 solves all 65 536 QPs on one GPU and 8 GPUs solve 8 192 each.
    named_total = total
    nshards = world

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = run_cpu(named_total, args.cpu_sample, args.steps, 1 if args.warmup else 0, args.cpu_form)      # (a CPU pass needs one warm-up at most)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sec_per_pass"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world),
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "form", "gflops_per_core", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import ssqp_b200 as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: this rank's shard of the named batch --------------------------------------
    idx = shard_indices(rank, nshards, named_total)
    c = S.workloads.config4(index=idx, total=named_total)
    nb = len(idx)
    ctx = S.Context([local_rank])
    ctx.set_shared(c["V"], c["A"], c["G"])
    st = S.Settings().to_c()

    host = {k: torch.from_numpy(np.ascontiguousarray(c[k])).pin_memory() for k in ("q", "b", "g", "d", "u")}
    devt = {k: v.to(dev) for k, v in host.items()}
    x_d = torch.empty((nb, N_), dtype=torch.float64, device=dev)
    S_d = torch.empty((nb, N_ + J_), dtype=torch.int32, device=dev)
    st_d = torch.empty((nb,), dtype=torch.int64, device=dev)
    x_h = torch.empty((nb, N_), dtype=torch.float64).pin_memory()
    S_h = torch.empty((nb, N_ + J_), dtype=torch.int32).pin_memory()
    st_h = torch.empty((nb,), dtype=torch.int64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)      # a real (non-NULL) stream: kernels and timing events share it
    torch.cuda.set_stream(stream)

    def step_device():
        ctx.solve_batch_device(nb, devt["q"].data_ptr(), devt["b"].data_ptr(), devt["g"].data_ptr(),
                               devt["d"].data_ptr(), devt["u"].data_ptr(), x_d.data_ptr(), S_d.data_ptr(),
                               st_d.data_ptr(), settings=st, stream=stream.cuda_stream)

    def step_host():
        L = ctx._L
        import ctypes as C
        vp = lambda t: C.c_void_p(t.data_ptr())
        rc = L.ssqp_solve_batch(ctx._h, nb, None, vp(host["q"]), vp(host["b"]), vp(host["g"]), vp(host["d"]),
                                vp(host["u"]), None, None, C.byref(st), None, vp(x_h), vp(S_h), vp(st_h))
        if rc != 0:
            raise RuntimeError("ssqp_solve_batch failed: %d" % rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()

    # ---- timed: HBM-resident value -----------------------------------------------------------------------
    # nvidia-smi numbers the physical GPUs; CUDA_VISIBLE_DEVICES may renumber what torch sees: address the GPU by its UUID
    try:
        smi_id = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        smi_id = vis.split(",")[local_rank].strip() if vis and len(vis.split(",")) > local_rank else str(local_rank)
    sampler = ClockSampler(smi_id)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        flush.fill_(1)                      # L2 flush between timed steps (not timed)
        e0.record(stream)
        step_device()
        e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    my_ms = float(np.mean(step_ms))
    if os.environ.get("SSQP_BENCH_RANKS"):      # diagnostics: every rank's own device time
        sys.stderr.write("rank %d (cuda:%d %s): %.1f ms per step %s\n" % (rank, local_rank, torch.cuda.get_device_name(local_rank), my_ms, ["%.0f" % t for t in step_ms]))
    kstats = ctx.stats(nb, device=True)
    status_dev = st_d.cpu().numpy()

    # ---- timed: end-to-end through the host-pointer C ABI ---------------------------------------------
    e2e_ms = None
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if not args.no_e2e:
        step_host()                          # warm the staging buffers
        barrier()
        ts = []
        for _ in range(e2e_steps):
            t = time.perf_counter()
            step_host()
            ts.append(time.perf_counter() - t)
        barrier()
        e2e_ms = float(np.mean(ts)) * 1e3
        if not np.array_equal(st_h.numpy(), status_dev):
            raise RuntimeError("host-path and device-path statuses differ")
    if rank == 0:
        sampler.stop_ev.set()
        sampler.join(timeout=5)

    # ---- max over ranks -----------------------------------------------------------------------------------
    ms, e2e_max, (n_ok, falg, bytes_streamed, trips, lploops) = reduce_over_ranks(
        my_ms, e2e_ms or 0.0,
        [float((status_dev > 0).sum()), float(kstats[:, 1].sum()), float(kstats[:, 10].sum()),
         float(kstats[:, 0].sum()), float(kstats[:, 4].sum())], world, dev)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
        fp64_peak = ctx.measure_fp64_peak()
        l2_peak = ctx.measure_read_bw(64, 20)
        # per-rank (single kernel launch per step) figures use this rank's own shard
        alg_bytes = nb * (8.0 * (3 * N_ + M_ + J_) + 8.0 * N_ + 4.0 * (N_ + J_) + 8.0)
        sec = my_ms * 1e-3
        falg_tf = float(kstats[:, 1].sum()) / sec / 1e12
        prof = ncu_profile_numbers()
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "solved_ok": int(n_ok), "trips_per_qp": trips / total, "lp_loops_per_qp": lploops / total,
            "gpu_launches": int(launches),
            # The binding resource of this path is not HBM (19.2 KB of algorithmic HBM bytes per QP, see roofline_hbm) and
            # not the tensor cores (tcgen05 has no FP64 kind; every dense operation is a GEMV or a rank-1 update): the
            # FP64 pipe fraction is reported in the contract key, with the algorithmic flops of SURVEY 8d (which credit
            # the from-scratch factorisation the rank-1 updates avoid) next to the flops the kernel actually executes.
            "roofline": {"bound": "fp64", "achieved": falg_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": falg_tf / fp64_peak if fp64_peak > 0 else None,
                         "traffic": dram_traffic(nb),
                         "kernel": "ssqp::ssqp_solve_kernel<512> (one launch per step, rank 0's shard)",
                         "what": "achieved = F_alg (SURVEY 8d: per trip K^3/3 + K^2 W + K W^2 + W^3/3 + 2K^2 + 4KW + 2W^2 + 2N^2 + "
                                 "2(N-K)W + 2 J_O (N+K), from each QP's own K_t, W_t, counted in-kernel) / kernel time: speed-up "
                                 "accounting; executed_* = the DFMA/DADD/DMUL the kernel issues (ncu, profiles/): pipe utilisation",
                         "peak_source": "DFMA microbenchmark in this run (no driver-measured FP64 peak exists)",
                         "algorithmic_gflop_per_qp": float(kstats[:, 1].sum()) / nb / 1e9,
                         "executed_gflop_per_qp": prof.get("executed_gflop_per_qp"),
                         "executed_tflops": (prof["executed_gflop_per_qp"] * nb / sec / 1e3) if prof.get("executed_gflop_per_qp") else None,
                         "executed_frac": (prof["executed_gflop_per_qp"] * nb / sec / 1e3 / fp64_peak) if prof.get("executed_gflop_per_qp") and fp64_peak > 0 else None,
                         "executed_source": prof.get("source"),
                         "traffic_note": "ncu dram__bytes_read+write per QP of the committed --set full capture x QPs per launch"},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / sec / 1e9 / hbm_peak, "peak_source": hbm_src,
                             "algorithmic_bytes_per_qp": alg_bytes / nb,
                             "note": "algorithmic HBM bytes are 19.2 KB/QP (DESIGN.md section 5): legitimately tiny"},
            "roofline_l2": {"achieved": float(kstats[:, 10].sum()) / sec / 1e9, "peak": l2_peak, "unit": "GB/s",
                            "frac": float(kstats[:, 10].sum()) / sec / 1e9 / l2_peak if l2_peak > 0 else None,
                            "what": "bytes streamed by the kernel's V / [A;G] / packed-inverse passes (counted in-kernel) / kernel time; "
                                    "peak = L2-resident read microbenchmark in this run"},
            "launch_config": ctx.last_launch_config(),
            "clocks": sampler.summary(),
            "wall_s_timed_region": t_wall,
        }
        if e2e_ms is not None:
            h2d = nb * 8 * (3 * N_ + M_ + J_)
            d2h = nb * (8 * N_ + 4 * (N_ + J_) + 8)
            line["e2e"] = {"value": total / (e2e_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_max, "steps": e2e_steps,
                           "api": "ssqp_solve_batch (host pointers, pinned), per rank"}
        if world == 1 and not args.no_cpu_baseline:
            cb = run_cpu(named_total, args.cpu_sample, 1, 0, args.cpu_form)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "form", "gflops_per_core", "sample")}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
