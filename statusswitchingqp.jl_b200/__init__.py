"""statusswitchingqp.jl_b200 — B200-native batched status-switching active-set QP (the `solveQP` hot path of
PharosAbad/StatusSwitchingQP.jl).  Import through the root shim:  `import ssqp_b200`.

Layout: csrc/ (CUDA kernels + C ABI), capi.py (ctypes == the Julia ccall surface), types.py / solver.py
(host-side mirror of the reference interface), moi.py (mirror of the MOI wrapper's optimize! / status layer), workloads.py (synthetic BASELINE configs), julia/ (ccall glue).
"""
from .capi import Context, SsqpError, CSettings, device_count, version, load, LIB_PATH, NSTATS, STAT_NAMES, EXPORTS
from .types import Status, Settings, QP, LP, IN, DN, UP, OE, EO
from .solver import solveQP, solveQP_batch, solveQP_sweep, initQP_batch, SimplexLP, SimplexLP_batch, context
from . import workloads
from . import moi
from .moi import Optimizer, optimize_batch
from .build import build

__all__ = ["Context", "SsqpError", "CSettings", "device_count", "version", "load", "Status", "Settings", "QP",
           "LP", "IN", "DN", "UP", "OE", "EO", "solveQP", "solveQP_batch", "solveQP_sweep", "initQP_batch", "SimplexLP", "SimplexLP_batch", "context", "workloads", "build",
           "moi", "Optimizer", "optimize_batch"]
