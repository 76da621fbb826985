"""ctypes binding of libssqp_b200.so (include/ssqp_b200.h) — the same calls a Julia `ccall` makes.

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, every entry
point raises.  PyTorch is not needed here; bench.py uses it only for pinned/device buffers and
torch.distributed plumbing.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSQP_LIB") or os.path.join(_HERE, "libssqp_b200.so")      # (SSQP_LIB: developer builds, e.g. the timeline library)
NSTATS = 56
STAT_NAMES = ("trips", "falg", "maxK", "maxW", "lp_loops", "lp_pivots", "updates", "rebuilds", "maxres",
              "cycles", "bytes", "degen", "cyc_p1", "cyc_vpass", "cyc_cpass", "cyc_symv", "cyc_syr", "cyc_gamma",
              "cyc_p1_price", "cyc_rebuild", "cyc_ratio", "cyc_events", "cyc_kkt", "n_symv", "n_syr")
EXPORTS = ("ssqp_default_settings", "ssqp_create", "ssqp_destroy", "ssqp_set_shared", "ssqp_solve_batch", "ssqp_solve_sweep",
           "ssqp_solve_batch_device", "ssqp_solve_lp_batch", "ssqp_init_batch", "ssqp_get_stats", "ssqp_get_stats_device",
           "ssqp_set_free_var_capacity",
           "ssqp_launch_count", "ssqp_last_kernel_ms", "ssqp_measure_fp64_peak", "ssqp_measure_read_bw",
           "ssqp_last_error", "ssqp_last_launch_config", "ssqp_device_count", "ssqp_version")


class SsqpError(RuntimeError):
    pass


class CSettings(C.Structure):
    """ssqp_settings == struct Settings{Float64} (src/types.jl:390-408)."""
    _fields_ = [("max_iter", C.c_int32), ("tol", C.c_double), ("tolG", C.c_double), ("rule", C.c_int32),
                ("pivot", C.c_int32)]


_lib = None


def load():
    """dlopen the in-tree library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SsqpError("libssqp_b200.so is not built (%s). Run `python __graft_entry__.py` or "
                        "`python statusswitchingqp.jl_b200/build.py`; there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip, lp, vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    sp = C.POINTER(CSettings)
    L.ssqp_default_settings.argtypes = [sp]; L.ssqp_default_settings.restype = None
    L.ssqp_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32]; L.ssqp_create.restype = C.c_int
    L.ssqp_destroy.argtypes = [C.c_void_p]; L.ssqp_destroy.restype = C.c_int
    L.ssqp_set_shared.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, dp, dp, dp]; L.ssqp_set_shared.restype = C.c_int
    L.ssqp_solve_batch.argtypes = [C.c_void_p, C.c_int64] + [dp] * 6 + [ip, dp, sp, sp, dp, ip, lp]
    L.ssqp_solve_batch.restype = C.c_int
    L.ssqp_solve_sweep.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [dp] * 6 + [sp, sp, dp, ip, lp]
    L.ssqp_solve_sweep.restype = C.c_int
    L.ssqp_solve_batch_device.argtypes = [C.c_void_p, C.c_int64] + [dp] * 6 + [ip, dp, sp, sp, dp, ip, lp, vp]
    L.ssqp_solve_batch_device.restype = C.c_int
    L.ssqp_solve_lp_batch.argtypes = [C.c_void_p, C.c_int64] + [dp] * 5 + [sp, dp, ip, lp]; L.ssqp_solve_lp_batch.restype = C.c_int
    L.ssqp_init_batch.argtypes = [C.c_void_p, C.c_int64] + [dp] * 4 + [sp, dp, ip, lp]; L.ssqp_init_batch.restype = C.c_int
    L.ssqp_get_stats.argtypes = [C.c_void_p, C.c_int64, dp]; L.ssqp_get_stats.restype = C.c_int
    L.ssqp_set_free_var_capacity.argtypes = [C.c_void_p, C.c_int32]; L.ssqp_set_free_var_capacity.restype = C.c_int
    L.ssqp_get_stats_device.argtypes = [C.c_void_p, C.c_int64, dp]; L.ssqp_get_stats_device.restype = C.c_int
    L.ssqp_launch_count.argtypes = [C.c_void_p]; L.ssqp_launch_count.restype = C.c_int64
    L.ssqp_last_kernel_ms.argtypes = [C.c_void_p]; L.ssqp_last_kernel_ms.restype = C.c_double
    L.ssqp_measure_fp64_peak.argtypes = [C.c_void_p]; L.ssqp_measure_fp64_peak.restype = C.c_double
    L.ssqp_measure_read_bw.argtypes = [C.c_void_p, C.c_int32, C.c_int32]; L.ssqp_measure_read_bw.restype = C.c_double
    L.ssqp_last_error.argtypes = [C.c_void_p]; L.ssqp_last_error.restype = C.c_char_p
    L.ssqp_last_launch_config.argtypes = [C.c_void_p]; L.ssqp_last_launch_config.restype = C.c_char_p
    L.ssqp_device_count.argtypes = []; L.ssqp_device_count.restype = C.c_int32
    L.ssqp_version.argtypes = []; L.ssqp_version.restype = C.c_char_p
    _lib = L
    return L


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Context:
    """Owns an ssqp_ctx.  One per host thread; not thread-safe (include/ssqp_b200.h)."""

    def __init__(self, devices=None):
        L = load()
        if devices is None:
            devices = [0]
        ids = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = L.ssqp_create(C.byref(h), ids, len(devices))
        if rc != 0:
            raise SsqpError("ssqp_create failed (%d): %s — a CUDA device is required, there is no CPU fallback"
                            % (rc, (L.ssqp_last_error(None) or b"").decode()))
        self._h = h
        self._L = L
        self.N = self.M = self.J = None
        self.n_devices = len(devices)

    def close(self):
        if getattr(self, "_h", None):
            self._L.ssqp_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != 0:
            raise SsqpError("%s failed (%d): %s" % (what, rc, (self._L.ssqp_last_error(self._h) or b"").decode()))

    def set_shared(self, V, A, G):
        """V (N,N) or None; A (M,N); G (J,N) — row-major numpy arrays are converted to column-major."""
        A = np.asarray(A, dtype=np.float64); G = np.asarray(G, dtype=np.float64)
        N = A.shape[1] if A.ndim == 2 and A.shape[1] else (G.shape[1] if G.ndim == 2 and G.shape[1] else np.asarray(V).shape[0])
        A = A.reshape(-1, N); G = G.reshape(-1, N)
        M, J = A.shape[0], G.shape[0]
        Vf = None if V is None else np.asfortranarray(V, dtype=np.float64)
        Af, Gf = np.asfortranarray(A), np.asfortranarray(G)
        self._check(self._L.ssqp_set_shared(self._h, N, M, J, _ptr(Vf), _ptr(Af) if M else None, _ptr(Gf) if J else None),
                    "ssqp_set_shared")
        self.N, self.M, self.J = N, M, J

    def solve_batch(self, q, b, g, d, u, V_per_qp=None, S0=None, x0=None, settings=None, settingsLP=None):
        N, M, J = self.N, self.M, self.J
        q = _f64(q, (-1, N)); nb = q.shape[0]
        b = _f64(b, (nb, M)); g = _f64(g, (nb, J)); d = _f64(d, (nb, N)); u = _f64(u, (nb, N))
        Vq = None
        if V_per_qp is not None:
            Vq = _f64(V_per_qp, (nb, N, N))           # symmetric: row-major == column-major
        S0a = None if S0 is None else np.ascontiguousarray(S0, dtype=np.int32).reshape(nb, N + J)
        x0a = _f64(x0, (nb, N))
        x = np.empty((nb, N)); S = np.empty((nb, N + J), dtype=np.int32); status = np.empty(nb, dtype=np.int64)
        sp = C.byref(settings) if settings is not None else None
        slp = C.byref(settingsLP) if settingsLP is not None else None
        self._check(self._L.ssqp_solve_batch(self._h, nb, _ptr(Vq), _ptr(q), _ptr(b) if M else None, _ptr(g) if J else None,
                                             _ptr(d), _ptr(u), _ptr(S0a), _ptr(x0a), sp, slp, _ptr(x), _ptr(S), _ptr(status)),
                    "ssqp_solve_batch")
        return x, S, status

    def solve_sweep(self, q, b, g, d, u, chain_len, V_per_qp=None, settings=None, settingsLP=None):
        """Chains of `chain_len` consecutive QPs sharing b, g, d, u: each QP after the first of its chain is warm-started
        from the previous one's (x, S) — solveQP(Q, S, x0), src/SSQP.jl:237 — chains in parallel (ssqp_solve_sweep)."""
        N, M, J = self.N, self.M, self.J
        q = _f64(q, (-1, N)); nb = q.shape[0]
        b = _f64(b, (nb, M)); g = _f64(g, (nb, J)); d = _f64(d, (nb, N)); u = _f64(u, (nb, N))
        Vq = None if V_per_qp is None else _f64(V_per_qp, (nb, N, N))
        x = np.empty((nb, N)); S = np.empty((nb, N + J), dtype=np.int32); status = np.empty(nb, dtype=np.int64)
        sp = C.byref(settings) if settings is not None else None
        slp = C.byref(settingsLP) if settingsLP is not None else None
        self._check(self._L.ssqp_solve_sweep(self._h, nb, int(chain_len), _ptr(Vq), _ptr(q), _ptr(b) if M else None,
                                             _ptr(g) if J else None, _ptr(d), _ptr(u), sp, slp, _ptr(x), _ptr(S), _ptr(status)),
                    "ssqp_solve_sweep")
        return x, S, status

    def solve_lp_batch(self, c, b, g, d, u, settings=None):
        """Batch of LPs sharing A, G (set_shared(None, A, G)).  c,d,u: (nb,N); b: (nb,M); g: (nb,J)."""
        N, M, J = self.N, self.M, self.J
        c = _f64(c, (-1, N)); nb = c.shape[0]
        b = _f64(b, (nb, M)); g = _f64(g, (nb, J)); d = _f64(d, (nb, N)); u = _f64(u, (nb, N))
        x = np.empty((nb, N)); S = np.empty((nb, N + J), dtype=np.int32); status = np.empty(nb, dtype=np.int64)
        sp = C.byref(settings) if settings is not None else None
        self._check(self._L.ssqp_solve_lp_batch(self._h, nb, _ptr(c), _ptr(b) if M else None, _ptr(g) if J else None, _ptr(d),
                                                _ptr(u), sp, _ptr(x), _ptr(S), _ptr(status)), "ssqp_solve_lp_batch")
        return x, S, status

    def init_batch(self, b, g, d, u, settingsLP=None):
        N, M, J = self.N, self.M, self.J
        d = _f64(d, (-1, N)); nb = d.shape[0]
        b = _f64(b, (nb, M)); g = _f64(g, (nb, J)); u = _f64(u, (nb, N))
        x = np.empty((nb, N)); S = np.empty((nb, N + J), dtype=np.int32); status = np.empty(nb, dtype=np.int64)
        slp = C.byref(settingsLP) if settingsLP is not None else None
        self._check(self._L.ssqp_init_batch(self._h, nb, _ptr(b) if M else None, _ptr(g) if J else None, _ptr(d), _ptr(u),
                                            slp, _ptr(x), _ptr(S), _ptr(status)), "ssqp_init_batch")
        return x, S, status

    def solve_batch_device(self, nb, q, b, g, d, u, x, S, status, V_per_qp=0, S0=0, x0=0, settings=None,
                           settingsLP=None, stream=0):
        """All array arguments are raw device addresses (ints), e.g. torch.Tensor.data_ptr().  Enqueues only."""
        sp = C.byref(settings) if settings is not None else None
        slp = C.byref(settingsLP) if settingsLP is not None else None
        vp = lambda a: C.c_void_p(a) if a else None
        self._check(self._L.ssqp_solve_batch_device(self._h, nb, vp(V_per_qp), vp(q), vp(b), vp(g), vp(d), vp(u), vp(S0),
                                                    vp(x0), sp, slp, vp(x), vp(S), vp(status), vp(stream)),
                    "ssqp_solve_batch_device")

    def set_free_var_capacity(self, n):
        """Free variables per QP the device-pointer entry sizes Phase 1 for (the host entries scan the bounds themselves)."""
        self._check(self._L.ssqp_set_free_var_capacity(self._h, int(n)), "ssqp_set_free_var_capacity")

    def stats(self, nb, device=False):
        out = np.zeros((nb, NSTATS))
        fn = self._L.ssqp_get_stats_device if device else self._L.ssqp_get_stats
        self._check(fn(self._h, nb, _ptr(out)), "ssqp_get_stats")
        return out

    def launch_count(self):
        return int(self._L.ssqp_launch_count(self._h))

    def last_kernel_ms(self):
        return float(self._L.ssqp_last_kernel_ms(self._h))

    def last_launch_config(self):
        return (self._L.ssqp_last_launch_config(self._h) or b"").decode()

    def measure_fp64_peak(self):
        return float(self._L.ssqp_measure_fp64_peak(self._h))

    def measure_read_bw(self, mbytes, reps):
        return float(self._L.ssqp_measure_read_bw(self._h, mbytes, reps))


def device_count():
    return int(load().ssqp_device_count())


def version():
    return load().ssqp_version().decode()
