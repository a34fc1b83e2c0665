"""ctypes binding of the C-ABI in include/xm_b200.h (libxm_b200.so).

Fails loudly: if the shared library is missing or no sm_100 GPU is present there is NO fallback — ``load()`` /
``Handle()`` raise.  Arrays cross the boundary in the reference's wire layouts (3N x r column-major, s length N)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxm_b200.so")
XM_LOG_CAP = 1002
XM_MAX_RANK = 20

ERRORS = {0: "XM_OK", -1: "XM_EINVAL", -2: "XM_ECUDA", -3: "XM_ENOMEM", -4: "XM_ENOGPU", -5: "XM_ESYNC", -6: "XM_EUNSUPPORTED"}
EXIT_CODES = {1: "gradtol", 2: "rdotr_tiny", 3: "maxtime", 4: "model_increase", 5: "delta_tiny", 6: "max_outer", -1: "linesearch_failed", 0: "none"}


class XmOptions(C.Structure):
    _fields_ = [("device", C.c_int), ("grid_ctas", C.c_int), ("ksplit", C.c_int), ("replicate_stale_sr", C.c_int),
                ("verbose", C.c_int), ("max_outer", C.c_int), ("max_inner", C.c_int), ("qy_variant", C.c_int),
                ("vec_in_global", C.c_int), ("profile", C.c_int), ("three_barrier_tcg", C.c_int)]


class XmLogRec(C.Structure):
    _fields_ = [("k", C.c_int), ("inner_shown", C.c_int), ("trstatus", C.c_int), ("endreason", C.c_int),
                ("loss", C.c_double), ("gradnorm", C.c_double), ("delta", C.c_double)]


class XmStats(C.Structure):
    _fields_ = [("exit_code", C.c_int), ("outer_iters", C.c_int), ("tcg_iters", C.c_int), ("qy_products", C.c_int),
                ("n_log", C.c_int), ("primal", C.c_double), ("gradnorm", C.c_double), ("solve_ms", C.c_double),
                ("qy_ms", C.c_double), ("sync_ms", C.c_double), ("grid_ctas", C.c_int), ("threads_per_cta", C.c_int),
                ("ksplit", C.c_int), ("launches", C.c_int), ("phase_ms", C.c_double * 4)]


class XmCertInfo(C.Structure):
    _fields_ = [("certified", C.c_int), ("method", C.c_int), ("products", C.c_int), ("converged", C.c_int),
                ("min_eig", C.c_double), ("dual", C.c_double), ("gap", C.c_double), ("residual", C.c_double), ("ms", C.c_double)]


class XmSolveResult(C.Structure):
    _fields_ = [("rank", C.c_int), ("status", C.c_int), ("n_solves", C.c_int), ("certified", C.c_int), ("cert_method", C.c_int),
                ("tcg_iters_total", C.c_int), ("qy_products_total", C.c_int), ("cert_products_total", C.c_int),
                ("primal", C.c_double), ("gradnorm", C.c_double), ("min_eig", C.c_double), ("dual", C.c_double), ("gap", C.c_double),
                ("solve_ms_total", C.c_double), ("cert_ms_total", C.c_double)]


CERT_METHODS = {"auto": 0, "dense": 1, "iterative": 2}
MODES = {"full": 0, "rank3": 1, "rebuttle": 2}

EXPORTS = [
    "xm_default_options", "xm_create", "xm_destroy", "xm_last_error", "xm_set_stream", "xm_set_q_dense",
    "xm_set_q_dense_dev", "xm_set_q_bsr", "xm_qy", "xm_qy_dev", "xm_trust_region", "xm_trust_region_dev",
    "xm_op_objective", "xm_op_rgrad", "xm_op_rhess", "xm_op_retract", "xm_certify", "xm_escape_scale", "xm_bench_qy",
    "xm_bench_barrier", "xm_debug_trace",
    "xm_partition", "xm_comm_init", "xm_comm_connect", "xm_comm_connect_ptrs", "xm_comm_arena", "xm_comm_info", "xm_comm_reset", "xm_comm_disconnect",
    "xm_set_q_dense_slab", "xm_set_q_dense_slab_dev", "xm_recover", "xm_residuals", "xm_debug_counters",
    "xm_certify_ex", "xm_op_diag_blocks", "xm_solve", "xm_create_matrix", "xm_comm_halo", "xm_rcm_order",
]
XM_IPC_HANDLE_BYTES = 64
XM_MAX_WORLD = 8

_lib = None


def load(path: str | None = None):
    """dlopen libxm_b200.so (raises OSError if it has not been built: run ``python __graft_entry__.py``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise OSError(f"{p} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                      "there is no CPU fallback for the XM hot path")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    lib.xm_default_options.argtypes = [C.POINTER(XmOptions)]
    lib.xm_default_options.restype = None
    lib.xm_create.argtypes = [C.POINTER(vp), C.POINTER(XmOptions)]
    lib.xm_destroy.argtypes = [vp]
    lib.xm_last_error.argtypes = [vp]
    lib.xm_last_error.restype = C.c_char_p
    lib.xm_set_stream.argtypes = [vp, vp]
    lib.xm_set_q_dense.argtypes = [vp, C.c_int, vp, C.c_int64]
    lib.xm_set_q_dense_dev.argtypes = [vp, C.c_int, vp, C.c_int64]
    lib.xm_set_q_bsr.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.xm_qy.argtypes = [vp, C.c_int, C.c_double, vp, vp]
    lib.xm_qy_dev.argtypes = [vp, C.c_int, C.c_double, vp, vp]
    tr_args = [vp, C.c_int, vp, vp, C.c_double, dp, C.c_double, vp, C.c_double, vp, vp, dp, C.POINTER(XmStats), vp]
    lib.xm_trust_region.argtypes = tr_args
    lib.xm_trust_region_dev.argtypes = tr_args
    lib.xm_op_objective.argtypes = [vp, C.c_int, vp, vp, C.c_double, dp]
    lib.xm_op_rgrad.argtypes = [vp, C.c_int, vp, vp, C.c_double, vp, vp, dp]
    lib.xm_op_rhess.argtypes = [vp, C.c_int, vp, vp, C.c_double, vp, vp, vp, vp]
    lib.xm_op_retract.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_double, vp, vp]
    lib.xm_certify.argtypes = [vp, C.c_int, vp, vp, C.c_double, C.c_double, vp, dp, dp, dp, C.POINTER(C.c_int)]
    lib.xm_escape_scale.argtypes = [C.c_int, vp, vp]
    lib.xm_bench_qy.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.xm_bench_barrier.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.xm_debug_trace.argtypes = [vp, vp]
    lib.xm_debug_counters.argtypes = [vp, vp]
    ip = C.POINTER(C.c_int)
    lib.xm_partition.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
    lib.xm_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    lib.xm_comm_connect.argtypes = [vp, vp]
    lib.xm_comm_connect_ptrs.argtypes = [vp, C.POINTER(vp)]
    lib.xm_comm_arena.argtypes = [vp]
    lib.xm_comm_arena.restype = vp
    lib.xm_comm_info.argtypes = [vp, ip, ip, ip, ip, ip]
    lib.xm_comm_reset.argtypes = [vp]
    lib.xm_comm_disconnect.argtypes = [vp]
    lib.xm_set_q_dense_slab.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64]
    lib.xm_set_q_dense_slab_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64]
    lib.xm_recover.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int64, vp, vp, vp, vp, ip]
    lib.xm_residuals.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.xm_certify_ex.argtypes = [vp, C.c_int, vp, vp, C.c_double, C.c_double, C.c_int, vp, C.POINTER(XmCertInfo)]
    lib.xm_op_diag_blocks.argtypes = [vp, vp]
    lib.xm_solve.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, vp, C.c_int, vp, vp, C.POINTER(XmSolveResult)]
    lib.xm_create_matrix.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, dp]
    lib.xm_comm_halo.argtypes = [vp, ip, C.POINTER(C.c_longlong), ip]
    lib.xm_rcm_order.argtypes = [C.c_int, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("xm_default_options", "xm_last_error", "xm_comm_arena"):
            fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


class XmError(RuntimeError):
    pass


def partition(n_cameras: int, world: int, ctas_per_rank: int, rank: int):
    """Cameras [lo, hi) owned by `rank` (pure host arithmetic in the library; works without a GPU)."""
    lib = load()
    lo, hi = C.c_int(), C.c_int()
    rc = lib.xm_partition(n_cameras, world, ctas_per_rank, rank, C.byref(lo), C.byref(hi))
    if rc != 0:
        raise XmError(f"xm_partition: {ERRORS.get(rc, rc)}")
    return lo.value, hi.value


def rcm_order(rowptr, colidx) -> np.ndarray:
    """xm_rcm_order: reverse Cuthill-McKee camera order of a block-CSR view graph (perm[new] = old); host only."""
    lib = load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32); colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    perm = np.empty(rowptr.size - 1, dtype=np.int32)
    rc = lib.xm_rcm_order(rowptr.size - 1, rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p), perm.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise XmError(f"xm_rcm_order: {ERRORS.get(rc, rc)}")
    return perm.astype(np.int64)


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS", "ALIGNED"])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class TRResult:
    R: np.ndarray          # 3N x r
    s: np.ndarray          # N, s[0] == 1
    primal: float
    gradtol: float
    stats: dict
    log: list


class Handle:
    """One xm_handle (one CUDA device)."""

    def __init__(self, device: int = 0, grid_ctas: int = 0, ksplit: int = 0, replicate_stale_sr: bool = True,
                 verbose: bool = False, max_outer: int = 1000, max_inner: int = 1000, qy_variant: int = 0,
                 vec_in_global: bool = False, profile: bool = False, three_barrier_tcg: bool = False):
        self.lib = load()
        opt = XmOptions()
        self.lib.xm_default_options(C.byref(opt))
        opt.device = device; opt.grid_ctas = grid_ctas; opt.ksplit = ksplit
        opt.replicate_stale_sr = int(replicate_stale_sr); opt.verbose = int(verbose)
        opt.max_outer = max_outer; opt.max_inner = max_inner; opt.qy_variant = qy_variant
        opt.vec_in_global = int(vec_in_global); opt.profile = int(profile); opt.three_barrier_tcg = int(three_barrier_tcg)
        self._h = C.c_void_p()
        rc = self.lib.xm_create(C.byref(self._h), C.byref(opt))
        if rc != 0:
            raise XmError(f"xm_create failed: {ERRORS.get(rc, rc)} (an sm_100 GPU is required; no CPU fallback)")
        self.N = 0
        self.is_bsr = False

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.xm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.xm_last_error(self._h)
            raise XmError(f"{what}: {ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")

    # ---- multi-GPU communicator (camera partition; include/xm_b200.h "multi-GPU")
    def comm_init(self, rank: int, world: int, n_cameras: int, max_r: int = 5) -> bytes:
        buf = (C.c_ubyte * XM_IPC_HANDLE_BYTES)()
        self._check(self.lib.xm_comm_init(self._h, rank, world, n_cameras, max_r, C.cast(buf, C.c_void_p)), "xm_comm_init")
        return bytes(buf)

    def comm_connect(self, handles):
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self.lib.xm_comm_connect(self._h, C.cast(buf, C.c_void_p)), "xm_comm_connect")

    def comm_connect_ptrs(self, ptrs):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._check(self.lib.xm_comm_connect_ptrs(self._h, arr), "xm_comm_connect_ptrs")

    def comm_arena(self) -> int:
        return int(self.lib.xm_comm_arena(self._h) or 0)

    def comm_info(self) -> dict:
        v = [C.c_int() for _ in range(5)]
        self._check(self.lib.xm_comm_info(self._h, *[C.byref(x) for x in v]), "xm_comm_info")
        return dict(zip(("rank", "world", "ctas_per_rank", "cam_lo", "cam_hi"), (x.value for x in v)))

    def comm_halo(self) -> dict:
        """Boundary-only exchange of the current block-CSR operator: cameras unpacked / (camera, peer) pairs pushed per exchange."""
        a, b, c = C.c_int(), C.c_longlong(), C.c_int()
        self._check(self.lib.xm_comm_halo(self._h, C.byref(a), C.byref(b), C.byref(c)), "xm_comm_halo")
        return dict(need=a.value, sent=b.value, remote=c.value)

    def comm_disconnect(self):
        self._check(self.lib.xm_comm_disconnect(self._h), "xm_comm_disconnect")

    def comm_reset(self):
        self._check(self.lib.xm_comm_reset(self._h), "xm_comm_reset")

    def set_q_dense_slab(self, Q_slab, row0: int):
        """Q_slab: this rank's rows of Q, shape (nrows, 3N)."""
        Q_slab = _f64(Q_slab)
        nrows, n3 = Q_slab.shape
        self._check(self.lib.xm_set_q_dense_slab(self._h, n3, row0, nrows, _ptr(Q_slab), nrows), "xm_set_q_dense_slab")
        self.N = n3 // 3
        self.is_bsr = False

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.xm_set_stream(self._h, C.c_void_p(cuda_stream)), "xm_set_stream")

    # ---- operator
    def set_q_dense(self, Q):
        Q = _f64(Q)
        n3 = Q.shape[0]
        self._check(self.lib.xm_set_q_dense(self._h, n3, _ptr(Q), n3), "xm_set_q_dense")
        self.N = n3 // 3
        self.is_bsr = False

    def set_q_dense_ptr(self, n3: int, host_ptr: int, ld: int | None = None):
        self._check(self.lib.xm_set_q_dense(self._h, n3, C.c_void_p(host_ptr), ld or n3), "xm_set_q_dense")
        self.N = n3 // 3
        self.is_bsr = False

    def set_q_dense_dev(self, n3: int, dev_ptr: int, ld: int | None = None):
        self._check(self.lib.xm_set_q_dense_dev(self._h, n3, C.c_void_p(dev_ptr), ld or n3), "xm_set_q_dense_dev")
        self.N = n3 // 3
        self.is_bsr = False

    def set_q_bsr(self, rowptr, colidx, vals, bdim: int):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        colidx = np.ascontiguousarray(colidx, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        nb = rowptr.size - 1
        self._check(self.lib.xm_set_q_bsr(self._h, nb, bdim, _ptr(rowptr), _ptr(colidx), _ptr(vals)), "xm_set_q_bsr")
        self.N = nb
        self.is_bsr = True

    # ---- ops
    def qy(self, X, alpha: float = 1.0):
        X = _f64(X)
        out = np.empty_like(X, order="F")
        self._check(self.lib.xm_qy(self._h, X.shape[1], alpha, _ptr(X), _ptr(out)), "xm_qy")
        return out

    def qy_dev(self, r: int, x_dev_ptr: int, out_dev_ptr: int, alpha: float = 1.0):
        self._check(self.lib.xm_qy_dev(self._h, r, alpha, C.c_void_p(x_dev_ptr), C.c_void_p(out_dev_ptr)), "xm_qy_dev")

    def bench_qy(self, r: int, iters: int) -> float:
        ms = C.c_double()
        self._check(self.lib.xm_bench_qy(self._h, r, iters, C.byref(ms)), "xm_bench_qy")
        return ms.value

    def debug_trace(self):
        buf = (C.c_ulonglong * 256)()
        self._check(self.lib.xm_debug_trace(self._h, C.cast(buf, C.c_void_p)), "xm_debug_trace")
        out = [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(128) if buf[2 * i + 1]]
        return out

    def debug_counters(self):
        buf = (C.c_ulonglong * 8)()
        self._check(self.lib.xm_debug_counters(self._h, C.cast(buf, C.c_void_p)), "xm_debug_counters")
        return [int(x) for x in buf]

    def bench_barrier(self, r: int, iters: int) -> float:
        us = C.c_double()
        self._check(self.lib.xm_bench_barrier(self._h, r, iters, C.byref(us)), "xm_bench_barrier")
        return us.value

    def objective(self, R, s, lam=0.0) -> float:
        R = _f64(R); s = _f64(s)
        f = C.c_double()
        self._check(self.lib.xm_op_objective(self._h, R.shape[1], _ptr(R), _ptr(s), lam, C.byref(f)), "xm_op_objective")
        return f.value

    def rgrad(self, R, s, lam=0.0):
        R = _f64(R); s = _f64(s)
        gR = np.empty_like(R, order="F"); gs = np.empty_like(s)
        gn = C.c_double()
        self._check(self.lib.xm_op_rgrad(self._h, R.shape[1], _ptr(R), _ptr(s), lam, _ptr(gR), _ptr(gs), C.byref(gn)), "xm_op_rgrad")
        return gR, gs, gn.value

    def rhess(self, R, s, P, ps, lam=0.0):
        R = _f64(R); s = _f64(s); P = _f64(P); ps = _f64(ps)
        HR = np.empty_like(R, order="F"); Hs = np.empty_like(s)
        self._check(self.lib.xm_op_rhess(self._h, R.shape[1], _ptr(R), _ptr(s), lam, _ptr(P), _ptr(ps), _ptr(HR), _ptr(Hs)), "xm_op_rhess")
        return HR, Hs

    def retract(self, R, s, etaR, etas, lr=1.0):
        R = _f64(R); s = _f64(s); etaR = _f64(etaR); etas = _f64(etas)
        Rn = np.empty_like(R, order="F"); sn = np.empty_like(s)
        self._check(self.lib.xm_op_retract(self._h, R.shape[1], _ptr(R), _ptr(s), _ptr(etaR), _ptr(etas), lr, _ptr(Rn), _ptr(sn)), "xm_op_retract")
        return Rn, sn

    # ---- solver
    def trust_region(self, R0, s0, lam=0.0, gradtol=1e-6, ls_step=0.0, v=None, max_time=1000.0) -> TRResult:
        R0 = _f64(R0); s0 = _f64(s0)
        r = R0.shape[1]
        R = np.empty_like(R0, order="F"); s = np.empty_like(s0)
        gt = C.c_double(gradtol); primal = C.c_double()
        st = XmStats()
        log = (XmLogRec * XM_LOG_CAP)()
        vv = _f64(v) if v is not None else None
        rc = self.lib.xm_trust_region(self._h, r, _ptr(R0), _ptr(s0), lam, C.byref(gt), ls_step,
                                      _ptr(vv) if vv is not None else None, max_time, _ptr(R), _ptr(s),
                                      C.byref(primal), C.byref(st), C.cast(log, C.c_void_p))
        self._check(rc, "xm_trust_region")
        stats = {f[0]: getattr(st, f[0]) for f in XmStats._fields_}
        stats["phase_ms"] = list(st.phase_ms)
        stats["exit"] = EXIT_CODES.get(st.exit_code, str(st.exit_code))
        lg = [(log[i].k, log[i].inner_shown, log[i].loss, log[i].gradnorm, log[i].trstatus, log[i].endreason, log[i].delta)
              for i in range(st.n_log)]
        return TRResult(R=R, s=s, primal=primal.value, gradtol=gt.value, stats=stats, log=lg)

    def trust_region_dev(self, r, R0_ptr, s0_ptr, R_ptr, s_ptr, lam=0.0, gradtol=1e-6, ls_step=0.0, v_ptr=None, max_time=1000.0):
        gt = C.c_double(gradtol); primal = C.c_double()
        st = XmStats()
        rc = self.lib.xm_trust_region_dev(self._h, r, C.c_void_p(R0_ptr), C.c_void_p(s0_ptr), lam, C.byref(gt), ls_step,
                                          C.c_void_p(v_ptr) if v_ptr else None, max_time, C.c_void_p(R_ptr), C.c_void_p(s_ptr),
                                          C.byref(primal), C.byref(st), None)
        self._check(rc, "xm_trust_region_dev")
        stats = {f[0]: getattr(st, f[0]) for f in XmStats._fields_}
        stats["phase_ms"] = list(st.phase_ms)
        stats["exit"] = EXIT_CODES.get(st.exit_code, str(st.exit_code))
        return primal.value, gt.value, stats

    def recover(self, R, s, Abar=None):
        """xm_recover: returns dict(R (3 x 3N), s (N,), t (3 x N) / p (3 x M) when Abar is given, eigvals, negative)."""
        R = _f64(R); s = _f64(np.asarray(s).reshape(-1))
        n3, r = R.shape
        N = n3 // 3
        Ro = np.empty((3, n3), order="F"); so = np.empty(N); eig = np.empty(r); neg = C.c_int()
        y = None; rows = 0; ab = None
        if Abar is not None:
            ab = _f64(Abar); rows = ab.shape[0]
            y = np.empty((3, rows + 1), order="F")
        self._check(self.lib.xm_recover(self._h, N, r, _ptr(R), _ptr(s), _ptr(ab) if ab is not None else None, rows, _ptr(Ro), _ptr(so),
                                        _ptr(y) if y is not None else None, _ptr(eig), C.byref(neg)), "xm_recover")
        out = dict(R=Ro, s=so, eigvals=eig if r > 3 else None, negative=neg.value, t=None, p=None)
        if y is not None:
            out["t"] = y[:, :N]; out["p"] = y[:, N:]
        return out

    def residuals(self, cam, lm, pts, w, R_real, s_real, t, p):
        """xm_residuals: weighted squared residual per observation (cam, lm 0-based)."""
        cam = np.ascontiguousarray(cam, dtype=np.int32); lm = np.ascontiguousarray(lm, dtype=np.int32)
        pts = np.ascontiguousarray(pts, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
        R_real = _f64(R_real); t = _f64(t); p = _f64(p); s_real = _f64(np.asarray(s_real).reshape(-1))
        N = s_real.size; M = p.shape[1]
        err = np.empty(cam.size)
        self._check(self.lib.xm_residuals(self._h, cam.size, N, M, _ptr(cam), _ptr(lm), _ptr(pts), _ptr(w), _ptr(R_real), _ptr(s_real),
                                          _ptr(t), _ptr(p), _ptr(err)), "xm_residuals")
        return err

    def certify(self, R, s, lam, primal, method: str = "auto"):
        """xm_certify_ex: method "auto" | "dense" (cuSOLVER syevd on the assembled dual slack) | "iterative" (block Davidson on the
        Q.Y operator; block-CSR / communicator capable)."""
        R = _f64(R); s = _f64(s)
        v = np.empty(R.shape[0]); ci = XmCertInfo()
        self._check(self.lib.xm_certify_ex(self._h, R.shape[1], _ptr(R), _ptr(s), lam, primal, CERT_METHODS[method], _ptr(v), C.byref(ci)),
                    "xm_certify_ex")
        return dict(certified=bool(ci.certified), min_eig=ci.min_eig, dual=ci.dual, gap=ci.gap, v=v,
                    method={1: "dense", 2: "iterative"}.get(ci.method, str(ci.method)), products=ci.products, converged=bool(ci.converged),
                    residual=ci.residual, ms=ci.ms)

    def diag_blocks(self):
        out = np.empty((self.N, 3, 3))
        self._check(self.lib.xm_op_diag_blocks(self._h, _ptr(out)), "xm_op_diag_blocks")
        return out

    def solve(self, max_rank: int, tol: float, lam: float, max_time: float = 1000.0, mode: str = "full", s_init=None, cert_method: str = "auto"):
        """xm_solve: the reference's rank staircase (solve / solve_rank3 / solve_rebuttle) on the handle's operator."""
        N = self.N
        cap = max(3, min(int(max_rank), XM_MAX_RANK))
        R = np.zeros((3 * N, cap), order="F"); s = np.empty(N); res = XmSolveResult()
        si = _f64(np.asarray(s_init).reshape(-1)) if s_init is not None else None
        self._check(self.lib.xm_solve(self._h, MODES[mode], int(max_rank), tol, lam, max_time, _ptr(si) if si is not None else None,
                                      CERT_METHODS[cert_method], _ptr(R), _ptr(s), C.byref(res)), "xm_solve")
        out = {f[0]: getattr(res, f[0]) for f in XmSolveResult._fields_}
        out["R"] = np.asfortranarray(R[:, :res.rank]); out["s"] = s
        out["certificate_method"] = {1: "dense", 2: "iterative"}.get(res.cert_method, "none")
        return out

    def create_matrix(self, n_cameras: int, n_landmarks: int, cam, lm, w, pts, want_q: bool = True, want_abar: bool = False):
        """xm_create_matrix: Q (and Abar) from observations, assembled on the device; the result becomes the handle's operator.
        cam, lm: 0-based.  Returns (Q or None, Abar or None, assemble_ms)."""
        cam = np.ascontiguousarray(cam, dtype=np.int32); lm = np.ascontiguousarray(lm, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float64); pts = np.ascontiguousarray(pts, dtype=np.float64)
        n3 = 3 * n_cameras
        Q = np.empty((n3, n3), order="F") if want_q else None
        A = np.empty((n_cameras + n_landmarks - 1, n3), order="F") if want_abar else None
        ms = C.c_double()
        self._check(self.lib.xm_create_matrix(self._h, n_cameras, n_landmarks, cam.size, _ptr(cam), _ptr(lm), _ptr(w), _ptr(pts),
                                              _ptr(Q) if Q is not None else None, _ptr(A) if A is not None else None, C.byref(ms)), "xm_create_matrix")
        self.N = n_cameras
        self.is_bsr = False
        return Q, A, ms.value
