"""ctypes wrapper of oracle/xm_oracle_c.c — the compiled (C + OpenMP) twin of oracle/xm_oracle.py (TEST / BASELINE
INFRASTRUCTURE ONLY: tests/ and bench.py's cpu_baseline / --impl reference legs; never the product path).

The shared object is compiled on the machine that uses it (``gcc -O3 -march=native -fopenmp``, < 1 s) into
``oracle/_build/libxm_oracle_c.<cpu-tag>.so``: the build container and the GPU box have different CPUs, so a
``-march=native`` binary must not travel between them."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "xm_oracle_c.c")
_lib = None


def _cpu_tag() -> str:
    model = flags = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and not model:
                model = line.split(":", 1)[1].strip()
            if line.startswith("flags") and not flags:
                flags = line.split(":", 1)[1].strip()
            if model and flags:
                break
    except OSError:
        pass
    return hashlib.sha1((model + "|" + flags).encode()).hexdigest()[:10]


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def lib_path() -> str:
    return os.path.join(_HERE, "_build", f"libxm_oracle_c.{_cpu_tag()}.so")


def ensure_built(force: bool = False) -> str:
    out = lib_path()
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tmp = out + f".{os.getpid()}.tmp"
        subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-std=c11", "-D_POSIX_C_SOURCE=200809L", _SRC, "-o", tmp, "-lm"],
                       check=True)
        os.replace(tmp, out)
    return out


class _Result(C.Structure):
    _fields_ = [("outer_iters", C.c_int), ("tcg_iters", C.c_int), ("qy_products", C.c_int), ("status", C.c_int), ("n_log", C.c_int),
                ("primal", C.c_double), ("gradtol", C.c_double), ("gradnorm", C.c_double), ("wall_s", C.c_double)]


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(ensure_built())
        vp = C.c_void_p
        lib.xmo_trust_region.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_double, vp, C.c_double, C.c_int, vp, vp,
                                         C.POINTER(_Result), vp]
        lib.xmo_trust_region.restype = C.c_int
        lib.xmo_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(load().xmo_num_threads())


@dataclass
class TRResult:
    Y: np.ndarray
    s: np.ndarray
    primal: float
    gradtol: float
    outer_iters: int = 0
    tcg_iters: int = 0
    qy_products: int = 0
    gradnorm: float = float("nan")
    status: int = 0
    log: list = field(default_factory=list)
    wall_s: float = 0.0


def trust_region(Q, Y0, s0, lam, gradtol, ls_step=0.0, v=None, max_time=1000.0, replicate_stale_sr=True) -> TRResult:
    """Same arguments and result fields as xm_oracle.trust_region (Y0: (N, 3, r), s0: (N,), v: (3N,) or None)."""
    lib = load()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    Y0 = np.ascontiguousarray(Y0, dtype=np.float64); s0 = np.ascontiguousarray(s0, dtype=np.float64)
    N, _, r = Y0.shape
    if Q.shape != (3 * N, 3 * N) or r < 3 or r > 20:
        raise ValueError("bad shapes")
    Yo = np.empty_like(Y0); so = np.empty_like(s0)
    log = np.zeros((1002, 6))
    vv = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
    res = _Result()
    rc = lib.xmo_trust_region(N, r, Q.ctypes.data, Y0.ctypes.data, s0.ctypes.data, float(lam), float(gradtol), float(ls_step),
                              vv.ctypes.data if vv is not None else None, float(max_time), int(bool(replicate_stale_sr)),
                              Yo.ctypes.data, so.ctypes.data, C.byref(res), log.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"xmo_trust_region failed ({rc})")
    lg = [(int(a[0]), int(a[1]), float(a[2]), float(a[3]), int(a[4]), int(a[5])) for a in log[: res.n_log]]
    return TRResult(Y=Yo, s=so, primal=res.primal, gradtol=res.gradtol, outer_iters=res.outer_iters, tcg_iters=res.tcg_iters,
                    qy_products=res.qy_products, gradnorm=res.gradnorm, status=res.status, log=lg, wall_s=res.wall_s)
