"""ctypes wrapper of oracle/_build/liboracle.so -- the CPU restatement of the reference's
SQP + OSQP path.  TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs only."""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
_lib = None

# layout shared with oracle_capi.cpp:apply_settings
SETTINGS_DEFAULT = dict(rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, eps_prim_inf=1e-4,
                        eps_dual_inf=1e-4, max_iter=10000, scaling=10, check_termination=25, adaptive_rho=1,
                        adaptive_rho_interval=0, adaptive_rho_tolerance=5.0)
_ORDER = list(SETTINGS_DEFAULT)


def settings_vector(**kw) -> np.ndarray:
    d = dict(SETTINGS_DEFAULT)
    for k, v in kw.items():
        if k not in d:
            raise KeyError(k)
        d[k] = v
    return np.array([float(d[k]) for k in _ORDER])


def settings_from_b200(s) -> np.ndarray:
    """Oracle settings equal to an optimal_control_problem_b200.Settings struct."""
    return settings_vector(rho=s.rho, sigma=s.sigma, alpha=s.relax, eps_abs=s.eps_abs, eps_rel=s.eps_rel,
                           eps_prim_inf=s.eps_prim_inf, eps_dual_inf=s.eps_dual_inf, max_iter=s.admm_max_iter,
                           scaling=s.scaling_iters, check_termination=s.check_termination,
                           adaptive_rho=s.adaptive_rho, adaptive_rho_interval=s.adaptive_rho_interval,
                           adaptive_rho_tolerance=s.adaptive_rho_tolerance)


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = ROOT / "oracle" / "_build" / "liboracle.so"
        if not path.exists():
            sys.path.insert(0, str(ROOT / "oracle"))
            import build as oracle_build  # noqa: WPS433
            oracle_build.build()
        L = C.CDLL(str(path))
        L.oracle_last_error.restype = C.c_char_p
        dptr, iptr, vptr = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.oracle_problem_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.POINTER(vptr)]
        L.oracle_problem_destroy.argtypes = [vptr]
        L.oracle_problem_dims.argtypes = [vptr, iptr]
        L.oracle_problem_patterns.argtypes = [vptr, iptr, iptr, iptr, iptr]
        L.oracle_problem_bounds.argtypes = [vptr, dptr, dptr, dptr, dptr]
        L.oracle_problem_set_qp_settings.argtypes = [vptr, dptr]
        L.oracle_problem_set_schedule.argtypes = [vptr, C.c_int, C.c_double, C.c_int]
        L.oracle_local_system.argtypes = [vptr] + [dptr] * 8
        L.oracle_objective.argtypes = [vptr, dptr, dptr, dptr]
        L.oracle_sqp_solve_batch.argtypes = [vptr, C.c_int, dptr, dptr, dptr, dptr, dptr, C.c_int, C.c_int]
        L.oracle_qp_solve.argtypes = [C.c_int, C.c_int, iptr, iptr, dptr, dptr, iptr, iptr, dptr, dptr, dptr, dptr,
                                      dptr, dptr, dptr, dptr, C.c_int, iptr, C.c_int]
        L.oracle_kat_solve.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, dptr, iptr, dptr, dptr]
        L.oracle_kat_expected.argtypes = [C.c_int, dptr, iptr]
        L.oracle_kat_patterns.argtypes = [C.c_int, iptr, iptr, iptr, iptr, iptr]
        L.oracle_sample_inputs.argtypes = [C.c_char_p, C.c_int, C.c_ulonglong, dptr, dptr]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().oracle_last_error().decode())


def hardware_threads() -> int:
    return int(lib().oracle_hardware_threads())


class OracleProblem:
    """Restated reference CPU path for one named problem."""

    def __init__(self, name: str, horizon: int = 0, alpha: float = 0.1, step_num: int = 10):
        L = lib()
        self.name = name
        h = C.c_void_p()
        _check(L.oracle_problem_create(name.encode(), horizon, alpha, step_num, C.byref(h)))
        self._h = h
        dims = (C.c_int * 8)()
        _check(L.oracle_problem_dims(h, dims))
        self.np_, self.nf, self.horizon, self.ng, self.n, self.m, self.nnz_h, self.nnz_a = list(dims)
        self.N = self.nf * self.horizon
        self.h_colptr = np.empty(self.n + 1, np.int32); self.h_rowidx = np.empty(self.nnz_h, np.int32)
        self.a_colptr = np.empty(self.n + 1, np.int32); self.a_rowidx = np.empty(self.nnz_a, np.int32)
        _check(L.oracle_problem_patterns(h, _ip(self.h_colptr), _ip(self.h_rowidx), _ip(self.a_colptr), _ip(self.a_rowidx)))
        self.lbx = np.empty(self.N); self.ubx = np.empty(self.N); self.lbg = np.empty(self.ng); self.ubg = np.empty(self.ng)
        _check(L.oracle_problem_bounds(h, _dp(self.lbx), _dp(self.ubx), _dp(self.lbg), _dp(self.ubg)))

    def __del__(self):
        try:
            if self._h:
                lib().oracle_problem_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_qp_settings(self, vec: np.ndarray) -> None:
        vec = np.ascontiguousarray(vec, np.float64)
        _check(lib().oracle_problem_set_qp_settings(self._h, _dp(vec)))

    def set_schedule(self, step_num: int, alpha: float, reuse_symbolic: bool = False) -> None:
        _check(lib().oracle_problem_set_schedule(self._h, step_num, alpha, int(reuse_symbolic)))

    def sample_inputs(self, B: int, seed: int):
        frames = np.empty((B, self.nf)); refs = np.empty((B, self.np_))
        _check(lib().oracle_sample_inputs(self.name.encode(), B, seed, _dp(frames), _dp(refs)))
        return frames, refs

    def local_system(self, frame, p, x):
        hv = np.empty(self.nnz_h); q = np.empty(self.n); av = np.empty(self.nnz_a); l = np.empty(self.m); u = np.empty(self.m)
        frame = None if frame is None else np.ascontiguousarray(frame, np.float64)
        p = np.ascontiguousarray(p, np.float64); x = np.ascontiguousarray(x, np.float64)
        _check(lib().oracle_local_system(self._h, _dp(frame), _dp(p), _dp(x), _dp(hv), _dp(q), _dp(av), _dp(l), _dp(u)))
        return hv, q, av, l, u

    def objective(self, p, x) -> float:
        f = np.zeros(1)
        _check(lib().oracle_objective(self._h, _dp(np.ascontiguousarray(p, np.float64)),
                                      _dp(np.ascontiguousarray(x, np.float64)), _dp(f)))
        return float(f[0])

    def solve_batch(self, frames, p, x0=None, nthreads: int = 1, use_float: bool = False):
        frames = np.ascontiguousarray(frames, np.float64)
        B = frames.size // self.nf
        p = np.ascontiguousarray(p, np.float64)
        x = np.zeros((B, self.N)) if x0 is None else np.array(x0, dtype=np.float64, order="C").reshape(B, self.N)
        f = np.empty(B); st = np.empty((B, 12))
        _check(lib().oracle_sqp_solve_batch(self._h, B, _dp(frames), _dp(p), _dp(x), _dp(f), _dp(st), nthreads,
                                            int(use_float)))
        return x, f, st


def qp_solve(n, m, hp, hi, hx, q, ap, ai, ax, l, u, settings=None, use_float=False, max_trace=512):
    """One QP through the OSQP restatement -> (x, y, info[8], trace[k, 6])."""
    hp = np.ascontiguousarray(hp, np.int32); hi = np.ascontiguousarray(hi, np.int32)
    ap = np.ascontiguousarray(ap, np.int32); ai = np.ascontiguousarray(ai, np.int32)
    f = lambda a: np.ascontiguousarray(a, np.float64)
    hx, q, ax, l, u = f(hx), f(q), f(ax), f(l), f(u)
    sv = None if settings is None else f(settings)
    x = np.empty(n); y = np.empty(m); info = np.zeros(8); trace = np.zeros((max_trace, 6)); nt = C.c_int(0)
    _check(lib().oracle_qp_solve(n, m, _ip(hp), _ip(hi), _dp(hx), _dp(q), _ip(ap), _ip(ai), _dp(ax), _dp(l), _dp(u),
                                 _dp(sv), _dp(x), _dp(y), _dp(info), _dp(trace), max_trace, C.byref(nt), int(use_float)))
    return x, y, info, trace[: nt.value]


def kat_solve(case: int, step_num: int = 1, alpha: float = 1.0, use_float: bool = False):
    x = np.zeros(8); n = C.c_int(0); f = np.zeros(1); st = np.zeros(12)
    _check(lib().oracle_kat_solve(case, step_num, alpha, int(use_float), _dp(x), C.byref(n), _dp(f), _dp(st)))
    return x[: n.value].copy(), float(f[0]), st


def kat_patterns(case: int):
    """(h_colptr, h_rowidx, a_colptr, a_rowidx) of the local system of test/test.cpp case `case` as casadi-lite builds it."""
    sizes = (C.c_int * 4)()
    hp = np.zeros(64, np.int32); hi = np.zeros(64, np.int32); ap = np.zeros(64, np.int32); ai = np.zeros(64, np.int32)
    _check(lib().oracle_kat_patterns(case, sizes, _ip(hp), _ip(hi), _ip(ap), _ip(ai)))
    n, m, nh, na = list(sizes)
    return hp[:n + 1].copy(), hi[:nh].copy(), ap[:n + 1].copy(), ai[:na].copy()


def kat_expected(case: int) -> np.ndarray:
    x = np.zeros(8); n = C.c_int(0)
    _check(lib().oracle_kat_expected(case, _dp(x), C.byref(n)))
    return x[: n.value].copy()
