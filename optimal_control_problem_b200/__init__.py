"""optimal_control_problem_b200 -- B200-native CUDA_SQP solve path of
LockedFlysher/optimal_control_problem.

The product is native: CUDA kernels + the C ABI of ``include/ocp_b200.h`` (``lib/libocp_b200.so``)
and the C++ front-end that keeps the reference's class surface (``lib/libocp_b200_host.so``).
This module is only the ctypes mirror used by the tests, ``bench.py`` and Python callers:

* :class:`Problem`  -- an ``OptimalControlProblem`` subclass instance (reference
  ``include/optimal_control_problem/OptimalControlProblem.h:65-107``): construct from YAML,
  ``genSolver()``, ``computeOptimalTrajectory()`` and its batched sibling;
* :class:`Solver`   -- the raw ``ocp_b200_*`` C ABI on numpy buffers / device pointers;
* :class:`KatProblem` -- the NLPs of the reference's ``test/test.cpp`` through
  ``SQPOptimizationSolver::getOptimalSolution``.

There is no CPU fallback: if the native libraries are missing, or no B200 is visible, the calls
raise.  Nothing in here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
ROOT = _PKG.parent
SHARE_DIR = Path(os.environ.get("OCP_B200_SHARE_DIR", _PKG / "share"))

NSTATS = 12
NINFO = 8
TRACE_WIDTH = 6
STAT = {"qp_status": 0, "sqp_steps": 1, "admm_iters": 2, "pcg_iters": 3, "prim_res": 4, "dual_res": 5,
        "objective": 6, "rho_updates": 7, "last_admm": 8, "last_rho": 9, "checks": 10, "step_norm": 11}
INFO = {"status": 0, "iters": 1, "pcg_iters": 2, "prim_res": 3, "dual_res": 4, "rho": 5, "rho_updates": 6, "checks": 7}
QP_SOLVED, QP_SOLVED_INACCURATE, QP_PRIMAL_INFEASIBLE, QP_DUAL_INFEASIBLE, QP_MAX_ITER, QP_UNSOLVED = 1, 2, 3, 5, 7, 11
PRECOND_DIAGONAL, PRECOND_BLOCK_JACOBI, PRECOND_BLOCK_TRIDIAG = 0, 1, 2
ERR_NO_DEVICE = 3


class Settings(C.Structure):
    """``ocp_b200_settings`` (include/ocp_b200.h)."""
    _fields_ = [
        ("sqp_alpha", C.c_double), ("sqp_step_num", C.c_int),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("admm_max_iter", C.c_int), ("rho", C.c_double), ("sigma", C.c_double), ("relax", C.c_double),
        ("scaling_iters", C.c_int), ("check_termination", C.c_int), ("adaptive_rho", C.c_int),
        ("adaptive_rho_interval", C.c_int), ("adaptive_rho_tolerance", C.c_double),
        ("pcg_max_iter", C.c_int), ("pcg_tol", C.c_double), ("pcg_precond", C.c_int),
    ]

    def copy(self) -> "Settings":
        out = Settings()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(Settings))
        return out


class ProblemDesc(C.Structure):
    """``ocp_b200_problem_desc`` (include/ocp_b200.h)."""
    _fields_ = [
        ("np", C.c_int), ("nf", C.c_int), ("horizon", C.c_int), ("ng", C.c_int),
        ("nnz_h", C.c_int), ("h_colptr", C.POINTER(C.c_int)), ("h_rowidx", C.POINTER(C.c_int)),
        ("nnz_a", C.c_int), ("a_colptr", C.POINTER(C.c_int)), ("a_rowidx", C.POINTER(C.c_int)),
        ("model_library", C.c_char_p),
        ("num_blocks", C.c_int), ("block_ptr", C.POINTER(C.c_int)),
        ("device", C.c_int),
    ]


class NativeLibraryMissing(RuntimeError):
    pass


_cuda = None
_host = None


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a, shape=None):
    if a is None:
        return None
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and out.size != int(np.prod(shape)):
        raise ValueError(f"expected {int(np.prod(shape))} values, got {out.size}")
    return out


def cuda_lib() -> C.CDLL:
    """``libocp_b200.so`` -- raises (never falls back) when it has not been built."""
    global _cuda
    if _cuda is None:
        # OCP_B200_LIB_DIR: another build of the two libraries (A/B measurements of kernel variants in one GPU session)
        path = Path(os.environ.get("OCP_B200_LIB_DIR", _PKG / "lib")) / "libocp_b200.so"
        if not path.exists():
            raise NativeLibraryMissing(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
        lib.ocp_b200_last_error.restype = C.c_char_p
        lib.ocp_b200_launch_count.restype = C.c_longlong
        lib.ocp_b200_launch_count.argtypes = [C.c_void_p]
        dptr, vptr = C.POINTER(C.c_double), C.c_void_p
        lib.ocp_b200_default_settings.argtypes = [C.POINTER(Settings)]
        lib.ocp_b200_default_settings.restype = None
        lib.ocp_b200_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(Settings), C.POINTER(vptr)]
        lib.ocp_b200_destroy.argtypes = [vptr]
        lib.ocp_b200_update_settings.argtypes = [vptr, C.POINTER(Settings)]
        lib.ocp_b200_get_settings.argtypes = [vptr, C.POINTER(Settings)]
        lib.ocp_b200_solve_batch.argtypes = [vptr, C.c_int] + [dptr] * 9
        lib.ocp_b200_solve_batch_device.argtypes = [vptr, C.c_int] + [vptr] * 9 + [vptr]
        lib.ocp_b200_shift_iterate_device.argtypes = [vptr, C.c_int, vptr, vptr]
        lib.ocp_b200_export_qp.argtypes = [vptr, C.c_int] + [dptr] * 12
        lib.ocp_b200_qp_solve_batch.argtypes = [vptr, C.c_int] + [dptr] * 8
        lib.ocp_b200_admm_trace.argtypes = [vptr] + [dptr] * 5 + [C.c_int, dptr, C.POINTER(C.c_int), dptr, dptr]
        lib.ocp_b200_get_dims.argtypes = [vptr] + [C.POINTER(C.c_int)] * 6
        lib.ocp_b200_get_plan.argtypes = [vptr, C.POINTER(C.c_int), C.c_int]
        lib.ocp_b200_create_multi.argtypes = [C.POINTER(ProblemDesc), C.POINTER(Settings), C.POINTER(C.c_int), C.c_int, C.POINTER(vptr)]
        lib.ocp_b200_destroy_multi.argtypes = [vptr]
        lib.ocp_b200_multi_update_settings.argtypes = [vptr, C.POINTER(Settings)]
        lib.ocp_b200_multi_partition.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.ocp_b200_multi_device_count.argtypes = [vptr]
        lib.ocp_b200_multi_handle.argtypes = [vptr, C.c_int]
        lib.ocp_b200_multi_handle.restype = vptr
        lib.ocp_b200_solve_batch_multi.argtypes = [vptr, C.c_int] + [dptr] * 9
        lib.ocp_b200_set_profiling.argtypes = [vptr, C.c_int]
        lib.ocp_b200_get_profile.argtypes = [vptr, dptr, C.POINTER(C.c_longlong), C.c_int]
        lib.ocp_b200_get_phase_cycles.argtypes = [vptr, C.POINTER(C.c_longlong)]
        _cuda = lib
    return _cuda


def host_lib() -> C.CDLL:
    """``libocp_b200_host.so`` (C++ front-end)."""
    global _host
    if _host is None:
        cuda_lib()
        path = Path(os.environ.get("OCP_B200_LIB_DIR", _PKG / "lib")) / "libocp_b200_host.so"
        if not path.exists():
            raise NativeLibraryMissing(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(str(path))
        lib.ocp_host_last_error.restype = C.c_char_p
        lib.ocp_host_problem_model_library.restype = C.c_char_p
        lib.ocp_host_problem_model_library.argtypes = [C.c_void_p]
        vptr, dptr, iptr = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        lib.ocp_host_set_share_dir.argtypes = [C.c_char_p]
        lib.ocp_host_problem_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.POINTER(vptr)]
        lib.ocp_host_problem_create_yaml.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vptr)]
        lib.ocp_host_problem_destroy.argtypes = [vptr]
        lib.ocp_host_problem_dims.argtypes = [vptr, iptr]
        lib.ocp_host_problem_patterns.argtypes = [vptr, iptr, iptr, iptr, iptr]
        lib.ocp_host_problem_bounds.argtypes = [vptr, dptr, dptr, dptr, dptr]
        lib.ocp_host_problem_handle.argtypes = [vptr, C.POINTER(vptr)]
        lib.ocp_host_problem_get_settings.argtypes = [vptr, C.POINTER(Settings)]
        lib.ocp_host_problem_set_settings.argtypes = [vptr, C.POINTER(Settings)]
        lib.ocp_host_problem_set_schedule.argtypes = [vptr, C.c_int, C.c_double]
        lib.ocp_host_compute_optimal_trajectory.argtypes = [vptr, dptr, dptr, dptr, dptr]
        lib.ocp_host_compute_optimal_trajectory_batch.argtypes = [vptr, C.c_int, dptr, dptr, dptr, dptr, dptr]
        lib.ocp_host_problem_reset.argtypes = [vptr]
        lib.ocp_host_problem_shift_batch.argtypes = [vptr]
        lib.ocp_host_problem_set_devices.argtypes = [vptr, C.POINTER(C.c_int), C.c_int]
        lib.ocp_host_problem_generate_c.argtypes = [vptr, C.c_char_p]
        lib.ocp_host_compile_casadi_c.argtypes = [C.c_char_p] * 4 + [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        pv = C.POINTER(vptr)
        lib.ocp_host_sx_sym.argtypes = [C.c_char_p, C.c_int, pv]
        lib.ocp_host_sx_const.argtypes = [dptr, C.c_int, pv]
        lib.ocp_host_sx_unary.argtypes = [C.c_char_p, vptr, pv]
        lib.ocp_host_sx_binary.argtypes = [C.c_char_p, vptr, vptr, pv]
        lib.ocp_host_sx_vertcat.argtypes = [pv, C.c_int, pv]
        lib.ocp_host_sx_slice.argtypes = [vptr, C.c_int, C.c_int, pv]
        lib.ocp_host_sx_size.argtypes = [vptr]
        lib.ocp_host_sx_free.argtypes = [vptr]
        lib.ocp_host_scripted_create.argtypes = [C.c_char_p, C.c_char_p, pv]
        lib.ocp_host_scripted_info.argtypes = [vptr, iptr, dptr, iptr]
        lib.ocp_host_scripted_variable.argtypes = [vptr, C.c_int, C.c_char_p, pv]
        lib.ocp_host_scripted_set_reference.argtypes = [vptr, vptr]
        lib.ocp_host_scripted_add_scalar_cost.argtypes = [vptr, vptr]
        lib.ocp_host_scripted_add_vector_cost.argtypes = [vptr, dptr, C.c_int, vptr]
        lib.ocp_host_scripted_add_inequality.argtypes = [vptr, C.c_char_p, dptr, vptr, dptr, C.c_int]
        lib.ocp_host_scripted_add_equation.argtypes = [vptr, C.c_char_p, vptr, vptr]
        lib.ocp_host_scripted_gen_solver.argtypes = [vptr]
        lib.ocp_host_default_yaml.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_char_p, C.c_int]
        lib.ocp_host_sample_inputs.argtypes = [C.c_char_p, C.c_int, C.c_ulonglong, dptr, dptr]
        lib.ocp_host_kat_create.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(vptr)]
        lib.ocp_host_kat_destroy.argtypes = [vptr]
        lib.ocp_host_kat_solve.argtypes = [vptr, dptr, iptr, dptr]
        lib.ocp_host_kat_set_settings.argtypes = [vptr, C.POINTER(Settings)]
        lib.ocp_host_cucaqp_solve.argtypes = [C.c_int, C.c_int, iptr, iptr, dptr, dptr, iptr, iptr, dptr, dptr, dptr,
                                              C.c_double, C.c_double, C.c_int, dptr, dptr, dptr]
        SHARE_DIR.mkdir(parents=True, exist_ok=True)
        lib.ocp_host_set_share_dir(str(SHARE_DIR).encode())
        os.environ.setdefault("OCP_B200_INCLUDE_DIR", str(ROOT / "include"))
        _host = lib
    return _host


class OcpB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ocp_b200 error {code}: {message}")
        self.code = code


def _check(rc: int) -> None:
    if rc != 0:
        raise OcpB200Error(rc, cuda_lib().ocp_b200_last_error().decode())


def _hcheck(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(host_lib().ocp_host_last_error().decode())


def multi_partition(B: int, ndev: int) -> np.ndarray:
    """``ocp_b200_multi_partition``: instance offsets of the contiguous block partition over ``ndev`` devices."""
    off = np.zeros(ndev + 1, np.int32)
    _check(cuda_lib().ocp_b200_multi_partition(int(B), int(ndev), _ip(off)))
    return off


class MultiSolver:
    """``ocp_b200_create_multi`` / ``ocp_b200_solve_batch_multi``: one process, several GPUs (or several handles on one)."""

    def __init__(self, problem: "Problem", devices, settings: Settings | None = None):
        hc, hr = np.ascontiguousarray(problem.h_colptr, np.int32), np.ascontiguousarray(problem.h_rowidx, np.int32)
        ac, ar = np.ascontiguousarray(problem.a_colptr, np.int32), np.ascontiguousarray(problem.a_rowidx, np.int32)
        d = ProblemDesc()
        d.np, d.nf, d.horizon, d.ng = problem.np_, problem.nf, problem.horizon, problem.ng
        d.nnz_h, d.h_colptr, d.h_rowidx = hr.size, _ip(hc), _ip(hr)
        d.nnz_a, d.a_colptr, d.a_rowidx = ar.size, _ip(ac), _ip(ar)
        d.model_library = problem.model_library.encode()
        dev = np.ascontiguousarray(list(devices), np.int32)
        s = settings if settings is not None else problem.get_settings()
        out = C.c_void_p()
        _check(cuda_lib().ocp_b200_create_multi(C.byref(d), C.byref(s), _ip(dev), int(dev.size), C.byref(out)))
        self._h, self.problem, self.devices = out, problem, dev.tolist()

    def update_settings(self, s: Settings) -> None:
        _check(cuda_lib().ocp_b200_multi_update_settings(self._h, C.byref(s)))

    def solve_batch(self, frames, p, x_inout, f_out=None, stats=None):
        pr = self.problem
        B = x_inout.size // pr.N
        frames = _f64(frames, (B, pr.nf)) if frames is not None else None
        _check(cuda_lib().ocp_b200_solve_batch_multi(self._h, B, _dp(frames), _dp(_f64(p, (B, pr.np_))), _dp(pr.lbx), _dp(pr.ubx),
                                                    _dp(pr.lbg), _dp(pr.ubg), _dp(x_inout), _dp(f_out), _dp(stats)))

    def close(self) -> None:
        if self._h:
            cuda_lib().ocp_b200_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compile_casadi_c(c_file: str, nf: int, horizon: int, name: str = "model", local_system_fn: str = "localSystemFunction",
                     objective_fn: str = "objective", code_dir: str | None = None) -> str:
    """A CasADi-generated C file (``Function::generate`` layout) -> stage library for ``Solver.create(model_library=...)``
    (host/src/CasadiCInterop.cpp).  Needs nvcc and a C compiler, no GPU."""
    code_dir = code_dir or str(_PKG / "share" / "code_gen")
    out = C.create_string_buffer(4096)
    _hcheck(host_lib().ocp_host_compile_casadi_c(str(c_file).encode(), local_system_fn.encode(), objective_fn.encode(), name.encode(),
                                                 int(nf), int(horizon), code_dir.encode(), out, 4096))
    return out.value.decode()


def default_settings() -> Settings:
    s = Settings()
    cuda_lib().ocp_b200_default_settings(C.byref(s))
    return s


class Solver:
    """The C ABI of include/ocp_b200.h on numpy buffers.  ``handle`` may be borrowed from a
    :class:`Problem` (then it is not destroyed here) or created from CCS patterns."""

    def __init__(self, handle: int, dims: dict, owned: bool):
        self._h = C.c_void_p(handle)
        self._owned = owned
        self.np_, self.nf, self.horizon, self.ng = dims["np"], dims["nf"], dims["horizon"], dims["ng"]
        self.n, self.m, self.nnz_h, self.nnz_a = dims["n"], dims["m"], dims["nnz_h"], dims["nnz_a"]
        self.N = self.nf * self.horizon

    @classmethod
    def create(cls, n: int, m: int, h_colptr, h_rowidx, a_colptr, a_rowidx, settings: Settings | None = None,
               np_: int = 0, nf: int | None = None, horizon: int = 1, model_library: str | None = None,
               block_ptr=None, device: int = 0) -> "Solver":
        hc = np.ascontiguousarray(h_colptr, dtype=np.int32)
        hr = np.ascontiguousarray(h_rowidx, dtype=np.int32)
        ac = np.ascontiguousarray(a_colptr, dtype=np.int32)
        ar = np.ascontiguousarray(a_rowidx, dtype=np.int32)
        nf = (n - np_) // horizon if nf is None else nf
        d = ProblemDesc()
        d.np, d.nf, d.horizon, d.ng = np_, nf, horizon, m - n
        d.nnz_h, d.h_colptr, d.h_rowidx = hr.size, _ip(hc), _ip(hr)
        d.nnz_a, d.a_colptr, d.a_rowidx = ar.size, _ip(ac), _ip(ar)
        d.model_library = model_library.encode() if model_library else None
        bp = None
        if block_ptr is not None:
            bp = np.ascontiguousarray(block_ptr, dtype=np.int32)
            d.num_blocks, d.block_ptr = bp.size - 1, _ip(bp)
        d.device = device
        out = C.c_void_p()
        _check(cuda_lib().ocp_b200_create(C.byref(d), C.byref(settings) if settings is not None else None,
                                          C.byref(out)))
        dims = dict(np=np_, nf=nf, horizon=horizon, ng=m - n, n=n, m=m, nnz_h=int(hr.size), nnz_a=int(ar.size))
        return cls(out.value, dims, owned=True)

    def close(self) -> None:
        if self._owned and self._h:
            cuda_lib().ocp_b200_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def get_settings(self) -> Settings:
        s = Settings()
        _check(cuda_lib().ocp_b200_get_settings(self._h, C.byref(s)))
        return s

    def update_settings(self, s: Settings) -> None:
        _check(cuda_lib().ocp_b200_update_settings(self._h, C.byref(s)))

    def launch_count(self) -> int:
        return int(cuda_lib().ocp_b200_launch_count(self._h))

    def set_profiling(self, enabled, phases: bool = False) -> None:
        """enabled: CUDA-event pairs around every launch; phases: also the per-phase cycle counters of CTA 0."""
        _check(cuda_lib().ocp_b200_set_profiling(self._h, 2 if (enabled and phases) else int(bool(enabled))))

    def get_profile(self, reset: bool = True) -> dict:
        """Accumulated device milliseconds / launch counts per kernel kind since the last reset."""
        ms = (C.c_double * 3)(); cnt = (C.c_longlong * 3)()
        _check(cuda_lib().ocp_b200_get_profile(self._h, ms, cnt, int(reset)))
        return {k: dict(ms=ms[i], launches=cnt[i]) for i, k in enumerate(("admm", "assemble", "objective"))}

    def get_phase_cycles(self) -> dict:
        """SM cycles per QP phase accumulated by CTA 0 of the direct kernel while profiling was on."""
        c = (C.c_longlong * 14)()
        _check(cuda_lib().ocp_b200_get_phase_cycles(self._h, c))
        return dict(zip(("load", "scale", "kkt_assemble", "factor", "rhs", "solve", "update", "check",
                         "solve_fwd", "solve_border", "solve_diag", "solve_bwd", "factor_invert", "factor_step"), list(c)))

    def device_dims(self) -> dict:
        v = [C.c_int() for _ in range(6)]
        _check(cuda_lib().ocp_b200_get_dims(self._h, *[C.byref(x) for x in v]))
        return dict(n=v[0].value, m=v[1].value, nnz_h=v[2].value, nnz_a=v[3].value, smem_bytes=v[4].value,
                    resident=v[5].value)

    def launch_plan(self) -> dict:
        """ocp_b200_get_plan: the two launch plans and the block structure of K (see OCP_B200_PLAN_*)."""
        v = (C.c_int * 16)()
        _check(cuda_lib().ocp_b200_get_plan(self._h, v, 16))
        keys = ("place", "threads", "smem_bytes", "ctas_per_sm", "slab_kb")
        return dict(wide=dict(zip(keys, v[0:5])), deep=dict(zip(keys, v[5:10])), tri_ok=v[10], tri_np=v[11], tri_bs=v[12],
                    tri_nb=v[13], nnz_p=v[14], num_sms=v[15])

    def solve_batch(self, frames, p, lbx, ubx, lbg, ubg, x_inout, f_out=None, stats=None):
        """ocp_b200_solve_batch: host buffers, x_inout [B, N] updated in place."""
        if x_inout.dtype != np.float64 or not x_inout.flags.c_contiguous:
            raise ValueError("x_inout must be a C-contiguous float64 array")
        B = x_inout.size // max(self.N, 1)
        frames = _f64(frames, (B, self.nf)) if frames is not None else None
        p = _f64(p, (B, self.np_))
        lbx, ubx = _f64(lbx, (self.N,)), _f64(ubx, (self.N,))
        lbg, ubg = _f64(lbg, (self.ng,)), _f64(ubg, (self.ng,))
        _check(cuda_lib().ocp_b200_solve_batch(self._h, B, _dp(frames), _dp(p), _dp(lbx), _dp(ubx), _dp(lbg),
                                               _dp(ubg), _dp(x_inout), _dp(f_out), _dp(stats)))
        return x_inout

    def solve_batch_device(self, B, d_frames, d_p, d_lbx, d_ubx, d_lbg, d_ubg, d_x, d_f, d_stats, stream=0):
        """ocp_b200_solve_batch_device: raw device pointers (ints), asynchronous on ``stream``."""
        _check(cuda_lib().ocp_b200_solve_batch_device(self._h, B, d_frames, d_p, d_lbx, d_ubx, d_lbg, d_ubg, d_x,
                                                      d_f, d_stats, stream))

    def shift_iterate_device(self, B, d_x, stream=0):
        """ocp_b200_shift_iterate_device: frame k <- frame k+1 in place on the device iterate."""
        _check(cuda_lib().ocp_b200_shift_iterate_device(self._h, B, d_x, stream))

    def export_qp(self, frames, p, lbx, ubx, lbg, ubg, x):
        """ocp_b200_export_qp: local system (H values, q, A values, l, u) of B instances."""
        x = _f64(x)
        B = x.size // self.N
        frames = _f64(frames, (B, self.nf)) if frames is not None else None
        p = _f64(p, (B, self.np_))
        hv = np.empty((B, self.nnz_h)); q = np.empty((B, self.n)); av = np.empty((B, self.nnz_a))
        l = np.empty((B, self.m)); u = np.empty((B, self.m))
        _check(cuda_lib().ocp_b200_export_qp(self._h, B, _dp(frames), _dp(p), _dp(_f64(lbx)), _dp(_f64(ubx)),
                                             _dp(_f64(lbg)), _dp(_f64(ubg)), _dp(x), _dp(hv), _dp(q), _dp(av),
                                             _dp(l), _dp(u)))
        return hv, q, av, l, u

    def qp_solve_batch(self, h_vals, q, a_vals, l, u):
        """ocp_b200_qp_solve_batch: B cold-started QPs sharing the create-time patterns."""
        q = _f64(q)
        B = q.size // self.n
        x = np.empty((B, self.n)); y = np.empty((B, self.m)); info = np.empty((B, NINFO))
        _check(cuda_lib().ocp_b200_qp_solve_batch(self._h, B, _dp(_f64(h_vals, (B, self.nnz_h))), _dp(q),
                                                  _dp(_f64(a_vals, (B, self.nnz_a))), _dp(_f64(l, (B, self.m))),
                                                  _dp(_f64(u, (B, self.m))), _dp(x), _dp(y), _dp(info)))
        return x, y, info

    def admm_trace(self, h_vals, q, a_vals, l, u, max_records: int = 512):
        trace = np.zeros((max_records, TRACE_WIDTH)); nrec = C.c_int(0)
        x = np.empty(self.n); y = np.empty(self.m)
        _check(cuda_lib().ocp_b200_admm_trace(self._h, _dp(_f64(h_vals)), _dp(_f64(q)), _dp(_f64(a_vals)),
                                              _dp(_f64(l)), _dp(_f64(u)), max_records, _dp(trace), C.byref(nrec),
                                              _dp(x), _dp(y)))
        return trace[: nrec.value], x, y


class Problem:
    """A concrete ``OptimalControlProblem`` (quadrotor / cartpole / centroidal, the shapes
    BASELINE.json names) built through the C++ front-end: YAML -> OCPConfig -> cost/constraint
    registration -> ``genSolver()`` (symbolic AD, CUDA stage code generation, nvcc)."""

    def __init__(self, name: str, horizon: int = 0, alpha: float = 0.1, step_num: int = 10, yaml_text: str | None = None):
        lib = host_lib()
        self.name = name
        out = C.c_void_p()
        if yaml_text is not None:
            _hcheck(lib.ocp_host_problem_create_yaml(name.encode(), yaml_text.encode(), C.byref(out)))
        else:
            _hcheck(lib.ocp_host_problem_create(name.encode(), horizon, alpha, step_num, C.byref(out)))
        self._h = out
        self._load_description()

    @classmethod
    def _adopt(cls, name: str, handle) -> "Problem":
        """Wraps a host problem that already went through genSolver() (symbolic.OptimalControlProblem)."""
        self = cls.__new__(cls)
        self.name, self._h = name, handle
        self._load_description()
        return self

    def _load_description(self) -> None:
        lib = host_lib()
        dims = (C.c_int * 8)()
        _hcheck(lib.ocp_host_problem_dims(self._h, dims))
        self.np_, self.nf, self.horizon, self.ng, self.n, self.m, self.nnz_h, self.nnz_a = list(dims)
        self.N = self.nf * self.horizon
        self.h_colptr = np.empty(self.n + 1, np.int32); self.h_rowidx = np.empty(self.nnz_h, np.int32)
        self.a_colptr = np.empty(self.n + 1, np.int32); self.a_rowidx = np.empty(self.nnz_a, np.int32)
        _hcheck(lib.ocp_host_problem_patterns(self._h, _ip(self.h_colptr), _ip(self.h_rowidx), _ip(self.a_colptr),
                                              _ip(self.a_rowidx)))
        self.lbx = np.empty(self.N); self.ubx = np.empty(self.N); self.lbg = np.empty(self.ng); self.ubg = np.empty(self.ng)
        _hcheck(lib.ocp_host_problem_bounds(self._h, _dp(self.lbx), _dp(self.ubx), _dp(self.lbg), _dp(self.ubg)))
        self.model_library = lib.ocp_host_problem_model_library(self._h).decode()
        self._solver = None

    def __del__(self):
        try:
            if self._h:
                host_lib().ocp_host_problem_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def dims(self) -> dict:
        return dict(np=self.np_, nf=self.nf, horizon=self.horizon, ng=self.ng, n=self.n, m=self.m,
                    nnz_h=self.nnz_h, nnz_a=self.nnz_a)

    @property
    def solver(self) -> Solver:
        """The device handle behind this problem (created on first use; needs a B200)."""
        if self._solver is None:
            out = C.c_void_p()
            _hcheck(host_lib().ocp_host_problem_handle(self._h, C.byref(out)))
            self._solver = Solver(out.value, self.dims, owned=False)
        return self._solver

    def get_settings(self) -> Settings:
        s = Settings()
        _hcheck(host_lib().ocp_host_problem_get_settings(self._h, C.byref(s)))
        return s

    def set_settings(self, s: Settings) -> None:
        _hcheck(host_lib().ocp_host_problem_set_settings(self._h, C.byref(s)))

    def set_schedule(self, step_num: int, alpha: float) -> None:
        _hcheck(host_lib().ocp_host_problem_set_schedule(self._h, step_num, alpha))

    def sample_inputs(self, B: int, seed: int):
        frames = np.empty((B, self.nf)); refs = np.empty((B, self.np_))
        _hcheck(host_lib().ocp_host_sample_inputs(self.name.encode(), B, seed, _dp(frames), _dp(refs)))
        return frames, refs

    def compute_optimal_trajectory(self, frame, reference):
        """``OptimalControlProblem::computeOptimalTrajectory`` -> (x [N], f)."""
        x = np.empty(self.N); f = C.c_double(0.0)
        _hcheck(host_lib().ocp_host_compute_optimal_trajectory(self._h, _dp(_f64(frame, (self.nf,))),
                                                               _dp(_f64(reference, (self.np_,))), _dp(x), C.byref(f)))
        return x, f.value

    def compute_optimal_trajectory_batch(self, frames, references):
        """Batched sibling -> (x [B, N], f [B], stats [B, NSTATS]); warm-starts from the last call."""
        frames = _f64(frames)
        B = frames.size // self.nf
        x = np.empty((B, self.N)); f = np.empty(B); st = np.empty((B, NSTATS))
        _hcheck(host_lib().ocp_host_compute_optimal_trajectory_batch(self._h, B, _dp(frames),
                                                                     _dp(_f64(references, (B, self.np_))), _dp(x),
                                                                     _dp(f), _dp(st)))
        return x, f, st

    def generate_c(self, path: str) -> str:
        """Writes localSystemFunction + objective as one C file in the layout of CasADi's code generator."""
        _hcheck(host_lib().ocp_host_problem_generate_c(self._h, str(path).encode()))
        return str(path)

    def set_devices(self, devices) -> None:
        """``SQPOptimizationSolver::setDevices``: the GPUs ``compute_optimal_trajectory_batch`` spreads a batch over
        (contiguous blocks, ``ocp_b200_create_multi``); a device may be listed twice."""
        d = np.ascontiguousarray(list(devices), dtype=np.int32)
        _hcheck(host_lib().ocp_host_problem_set_devices(self._h, _ip(d), int(d.size)))

    def reset(self) -> None:
        _hcheck(host_lib().ocp_host_problem_reset(self._h))

    def shift_batch_trajectory(self) -> None:
        """``OptimalControlProblem::shiftBatchTrajectory``: receding-horizon shift of the stored batch iterates."""
        _hcheck(host_lib().ocp_host_problem_shift_batch(self._h))


class MpcLoop:
    """Device-resident receding-horizon loop over B instances of one ``Problem`` (SURVEY.md 8f items
    1 and 3): torch CUDA tensors in and out, no host copies between ticks.

    The iterate ``x`` [B, N] lives on the device across ticks, which is what the reference's
    persistent ``result_["x"]`` (SQPOptimizationSolver.cpp:215) amounts to for one instance; with
    ``shift=True`` a tick first moves every stage one step towards the start of the horizon."""

    def __init__(self, problem: "Problem", batch: int, device=None):
        import torch
        if not torch.cuda.is_available():
            raise NativeLibraryMissing("MpcLoop needs a CUDA device: there is no CPU path")
        self.torch = torch
        self.problem, self.B = problem, int(batch)
        self.solver = problem.solver
        self.device = torch.device(device if device is not None else "cuda:0")
        # the solver handle of a Problem lives on device 0 (SQPOptimizationSolver's default): tensors and the
        # stream of another GPU must not be handed to its kernels
        if self.device.type != "cuda" or (self.device.index or 0) != 0:
            raise ValueError(f"MpcLoop: the solver of this Problem is on cuda:0, got device {self.device}")
        f64 = dict(dtype=torch.float64, device=self.device)
        self.x = torch.zeros((self.B, problem.N), **f64)
        self.f = torch.zeros(self.B, **f64)
        self.stats = torch.zeros((self.B, NSTATS), **f64)
        self._bounds = [torch.from_numpy(np.ascontiguousarray(b)).to(self.device)
                        for b in (problem.lbx, problem.ubx, problem.lbg, problem.ubg)]
        self.ticks = 0

    def _ptr(self, t, shape):
        torch = self.torch
        if t is None:
            return None
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
                and t.device == self.device and tuple(t.shape) == tuple(shape)):
            raise ValueError(f"expected a contiguous float64 CUDA tensor of shape {tuple(shape)} on {self.device}")
        return C.c_void_p(t.data_ptr())

    def reset(self) -> None:
        self.x.zero_()
        self.ticks = 0

    def tick(self, frames, references, shift: bool = False):
        """One MPC tick: (optionally) shift the iterate, then ``step_num`` SQP steps from it with the
        first frame of every instance pinned to ``frames`` [B, nf].  Asynchronous on torch's current
        stream; returns (x [B, N], f [B], stats [B, NSTATS]) -- views of buffers the next tick overwrites."""
        pr = self.problem
        stream = C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)
        d_fr = self._ptr(frames, (self.B, pr.nf))
        d_p = self._ptr(references, (self.B, pr.np_))
        if shift and self.ticks > 0:
            self.solver.shift_iterate_device(self.B, C.c_void_p(self.x.data_ptr()), stream)
        lbx, ubx, lbg, ubg = (C.c_void_p(b.data_ptr()) for b in self._bounds)
        self.solver.solve_batch_device(self.B, d_fr, d_p, lbx, ubx, lbg, ubg, C.c_void_p(self.x.data_ptr()),
                                       C.c_void_p(self.f.data_ptr()), C.c_void_p(self.stats.data_ptr()), stream)
        self.ticks += 1
        return self.x, self.f, self.stats

    def first_frame(self):
        """Stage 0 of every trajectory [B, nf] (state and the control to apply): the reference's
        ``getOptimalInputFirstFrame`` for the batch."""
        return self.x[:, : self.problem.nf]

    def frame(self, k: int):
        nf = self.problem.nf
        return self.x[:, k * nf:(k + 1) * nf]


class KatProblem:
    """NLP number ``case`` of the reference's test/test.cpp through SQPOptimizationSolver."""

    def __init__(self, case: int, step_num: int = 1, alpha: float = 1.0):
        out = C.c_void_p()
        _hcheck(host_lib().ocp_host_kat_create(case, step_num, alpha, C.byref(out)))
        self._h = out

    def set_settings(self, s: Settings) -> None:
        _hcheck(host_lib().ocp_host_kat_set_settings(self._h, C.byref(s)))

    def solve(self):
        x = np.zeros(8); n = C.c_int(0); f = C.c_double(0.0)
        _hcheck(host_lib().ocp_host_kat_solve(self._h, _dp(x), C.byref(n), C.byref(f)))
        return x[: n.value].copy(), f.value

    def __del__(self):
        try:
            if self._h:
                host_lib().ocp_host_kat_destroy(self._h)
                self._h = None
        except Exception:
            pass


def cucaqp_solve(n, m, h_colptr, h_rowidx, h_vals, q, a_colptr, a_rowidx, a_vals, l, u, eps_abs=1e-3, eps_rel=1e-3,
                 max_iter=10000):
    """One QP through the ``CuCaQP`` class life cycle (setDimension/setSystem/initSolver/solve)."""
    x = np.empty(n); y = np.empty(m); info = np.empty(NINFO)
    hc = np.ascontiguousarray(h_colptr, np.int32); hr = np.ascontiguousarray(h_rowidx, np.int32)
    ac = np.ascontiguousarray(a_colptr, np.int32); ar = np.ascontiguousarray(a_rowidx, np.int32)
    _hcheck(host_lib().ocp_host_cucaqp_solve(n, m, _ip(hc), _ip(hr), _dp(_f64(h_vals)), _dp(_f64(q)), _ip(ac), _ip(ar),
                                             _dp(_f64(a_vals)), _dp(_f64(l)), _dp(_f64(u)), eps_abs, eps_rel, max_iter,
                                             _dp(x), _dp(y), _dp(info)))
    return x, y, info
