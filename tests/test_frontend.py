"""CPU tests of the host side: the reference class surface (YAML -> OCPConfig ->
OptimalControlProblem -> SQPOptimizationSolver set-up), the structural sparsity patterns, the
stage code generator, the C ABI's symbol table, and an independent numpy restatement of one
benchmark problem against the symbolic layer.  No test here computes on a GPU."""
import ctypes as C
import hashlib
import json
import re
from pathlib import Path

import numpy as np
import pytest

import _oracle

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---------------------------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol(native):
    lib = native.cuda_lib()
    header = (ROOT / "include" / "ocp_b200.h").read_text()
    names = set(re.findall(r"\b(ocp_b200_[a-z_0-9]+)\s*\(", header))
    assert len(names) >= 16
    for name in sorted(names):
        assert hasattr(lib, name), f"{name} is declared in include/ocp_b200.h but not exported"
    assert lib.ocp_b200_abi_version() == 1


def test_default_settings_are_the_reference_values(native):
    s = native.default_settings()
    # SQPOptimizationSolver.cpp:83-85 and OptimalControlProblem.h:24-27; OSQP v1.0.0.beta1 defaults
    assert (s.eps_abs, s.eps_rel, s.admm_max_iter) == (1e-3, 1e-3, 10000)
    assert (s.sqp_alpha, s.sqp_step_num) == (0.1, 10)
    assert (s.rho, s.sigma, s.relax, s.scaling_iters, s.check_termination) == (0.1, 1e-6, 1.6, 10, 25)
    assert (s.adaptive_rho, s.adaptive_rho_interval, s.adaptive_rho_tolerance) == (1, 0, 5.0)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a machine without a CUDA device")
def test_no_device_is_an_error_not_a_fallback(native):
    """There is no CPU path behind the ABI: creating a solver without a GPU fails loudly."""
    with pytest.raises(native.OcpB200Error) as e:
        native.Solver.create(1, 2, [0, 1], [0], [0, 2], [0, 1])
    assert e.value.code == native.ERR_NO_DEVICE
    prob = native.Problem("quadrotor")
    with pytest.raises(RuntimeError, match="no CUDA device|ocp_b200_create"):
        prob.compute_optimal_trajectory(np.zeros(prob.nf), np.zeros(prob.np_))


def test_problem_too_large_for_16_bit_indices_is_unsupported(native):
    """Maximum size: index structures are 16-bit, n + ng must stay below 65535 (checked before any device work)."""
    n = 70000
    hc = np.arange(n + 1, dtype=np.int32); hr = np.arange(n, dtype=np.int32)
    with pytest.raises(native.OcpB200Error) as e:
        native.Solver.create(n, n, hc, hr, hc, hr)
    assert e.value.code == 5   # OCP_B200_ERR_UNSUPPORTED


def test_invalid_descriptions_are_rejected(native):
    for args in [dict(n=2, m=3, hc=[0, 1, 1], hr=[0], ac=[0, 1, 3], ar=[0, 2, 1]),   # rows not increasing
                 dict(n=2, m=3, hc=[0, 1, 2], hr=[0, 1], ac=[0, 1], ar=[0])]:          # colptr too short -> nnz mismatch
        with pytest.raises(native.OcpB200Error) as e:
            native.Solver.create(args["n"], args["m"], args["hc"], args["hr"], args["ac"] + [0] * (3 - len(args["ac"])), args["ar"])
        assert e.value.code in (1, native.ERR_NO_DEVICE)


# ---------------------------------------------------------------------------------------------
# front-end: YAML, frame layout, registration, patterns
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["quadrotor", "cartpole", "centroidal"])
def test_patterns_match_oracle_and_golden(problems, name):
    prob, ora = problems(name)
    gold = json.loads((GOLDEN / "patterns.json").read_text())[name]
    dims = prob.dims
    for k in ("np", "nf", "horizon", "ng", "n", "m", "nnz_h", "nnz_a"):
        assert dims[k] == gold[k]
    for a in ("h_colptr", "h_rowidx", "a_colptr", "a_rowidx"):
        assert np.array_equal(getattr(prob, a), getattr(ora, a))
        assert hashlib.sha256(getattr(prob, a).astype(np.int32).tobytes()).hexdigest() == gold[a]


@pytest.mark.parametrize("name", ["quadrotor", "cartpole", "centroidal"])
def test_pattern_invariants(problems, name):
    """w = [p; x], c = [p; x; g] (SQPOptimizationSolver.cpp:50-54): the first n rows of J are the
    identity; CCS columns have strictly increasing rows; the Hessian pattern is symmetric."""
    prob, _ = problems(name)
    n, m = prob.n, prob.m
    assert n == prob.np_ + prob.nf * prob.horizon and m == n + prob.ng
    for ptr, idx, nrow in ((prob.a_colptr, prob.a_rowidx, m), (prob.h_colptr, prob.h_rowidx, n)):
        assert ptr[0] == 0 and ptr[-1] == idx.size and (np.diff(ptr) >= 0).all()
        for j in range(n):
            col = idx[ptr[j]:ptr[j + 1]]
            assert (np.diff(col) > 0).all() and (col < nrow).all()
    for j in range(n):
        col = prob.a_rowidx[prob.a_colptr[j]:prob.a_colptr[j + 1]]
        assert col[0] == j and (col[1:] >= n).all()
    H = {(int(i), j) for j in range(n) for i in prob.h_rowidx[prob.h_colptr[j]:prob.h_colptr[j + 1]]}
    assert all((j, i) in H for (i, j) in H)


def test_bounds_and_first_frame_layout(problems):
    prob, ora = problems("quadrotor")
    assert prob.lbx.shape == (320,) and np.array_equal(prob.lbx, ora.lbx) and np.array_equal(prob.ubx, ora.ubx)
    frame_lb = prob.lbx[:16]
    assert np.array_equal(prob.lbx.reshape(20, 16), np.tile(frame_lb, (20, 1)))   # OCPConfig.cpp:301-304
    assert np.isneginf(frame_lb[:3]).all() and frame_lb[3] == -0.8 and (prob.lbx.reshape(20, 16)[:, 12:] == 0).all()
    assert (prob.lbg == 0).all() and (prob.ubg == 0).all()                        # equality constraints: [0, 0]


YAML_README = """
optimal_control_problem:
  discretization_settings:
    dt: 0.01
    horizon: 6
  solver_settings:
    max_iter: 1000
    warm_start: true
    verbose: false
    gen_code: false
    load_lib: false
    solve_method: CUDA_SQP
    SQP_step: 0.25
    ADMM_step: 4
    SQP_settings:
      alpha: 0.1
      step_num: 10
  OCP_variables:
    - name: state
      size: 4
      lower_bound: [-2.4, -.inf, -.inf, -.inf]
      upper_bound: [2.4, .inf, .inf, .inf]
    - name: force
      size: 1
      lower_bound: [-20.0]
      upper_bound: [20.0]
"""


def test_yaml_constructor_and_readme_aliases(native):
    """readme.md:55-62 spells the SQP settings SQP_step / ADMM_step; they alias
    SQP_settings.alpha / step_num (OptimalControlProblem.cpp:23-32)."""
    prob = native.Problem("cartpole", yaml_text=YAML_README)
    assert (prob.horizon, prob.nf, prob.N, prob.ng) == (6, 5, 30, 20)
    s = prob.get_settings()
    assert s.sqp_alpha == 0.25 and s.sqp_step_num == 4
    assert s.eps_abs == 1e-3 and s.eps_rel == 1e-3 and s.admm_max_iter == 10000   # SQPOptimizationSolver.cpp:83-85


def test_yaml_errors(native):
    with pytest.raises(RuntimeError, match="Invalid configuration"):
        native.Problem("cartpole", yaml_text=YAML_README.replace("    gen_code: false\n", ""))
    with pytest.raises(RuntimeError, match="not part of this build|nlpsol"):
        native.Problem("cartpole", yaml_text=YAML_README.replace("CUDA_SQP", "IPOPT"))
    with pytest.raises(RuntimeError, match="Unknown solver type"):
        native.Problem("cartpole", yaml_text=YAML_README.replace("CUDA_SQP", "NEWTON"))


# ---------------------------------------------------------------------------------------------
# symbolic layer vs finite differences and vs an independent numpy model
# ---------------------------------------------------------------------------------------------
def _cartpole_rk4(x, u, dt):
    M, mp, lp, g = 1.0, 0.1, 0.5, 9.81

    def f(s):
        th, ds, dth = s[1], s[2], s[3]
        sn, cs = np.sin(th), np.cos(th)
        tot = M + mp
        temp = (u + mp * lp * dth * dth * sn) / tot
        ddth = (g * sn - cs * temp) / (lp * (4.0 / 3.0 - mp * cs * cs / tot))
        dds = temp - mp * lp * ddth * cs / tot
        return np.array([ds, dth, dds, ddth])

    k1 = f(x); k2 = f(x + 0.5 * dt * k1); k3 = f(x + 0.5 * dt * k2); k4 = f(x + dt * k3)
    return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


def test_cartpole_matches_independent_numpy_model():
    """Objective and constraint values of the cart-pole OCP, written again in plain numpy."""
    H = 7
    ora = _oracle.OracleProblem("cartpole", horizon=H)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(ora.N) * 0.3
    p = rng.standard_normal(4) * 0.2
    X = x.reshape(H, 5)
    Q, R = np.array([1.0, 10.0, 0.1, 0.1]), 0.01
    f = sum((Q * (X[k, :4] - p) ** 2).sum() + R * X[k, 4] ** 2 for k in range(H))
    g = np.concatenate([X[k + 1, :4] - _cartpole_rk4(X[k, :4], X[k, 4], 0.01) for k in range(H - 1)])
    assert abs(ora.objective(p, x) - f) < 1e-12 * max(1.0, abs(f))
    hv, q, av, l, u = ora.local_system(None, p, x)
    c = np.concatenate([p, x, g])
    lfull = np.concatenate([p, ora.lbx, ora.lbg]); ufull = np.concatenate([p, ora.ubx, ora.ubg])
    fin = np.isfinite(lfull)
    assert np.abs(l[fin] - (lfull - c)[fin]).max() < 1e-12       # l' = l - c (SQPOptimizationSolver.cpp:66-71)
    fin = np.isfinite(ufull)
    assert np.abs(u[fin] - (ufull - c)[fin]).max() < 1e-12


def _dense(ptr, idx, vals, nrow, ncol):
    M = np.zeros((nrow, ncol))
    for j in range(ncol):
        M[idx[ptr[j]:ptr[j + 1]], j] = vals[ptr[j]:ptr[j + 1]]
    return M


@pytest.mark.parametrize("name,horizon", [("quadrotor", 4), ("centroidal", 3), ("cartpole", 6)])
def test_local_system_against_finite_differences(name, horizon):
    """grad f, hess f and dc/dw of the local system (AutoDifferentiator.cpp:16-27) vs central
    differences of the objective / constraint VALUES."""
    ora = _oracle.OracleProblem(name, horizon=horizon)
    frames, refs = ora.sample_inputs(1, 5)
    rng = np.random.default_rng(2)
    x = np.tile(frames[0], horizon) + 0.05 * rng.standard_normal(ora.N)
    p = refs[0] + 0.05 * rng.standard_normal(ora.np_)
    hv, q, av, l, u = ora.local_system(None, p, x)
    n, m, np_ = ora.n, ora.m, ora.np_
    J = _dense(ora.a_colptr, ora.a_rowidx, av, m, n)
    Hd = _dense(ora.h_colptr, ora.h_rowidx, hv, n, n)
    w0 = np.concatenate([p, x])

    def fval(w):
        return ora.objective(w[:np_], w[np_:])

    def cval(w):
        _, _, _, ll, _ = ora.local_system(None, w[:np_], w[np_:])
        lfull = np.concatenate([w[:np_], ora.lbx, ora.lbg])
        cc = lfull - ll
        return cc[n:]           # g rows (finite: dynamics are equalities, friction rows have one finite side)

    def gval(w):
        return ora.local_system(None, w[:np_], w[np_:])[1]

    h = 1e-6
    idx = rng.choice(n, size=min(n, 25), replace=False)
    for j in idx:
        e = np.zeros(n); e[j] = h
        assert abs((fval(w0 + e) - fval(w0 - e)) / (2 * h) - q[j]) < 1e-5 * max(1.0, abs(q[j]))
        dg = (gval(w0 + e) - gval(w0 - e)) / (2 * h)
        assert np.abs(dg - Hd[:, j]).max() < 1e-5 * max(1.0, np.abs(Hd[:, j]).max())
    if name != "centroidal":   # friction rows have an infinite side on l: use u for those
        for j in idx:
            e = np.zeros(n); e[j] = h
            dc = (cval(w0 + e) - cval(w0 - e)) / (2 * h)
            assert np.abs(dc - J[n:, j]).max() < 1e-5 * max(1.0, np.abs(J[n:, j]).max())
    assert np.array_equal(J[:n, :], np.eye(n))


# ---------------------------------------------------------------------------------------------
# stage code generator
# ---------------------------------------------------------------------------------------------
def test_stage_library_is_generated_and_describes_the_problem(native):
    prob = native.Problem("cartpole", horizon=5)
    lib_path = Path(prob.model_library)
    assert lib_path.exists() and lib_path.with_suffix(".cu").exists()
    src = lib_path.with_suffix(".cu").read_text()
    assert "__global__" in src and "assemble_kernel" in src and "stage_tmpl_" in src
    # interior stages share one template: far fewer templates than stages
    m = re.search(r"groups: (\d+), stage templates: (\d+)", src)
    assert int(m.group(1)) == 6 and int(m.group(2)) <= 4

    class Info(C.Structure):
        _fields_ = [("abi", C.c_int), ("np", C.c_int), ("nf", C.c_int), ("horizon", C.c_int), ("ng", C.c_int),
                    ("n", C.c_int), ("m", C.c_int), ("nnz_h", C.c_int), ("nnz_a", C.c_int),
                    ("h_colptr", C.POINTER(C.c_int)), ("h_rowidx", C.POINTER(C.c_int)),
                    ("a_colptr", C.POINTER(C.c_int)), ("a_rowidx", C.POINTER(C.c_int))]

    lib = C.CDLL(str(lib_path))
    lib.ocp_b200_model_get_info.restype = C.POINTER(Info)
    info = lib.ocp_b200_model_get_info().contents
    assert (info.abi, info.np, info.nf, info.horizon, info.ng) == (1, prob.np_, prob.nf, prob.horizon, prob.ng)
    assert (info.n, info.m, info.nnz_h, info.nnz_a) == (prob.n, prob.m, prob.nnz_h, prob.nnz_a)
    assert np.array_equal(np.ctypeslib.as_array(info.a_rowidx, (prob.nnz_a,)), prob.a_rowidx)
    assert np.array_equal(np.ctypeslib.as_array(info.h_colptr, (prob.n + 1,)), prob.h_colptr)
    for sym in ("ocp_b200_model_assemble", "ocp_b200_model_objective"):
        assert hasattr(lib, sym)


def test_sample_inputs_are_deterministic(native):
    prob = native.Problem("quadrotor")
    a, ra = prob.sample_inputs(16, 0xB202)
    b, rb = prob.sample_inputs(16, 0xB202)
    c, _ = prob.sample_inputs(16, 0xB203)
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and not np.array_equal(a, c)
    assert (np.abs(a[:, :3]) <= 1.0).all() and (np.abs(a[:, 3:6]) <= 0.3).all() and np.allclose(a[:, 12:], 9.81 / 4)
    oa, _ = _oracle.OracleProblem("quadrotor").sample_inputs(16, 0xB202)
    assert np.array_equal(a, oa)


# ---------------------------------------------------------------------------------------------
# stage-periodic index templates of the compact throughput kernel (csrc/periodic_index.h)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,horizon", [("quadrotor", 20), ("cartpole", 40), ("centroidal", 16)])
def test_periodic_index_compression_round_trips(native, name, horizon):
    """The builder expands what it compressed and compares (values through the outer index AND through the
    position); here on the real CSC / CSR structures of the benchmark problems, which must shrink a lot."""
    lib = native.cuda_lib()
    lib.ocp_b200_internal_compress_index.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4
    prob = native.Problem(name, horizon=horizon)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    regions = C.c_int(0)
    ac, ar = prob.a_colptr.astype(np.int32), prob.a_rowidx.astype(np.int32)
    words = lib.ocp_b200_internal_compress_index(prob.n, ip(ac), ip(ar), None, C.byref(regions))
    assert 0 < words < (prob.nnz_a + prob.n) // 2 and regions.value <= 12
    # CSR view with the CSR -> CSC position permutation as second payload
    order = np.lexsort((np.repeat(np.arange(prob.n), np.diff(ac)), ar))
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(ar, minlength=prob.m))]).astype(np.int32)
    colidx = np.repeat(np.arange(prob.n), np.diff(ac))[order].astype(np.int32)
    perm = order.astype(np.int32)
    words = lib.ocp_b200_internal_compress_index(prob.m, ip(rowptr), ip(colidx), ip(perm), C.byref(regions))
    assert 0 < words < (2 * prob.nnz_a + prob.m) // 2 and regions.value <= 12
    # a pattern without stage structure stays explicit (one region, no saving) but still round-trips
    rng = np.random.default_rng(3)
    cnt = rng.integers(0, 5, size=40)
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    val = np.concatenate([np.sort(rng.choice(60, size=c, replace=False)) for c in cnt] + [np.zeros(0, int)]).astype(np.int32)
    words = lib.ocp_b200_internal_compress_index(40, ip(ptr), ip(val), None, C.byref(regions))
    assert words >= val.size + 41
