"""a CasADi-generated C file behind include/ocp_b200_model.h.

CPU: the C file casadi-lite emits follows the layout of CasADi's code generator -- compiled as plain C with the
system compiler and called through the `(arg, res, iw, w, mem)` signature it reproduces the oracle's local system,
its compact-CCS `_sparsity_out` arrays are the patterns; the committed fixture tests/golden/casadi_format/ is such a
file; CasadiCInterop turns it into a stage library exporting the model ABI (nvcc cross-compiles without a GPU).
GPU: local system and SQP solve through that library against the oracle."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import _oracle

ROOT = Path(__file__).resolve().parent.parent
FIXTURE = ROOT / "tests" / "golden" / "casadi_format" / "localSystemFunction_cartpole_h3.c"
H = 3


def _decode(sp):
    nrow, ncol = sp[0], sp[1]
    if sp[2] != 0:
        return nrow, ncol, np.arange(ncol + 1) * nrow, np.tile(np.arange(nrow), ncol)
    colind = np.array([sp[2 + j] for j in range(ncol + 1)])
    return nrow, ncol, colind, np.array([sp[2 + ncol + 1 + k] for k in range(colind[-1])])


def _host_library(c_file, tmp_path):
    so = tmp_path / "f.so"
    r = subprocess.run(["cc", "-shared", "-fPIC", "-O1", "-x", "c", str(c_file), "-o", str(so), "-lm"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return C.CDLL(str(so))


@pytest.mark.parametrize("source", ["emitted", "fixture"])
def test_generated_c_file_follows_the_casadi_layout(native, tmp_path, source):
    prob = native.Problem("cartpole", horizon=H)
    c_file = Path(prob.generate_c(str(tmp_path / "localSystemFunction.c"))) if source == "emitted" else FIXTURE
    text = c_file.read_text()
    for token in ("#define casadi_real double", "#define casadi_int long long int", "CASADI_SYMBOL_EXPORT int localSystemFunction(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem)",
                  "localSystemFunction_sparsity_out", "localSystemFunction_work", "CASADI_SYMBOL_EXPORT int objective("):
        assert token in text
    lib = _host_library(c_file, tmp_path)
    ll = C.c_longlong
    lib.localSystemFunction_sparsity_out.restype = C.POINTER(ll); lib.localSystemFunction_sparsity_out.argtypes = [ll]
    lib.localSystemFunction_sparsity_in.restype = C.POINTER(ll); lib.localSystemFunction_sparsity_in.argtypes = [ll]
    lib.localSystemFunction_n_in.restype = ll; lib.localSystemFunction_n_out.restype = ll
    assert lib.localSystemFunction_n_in() == 4 and lib.localSystemFunction_n_out() == 5
    ora = _oracle.OracleProblem("cartpole", horizon=H)
    nr, nc, hc, hr = _decode(lib.localSystemFunction_sparsity_out(0))
    assert (nr, nc) == (ora.n, ora.n) and np.array_equal(hc, ora.h_colptr) and np.array_equal(hr, ora.h_rowidx)
    nr, nc, ac, ar = _decode(lib.localSystemFunction_sparsity_out(2))
    assert (nr, nc) == (ora.m, ora.n) and np.array_equal(ac, ora.a_colptr) and np.array_equal(ar, ora.a_rowidx)
    assert _decode(lib.localSystemFunction_sparsity_in(1))[:2] == (ora.N, 1)
    sz = [ll(0) for _ in range(4)]
    lib.localSystemFunction_work(*[C.byref(v) for v in sz])
    w = np.zeros(max(1, sz[3].value)); iw = (ll * max(1, sz[2].value))()
    frames, refs = ora.sample_inputs(1, 9)
    rng = np.random.default_rng(4)
    x = np.tile(frames[0], H) + 0.05 * rng.standard_normal(ora.N)
    p = refs[0] + 0.1
    lbx, ubx = ora.lbx.copy(), ora.ubx.copy()
    lbx[:ora.nf] = frames[0]; ubx[:ora.nf] = frames[0]
    lfull = np.concatenate([p, lbx, ora.lbg]); ufull = np.concatenate([p, ubx, ora.ubg])
    hv = np.zeros(ora.nnz_h); q = np.zeros(ora.n); av = np.zeros(ora.nnz_a); l = np.zeros(ora.m); u = np.zeros(ora.m)
    dp = C.POINTER(C.c_double)
    arg = (dp * 4)(*[a.ctypes.data_as(dp) for a in (p, x, lfull, ufull)])
    res = (dp * 5)(*[a.ctypes.data_as(dp) for a in (hv, q, av, l, u)])
    assert lib.localSystemFunction(arg, res, iw, w.ctypes.data_as(dp), 0) == 0
    ohv, oq, oav, ol, ou = ora.local_system(frames[0], p, x)
    for mine, ref in ((hv, ohv), (q, oq), (av, oav)):
        assert np.abs(mine - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    fin = np.isfinite(ol)
    assert np.array_equal(np.isfinite(l), fin) and np.abs(l[fin] - ol[fin]).max() <= 1e-13
    f = np.zeros(1)
    lib.objective_work(*[C.byref(v) for v in sz])
    w2 = np.zeros(max(1, sz[3].value))
    arg2 = (dp * 2)(p.ctypes.data_as(dp), x.ctypes.data_as(dp)); res2 = (dp * 1)(f.ctypes.data_as(dp))
    assert lib.objective(arg2, res2, iw, w2.ctypes.data_as(dp), 0) == 0
    assert abs(f[0] - ora.objective(p, x)) <= 1e-12 * max(1.0, abs(f[0]))


def test_interop_builds_a_stage_library_with_the_model_abi(native, tmp_path):
    so = native.compile_casadi_c(str(FIXTURE), nf=5, horizon=H, name="cartpole_c", code_dir=str(tmp_path))
    assert Path(so).exists() and Path(so).with_suffix(".cu").exists()
    sym = subprocess.run(["nm", "-D", "--defined-only", so], stdout=subprocess.PIPE, text=True).stdout
    for name in ("ocp_b200_model_get_info", "ocp_b200_model_assemble", "ocp_b200_model_objective"):
        assert name in sym
    sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in sass and "casadi_assemble_kernel" in sass
    with pytest.raises(RuntimeError, match="is missing|must map"):
        native.compile_casadi_c(str(FIXTURE), nf=5, horizon=H, name="bad", local_system_fn="objective", code_dir=str(tmp_path))
    with pytest.raises(RuntimeError, match="stage layout"):
        native.compile_casadi_c(str(FIXTURE), nf=4, horizon=H, name="bad2", code_dir=str(tmp_path))


@pytest.mark.gpu
def test_casadi_c_stage_library_matches_the_oracle(native, tmp_path):
    prob = native.Problem("cartpole", horizon=H)
    ora = _oracle.OracleProblem("cartpole", horizon=H)
    so = native.compile_casadi_c(str(FIXTURE), nf=prob.nf, horizon=H, name="cartpole_c", code_dir=str(tmp_path))
    s = prob.get_settings()
    s.sqp_alpha, s.sqp_step_num = 0.5, 3
    sol = native.Solver.create(prob.n, prob.m, prob.h_colptr, prob.h_rowidx, prob.a_colptr, prob.a_rowidx, settings=s,
                               np_=prob.np_, nf=prob.nf, horizon=H, model_library=so)
    B = 5
    frames, refs = prob.sample_inputs(B, 0xB200 + 41)
    rng = np.random.default_rng(8)
    x = np.tile(frames, (1, H)) + 0.05 * rng.standard_normal((B, prob.N))
    hv, q, av, l, u = sol.export_qp(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x)
    for b in range(B):
        ohv, oq, oav, ol, ou = ora.local_system(frames[b], refs[b], x[b])
        for mine, ref in ((hv[b], ohv), (q[b], oq), (av[b], oav)):
            assert np.abs(mine - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        fin = np.isfinite(ol)
        assert np.array_equal(np.isfinite(l[b]), fin) and np.abs(l[b][fin] - ol[fin]).max() <= 1e-12
        fin = np.isfinite(ou)
        assert np.abs(u[b][fin] - ou[fin]).max() <= 1e-12
    x0 = np.tile(frames, (1, H))
    xs = x0.copy(); f = np.zeros(B); st = np.zeros((B, native.NSTATS))
    sol.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, xs, f, st)
    ora.set_schedule(3, 0.5)
    ox, of, ost = ora.solve_batch(frames, refs, x0=x0)
    assert np.array_equal(st[:, native.STAT["admm_iters"]], ost[:, 2])
    assert np.abs(xs - ox).max() < 1e-6 * max(1.0, np.abs(ox).max()) and np.allclose(f, of, rtol=1e-6, atol=1e-9)
    # and the same bits as the default (warp-per-stage) stage library
    xs2 = x0.copy()
    prob.solver.update_settings(s)
    prob.solver.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, xs2)
    assert np.abs(xs - xs2).max() < 1e-9
