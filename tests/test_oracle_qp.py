"""Independent cross-checks of the OSQP restatement (oracle/osqp_restate.hpp) -- the real OSQP is
not in this image (parity unpinned), so the restatement is checked against numpy/scipy solutions
and against the optimality conditions of the QP itself."""
import numpy as np
import pytest
import scipy.optimize as so
import scipy.sparse as sp

import _oracle


def random_qp(seed, n=8, m_extra=5, n_eq=2):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    P = M @ M.T + 0.1 * np.eye(n)
    P[np.abs(P) < 0.3] = 0.0
    P = (P + P.T) / 2 + n * np.eye(n)
    q = rng.standard_normal(n)
    G = rng.standard_normal((m_extra, n)) * (rng.random((m_extra, n)) < 0.6)
    A = np.vstack([np.eye(n), G])
    x_feas = rng.standard_normal(n)
    l = np.concatenate([x_feas - rng.random(n), G @ x_feas - rng.random(m_extra)])
    u = np.concatenate([x_feas + rng.random(n), G @ x_feas + rng.random(m_extra)])
    l[n:n + n_eq] = u[n:n + n_eq] = (G @ x_feas)[:n_eq]        # equalities
    u[n + n_eq] = np.inf                                        # one one-sided row
    l[0] = -np.inf; u[0] = np.inf                               # one free row
    return P, q, A, l, u


def csc(M):
    S = sp.csc_matrix(M)
    S.sort_indices()
    return S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.astype(np.float64)


def solve_with_oracle(P, q, A, l, u, **kw):
    n, m = P.shape[0], A.shape[0]
    hp, hi, hx = csc(P)
    ap, ai, ax = csc(A)
    return _oracle.qp_solve(n, m, hp, hi, hx, q, ap, ai, ax, l, u, settings=_oracle.settings_vector(**kw))


@pytest.mark.parametrize("seed", range(6))
def test_against_scipy(seed):
    P, q, A, l, u = random_qp(seed)
    x, y, info, _ = solve_with_oracle(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=50000)
    assert info[0] == 1
    n = P.shape[0]
    cons = so.LinearConstraint(A[n:], l[n:], u[n:])
    res = so.minimize(lambda z: 0.5 * z @ P @ z + q @ z, np.zeros(n), jac=lambda z: P @ z + q, hess=lambda z: P,
                      bounds=so.Bounds(l[:n], u[:n]), constraints=[cons], method="trust-constr",
                      options=dict(gtol=1e-12, xtol=1e-14, maxiter=5000))
    assert np.abs(x - res.x).max() < 2e-5
    assert abs((0.5 * x @ P @ x + q @ x) - res.fun) < 1e-5


@pytest.mark.parametrize("seed", range(6))
def test_kkt_conditions(seed):
    P, q, A, l, u = random_qp(100 + seed, n=12, m_extra=9, n_eq=3)
    x, y, info, _ = solve_with_oracle(P, q, A, l, u, eps_abs=1e-8, eps_rel=1e-8, max_iter=50000)
    assert info[0] == 1
    z = A @ x
    assert np.abs(P @ x + q + A.T @ y).max() < 1e-6                  # stationarity
    assert (z >= l - 1e-6).all() and (z <= u + 1e-6).all()           # primal feasibility
    act_l, act_u = np.abs(z - l) < 1e-5, np.abs(z - u) < 1e-5
    assert (np.abs(y[~act_l & ~act_u]) < 1e-5).all()                 # complementarity
    assert (y[act_l & ~act_u] <= 1e-6).all() and (y[act_u & ~act_l] >= -1e-6).all()


def test_equality_only_qp_is_the_kkt_solve():
    rng = np.random.default_rng(3)
    n, me = 10, 4
    M = rng.standard_normal((n, n)); P = M @ M.T + np.eye(n)
    q = rng.standard_normal(n); E = rng.standard_normal((me, n)); b = rng.standard_normal(me)
    K = np.block([[P, E.T], [E, np.zeros((me, me))]])
    sol = np.linalg.solve(K, np.concatenate([-q, b]))
    x, y, info, _ = solve_with_oracle(P, q, E, b, b, eps_abs=1e-10, eps_rel=1e-10, max_iter=100000)
    assert np.abs(x - sol[:n]).max() < 1e-7
    assert np.abs(y - sol[n:]).max() < 1e-6


def test_only_upper_triangle_of_P_is_read():
    """OsqpEigen hands OSQP the upper triangle: garbage below the diagonal is ignored."""
    P, q, A, l, u = random_qp(7)
    x0, *_ = solve_with_oracle(P, q, A, l, u)
    P2 = np.triu(P) + 17.0 * np.tril(np.where(P != 0, 1.0, 0.0), -1)
    x1, *_ = solve_with_oracle(P2, q, A, l, u)
    assert np.array_equal(x0, x1)


def test_infeasibility_certificates():
    # primal infeasible: x in [0, 1] and x >= 2
    x, y, info, _ = _oracle.qp_solve(1, 2, [0, 1], [0], [1.0], [0.0], [0, 2], [0, 1], [1.0, 1.0], [0.0, 2.0], [1.0, 3.0])
    assert info[0] == 3 and np.isnan(x).all()
    # dual infeasible: min -x, x >= 0, no curvature
    x, y, info, _ = _oracle.qp_solve(1, 1, [0, 0], [], [], [-1.0], [0, 1], [0], [1.0], [0.0], [np.inf])
    assert info[0] == 5 and np.isnan(x).all()


def test_lower_above_upper_is_rejected():
    with pytest.raises(RuntimeError):
        _oracle.qp_solve(1, 1, [0, 1], [0], [1.0], [0.0], [0, 1], [0], [1.0], [2.0], [1.0])


def test_termination_checked_every_25_iterations_and_rho_adapts():
    P, q, A, l, u = random_qp(11, n=10, m_extra=8)
    x, y, info, trace = solve_with_oracle(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=20000)
    assert (trace[:, 0] % 25 == 0).all() and info[1] % 25 == 0
    x2, y2, info2, trace2 = solve_with_oracle(P, q, A, l, u, eps_abs=1e-9, eps_rel=1e-9, max_iter=20000, adaptive_rho=0)
    assert info2[6] == 0 and np.allclose(trace2[:, 3], 0.1)
    assert np.abs(x - x2).max() < 1e-6


def test_float_build_tracks_double():
    """Real = float mirrors the reference's OSQP_USE_FLOAT=ON build (cpu_install.sh:44)."""
    P, q, A, l, u = random_qp(5)
    n, m = P.shape[0], A.shape[0]
    hp, hi, hx = csc(P); ap, ai, ax = csc(A)
    xd, *_ = _oracle.qp_solve(n, m, hp, hi, hx, q, ap, ai, ax, l, u)
    xf, *_ = _oracle.qp_solve(n, m, hp, hi, hx, q, ap, ai, ax, l, u, use_float=True)
    assert np.abs(xd - xf).max() < 5e-3
