"""The symbolic layer (casadi-lite: patterns and values of the local system) against witnesses that do not link it:
hand-derived literal CCS arrays (tests/golden/hand_patterns.py) and an independent Python model
(tests/indep_models.py: structural dependency propagation for the patterns, complex-step differentiation for the
values).  SURVEY.md 8c: "sparsity pattern and CSC index assembly bit-exact" needs a witness outside the product."""
import sys
from pathlib import Path

import numpy as np
import pytest

import _oracle
import indep_models as W

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import hand_patterns as HP  # noqa: E402


def _same(a, b):
    return all(np.array_equal(np.asarray(x, np.int64), np.asarray(y, np.int64)) for x, y in zip(a, b))


@pytest.mark.parametrize("case", [1, 2, 3, 4, 5, 6, 7])
def test_kat_patterns_hand_vs_witness_vs_casadi_lite(case):
    hand = HP.KAT[case]
    assert _same(hand, W.kat_patterns(case))            # the propagator reproduces the hand derivation
    assert _same(hand, _oracle.kat_patterns(case))      # casadi-lite (as linked into the oracle) does too


def test_cartpole_h2_patterns_hand_vs_witness_vs_casadi_lite(native):
    hand = HP.CARTPOLE_H2
    assert _same(hand, W.model_patterns("cartpole", 2))
    prob = native.Problem("cartpole", horizon=2)
    assert _same(hand, (prob.h_colptr, prob.h_rowidx, prob.a_colptr, prob.a_rowidx))
    ora = _oracle.OracleProblem("cartpole", horizon=2)
    assert _same(hand, (ora.h_colptr, ora.h_rowidx, ora.a_colptr, ora.a_rowidx))


@pytest.mark.parametrize("name,horizon", [("quadrotor", 20), ("cartpole", 200), ("centroidal", 50), ("quadrotor", 3),
                                          ("centroidal", 2)])
def test_model_patterns_match_the_independent_propagator(native, name, horizon):
    """Full-size patterns of the three benchmark problems, bit-exact (colptr and rowidx of H and J)."""
    wit = W.model_patterns(name, horizon)
    prob = native.Problem(name, horizon=horizon)
    assert _same(wit, (prob.h_colptr, prob.h_rowidx, prob.a_colptr, prob.a_rowidx))
    ora = _oracle.OracleProblem(name, horizon=horizon)
    assert _same(wit, (ora.h_colptr, ora.h_rowidx, ora.a_colptr, ora.a_rowidx))
    if name == "quadrotor":
        # J[1] == J[0] makes the gyroscopic term of the yaw rate 0 * wx * wy: no structural dependence of
        # omega_z' on omega_x / omega_y beyond the identity -- a case where constant folding decides the pattern
        n = prob.n
        col = prob.np_ + 9                                   # omega_x of stage 0
        rows = prob.a_rowidx[prob.a_colptr[col]:prob.a_colptr[col + 1]]
        assert n + 11 not in rows                            # defect row of omega_z


def _gather(dense, colptr, rowidx):
    out = np.empty(len(rowidx))
    for j in range(len(colptr) - 1):
        for k in range(colptr[j], colptr[j + 1]):
            out[k] = dense[rowidx[k], j]
    return out


@pytest.mark.parametrize("name,horizon", [("quadrotor", 20), ("cartpole", 12), ("centroidal", 4)])
def test_local_system_values_match_the_independent_model(name, horizon):
    """H, grad f, J, l - c, u - c of the oracle (casadi-lite AD on the CPU) against hand-written derivatives of the
    cost and complex-step derivatives of the RK4 defects; structural zeros of the pattern must be zeros of the
    independent model as well (nothing outside the pattern)."""
    ora = _oracle.OracleProblem(name, horizon=horizon)
    frames, refs = ora.sample_inputs(2, 0xB200 + 3)
    rng = np.random.default_rng(7)
    for b in range(2):
        x = np.tile(frames[b], horizon) + 0.05 * rng.standard_normal(ora.N)
        p = refs[b] + 0.05 * rng.standard_normal(ora.np_)
        hv, q, av, l, u = ora.local_system(frames[b], p, x)
        Hd, grad, J, ld, ud = W.local_system_dense(name, horizon, p, x, frame=frames[b])
        scale = lambda a: max(1.0, float(np.abs(a).max()))
        assert np.abs(hv - _gather(Hd, ora.h_colptr, ora.h_rowidx)).max() < 1e-11 * scale(hv)
        assert np.abs(av - _gather(J, ora.a_colptr, ora.a_rowidx)).max() < 1e-10 * scale(av)
        assert np.abs(q - grad).max() < 1e-11 * scale(q)
        # nothing of the independent model lies outside the structural pattern
        mask = np.zeros_like(J, bool)
        for j in range(ora.n):
            mask[ora.a_rowidx[ora.a_colptr[j]:ora.a_colptr[j + 1]], j] = True
        assert np.abs(J[~mask]).max() == 0.0
        maskh = np.zeros_like(Hd, bool)
        for j in range(ora.n):
            maskh[ora.h_rowidx[ora.h_colptr[j]:ora.h_colptr[j + 1]], j] = True
        assert np.abs(Hd[~maskh]).max() == 0.0
        for mine, theirs in ((l, ld), (u, ud)):
            fin = np.isfinite(theirs)
            assert np.array_equal(np.isfinite(mine), fin)
            assert np.abs(mine[fin] - theirs[fin]).max() < 1e-11 * scale(theirs[fin])
        assert abs(ora.objective(p, x) - W.objective_value(name, horizon, p, x)) < 1e-11 * max(1.0, abs(ora.objective(p, x)))


def test_bounds_match_the_independent_model():
    for name, H in (("quadrotor", 5), ("cartpole", 7), ("centroidal", 3)):
        ora = _oracle.OracleProblem(name, horizon=H)
        lbx, ubx, lbg, ubg = W.model_bounds(name, H)
        assert np.array_equal(lbx, ora.lbx) and np.array_equal(ubx, ora.ubx)
        assert np.array_equal(lbg, ora.lbg) and np.array_equal(ubg, ora.ubg)
