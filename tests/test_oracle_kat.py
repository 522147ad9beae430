"""Pins the oracle (CPU restatement of the reference's SQP + OSQP path) on the only known answers
the reference itself carries: the analytic optima printed by test/test.cpp:13-185 (fixture
tests/golden/kat_expected.json).  The reference prints and never asserts; the tolerance here is
what OSQP at eps_abs = eps_rel = 1e-3 (SQPOptimizationSolver.cpp:83-84) guarantees on these
well-scaled 2-3 variable QPs."""
import json
from pathlib import Path

import numpy as np
import pytest

import _oracle

GOLDEN = json.loads((Path(__file__).parent / "golden" / "kat_expected.json").read_text())


@pytest.mark.parametrize("case", sorted(GOLDEN))
@pytest.mark.parametrize("use_float", [False, True])
def test_kat_one_full_step(case, use_float):
    """alpha = 1, one SQP step: cases 1-7 are QPs with linear constraints, one step solves them."""
    x, f, st = _oracle.kat_solve(int(case), 1, 1.0, use_float)
    assert st[0] == 1  # OSQP_SOLVED
    assert np.abs(x - np.array(GOLDEN[case]["x"])).max() < 2e-3
    assert np.array_equal(_oracle.kat_expected(int(case)), np.array(GOLDEN[case]["x"]))


@pytest.mark.parametrize("case", sorted(GOLDEN))
def test_kat_reference_schedule_converges(case):
    """The reference's own schedule: step_num fixed steps of alpha = 0.1 (OptimalControlProblem.h:24-27)
    shrink the distance to the optimum by 0.9 per step: x_k = x* (1 - 0.9^k) from x_0 = 0."""
    target = np.array(GOLDEN[case]["x"])
    x10, _, st = _oracle.kat_solve(int(case), 10, 0.1)
    assert st[1] == 10
    assert np.abs(x10 - target * (1 - 0.9 ** 10)).max() < 5e-3
    x60, _, _ = _oracle.kat_solve(int(case), 60, 0.1)
    assert np.abs(x60 - target).max() < 1e-2


def test_kat_nonconvex_case_runs():
    """test/test.cpp:187-211: indefinite Hessian, outside OSQP's domain; only 'does not crash'."""
    x, f, st = _oracle.kat_solve(8, 2, 0.1)
    assert x.shape == (2,)


def test_oracle_regression_fixture():
    """The restatement reproduces its own committed iterates (compiler / refactoring guard)."""
    g = np.load(Path(__file__).parent / "golden" / "oracle_sqp.npz")
    o = _oracle.OracleProblem("quadrotor", alpha=0.1, step_num=10)
    x, f, st = o.solve_batch(g["frames"], g["refs"])
    assert np.allclose(x, g["x_a01"], rtol=0, atol=1e-9)
    assert np.array_equal(st[:, :3], g["st_a01"][:, :3])
    o.set_schedule(5, 1.0)
    x1, f1, st1 = o.solve_batch(g["frames"], g["refs"])
    assert np.allclose(x1, g["x_a1"], rtol=0, atol=1e-8)
    assert np.allclose(f1, g["f_a1"], rtol=1e-9)
