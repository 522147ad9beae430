"""The Python problem-definition front-end (optimal_control_problem_b200.symbolic):
a cart-pole written in Python with the reference's method names must generate exactly the problem the
C++ subclass of problems/problems.cpp generates."""
import numpy as np
import pytest

H = 12


def make_python_cartpole(native, horizon=H, name="cartpole"):
    sym = __import__("optimal_control_problem_b200.symbolic", fromlist=["x"])
    SX, sin, cos, vertcat = sym.SX, sym.sin, sym.cos, sym.vertcat
    cart_m, pole_m, pole_l, g = 1.0, 0.1, 0.5, 9.81

    def ode(x, u):
        dth, ds = x[3], x[2]
        s, c = sin(x[1]), cos(x[1])
        total = cart_m + pole_m
        temp = (u[0] + pole_m * pole_l * dth * dth * s) / total
        ddth = (g * s - c * temp) / (pole_l * (4.0 / 3.0 - pole_m * c * c / total))
        dds = temp - pole_m * pole_l * ddth * c / total
        return vertcat([ds, dth, dds, ddth])

    def rk4(x, u, dt):
        k1 = ode(x, u)
        k2 = ode(x + (0.5 * dt) * k1, u)
        k3 = ode(x + (0.5 * dt) * k2, u)
        k4 = ode(x + dt * k3, u)
        return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)

    class CartPole(sym.OptimalControlProblem):
        def deploy_constraints_and_add_cost(self):
            for k in range(self.horizon):
                xk, uk = self.get_variable(k, "state"), self.get_variable(k, "force")
                self.add_vector_cost([1.0, 10.0, 0.1, 0.1], xk - self.reference_)
                self.add_vector_cost([0.01], uk)
                if k + 1 < self.horizon:
                    self.add_equation_constraint("dynamics", self.get_variable(k + 1, "state"), rk4(xk, uk, self.dt))

    ocp = CartPole(sym.default_yaml("cartpole", horizon=horizon), name=name)
    ocp.set_reference(SX.sym("ref", 4))
    ocp.deploy_constraints_and_add_cost()
    ocp.gen_solver()
    return ocp


def test_sx_handles(native):
    sym = __import__("optimal_control_problem_b200.symbolic", fromlist=["x"])
    x = sym.SX.sym("x", 3)
    assert len(x) == 3 and len(x[1:3]) == 2 and len(x[-1]) == 1
    y = sym.vertcat([x, 2.0 * x[0] - 1.0, sym.sin(x[1]) ** 2])
    assert len(y) == 5
    with pytest.raises(IndexError):
        x[3]
    with pytest.raises(RuntimeError, match="unknown SX operation"):
        sym.SX._unary("no_such_op", x)


def test_python_cartpole_equals_cpp_cartpole(native):
    ocp = make_python_cartpole(native)
    ref = native.Problem("cartpole", horizon=H)
    p = ocp.problem
    assert p.dims == ref.dims
    for a in ("h_colptr", "h_rowidx", "a_colptr", "a_rowidx"):
        assert np.array_equal(getattr(p, a), getattr(ref, a)), a
    for a in ("lbx", "ubx", "lbg", "ubg"):
        assert np.array_equal(getattr(p, a), getattr(ref, a)), a
    # (the generated sources are not textually identical: the symbolic layer orders the operands of
    # commutative operations by node id, and Python creates its constants in a different order)
    assert p.model_library.endswith(".so") and p.model_library != ref.model_library
    assert ocp.get_horizon() == H and ocp.get_frame_size() == 5 and ocp.get_dt() == 0.01
    assert np.array_equal(ocp.get_constraint_lower_bounds(), ref.lbg)


def test_registration_errors_surface_as_exceptions(native):
    sym = __import__("optimal_control_problem_b200.symbolic", fromlist=["x"])
    ocp = sym.OptimalControlProblem(sym.default_yaml("cartpole", horizon=4), name="broken")
    with pytest.raises(NotImplementedError):
        ocp.deploy_constraints_and_add_cost()
    with pytest.raises(RuntimeError):
        ocp.get_variable(0, "no_such_field")
    with pytest.raises(RuntimeError, match="gen_solver"):
        ocp.problem


@pytest.mark.gpu
def test_python_cartpole_solves_like_cpp_cartpole(native):
    ocp = make_python_cartpole(native)
    ref = native.Problem("cartpole", horizon=H)
    frames, refs = ref.sample_inputs(3, 0xB200 + 5)
    out = []
    for prob in (ocp.problem, ref):
        x = np.tile(frames, (1, H))   # x = 0 would make the first cart-pole QP infeasible
        f = np.zeros(3); st = np.zeros((3, native.NSTATS))
        prob.solver.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x, f, st)
        out.append((x, f, st))
    (xa, fa, sa), (xb, fb, sb) = out
    assert np.isfinite(xa).all()
    # same expressions up to the operand order of commutative operations (nvcc may contract a different
    # product of a sum into the FMA): agreement to rounding, identical iteration counts
    assert np.abs(xa - xb).max() <= 1e-9 * max(1.0, np.abs(xb).max())
    assert np.allclose(fa, fb, rtol=1e-9)
    for k in ("qp_status", "sqp_steps", "admm_iters", "checks"):
        assert np.array_equal(sa[:, native.STAT[k]], sb[:, native.STAT[k]]), k
    # ... and against the CPU oracle, not only GPU against GPU: the Python-defined problem solves to the oracle's answer
    import _oracle
    ora = _oracle.OracleProblem("cartpole", horizon=H)
    ora.set_qp_settings(_oracle.settings_from_b200(ocp.problem.get_settings()))
    ox, of, ost = ora.solve_batch(frames, refs, x0=np.tile(frames, (1, H)))
    assert np.abs(xa - ox).max() <= 1e-6 * max(1.0, np.abs(ox).max())
    assert np.allclose(fa, of, rtol=1e-6, atol=1e-9)
    assert np.array_equal(sa[:, native.STAT["admm_iters"]], ost[:, 2])
    # the class API of the reference's pybind module
    x1 = ocp.compute_optimal_trajectory(frames[0], refs[0])
    assert x1.shape == (5 * H,)
    assert np.array_equal(ocp.get_optimal_input_first_frame(), x1[:5])
    assert np.array_equal(ocp.get_optimal_trajectory(), x1)
