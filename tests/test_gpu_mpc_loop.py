"""MPC tick loop on the device: receding-horizon shift of the iterates and the
torch-tensor API (`MpcLoop`), against the oracle on the same shifted iterates and in closed loop
against an independent numpy plant."""
import numpy as np
import pytest

import _oracle

pytestmark = pytest.mark.gpu

REL_SOLUTION = 1e-6


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max()))


def shift_numpy(x, nf):
    y = x.copy()
    y[:, :-nf] = x[:, nf:]
    return y


def quadrotor_ode(x, u):
    """numpy restatement of the quadrotor model of problems/problems.cpp (the plant of the closed loop)"""
    m, arm, kappa, g, J = 1.0, 0.17, 0.016, 9.81, (0.01, 0.01, 0.02)
    r, p, yw = x[:, 3], x[:, 4], x[:, 5]
    v, w = x[:, 6:9], x[:, 9:12]
    sr, cr, sp, cp, sy, cy = np.sin(r), np.cos(r), np.sin(p), np.cos(p), np.sin(yw), np.cos(yw)
    tp = np.tan(p)
    T = u.sum(axis=1)
    acc = np.stack([(cy * sp * cr + sy * sr) * T / m, (sy * sp * cr - cy * sr) * T / m, cp * cr * T / m - g], axis=1)
    rates = np.stack([w[:, 0] + sr * tp * w[:, 1] + cr * tp * w[:, 2], cr * w[:, 1] - sr * w[:, 2],
                      (sr / cp) * w[:, 1] + (cr / cp) * w[:, 2]], axis=1)
    tx, ty = arm * (u[:, 1] - u[:, 3]), arm * (u[:, 2] - u[:, 0])
    tz = kappa * (u[:, 0] - u[:, 1] + u[:, 2] - u[:, 3])
    dw = np.stack([(tx - (J[2] - J[1]) * w[:, 1] * w[:, 2]) / J[0], (ty - (J[0] - J[2]) * w[:, 2] * w[:, 0]) / J[1],
                   (tz - (J[1] - J[0]) * w[:, 0] * w[:, 1]) / J[2]], axis=1)
    return np.concatenate([v, rates, acc, dw], axis=1)


def rk4(x, u, dt):
    k1 = quadrotor_ode(x, u)
    k2 = quadrotor_ode(x + 0.5 * dt * k1, u)
    k3 = quadrotor_ode(x + 0.5 * dt * k2, u)
    k4 = quadrotor_ode(x + dt * k3, u)
    return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)


@pytest.mark.parametrize("name,B", [("quadrotor", 7), ("cartpole", 3), ("centroidal", 2)])
def test_shift_iterate_matches_numpy(problems, native, name, B):
    import torch
    prob, _ = problems(name)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, prob.N))
    d_x = torch.from_numpy(x).cuda()
    import ctypes as C
    prob.solver.shift_iterate_device(B, C.c_void_p(d_x.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(d_x.cpu().numpy(), shift_numpy(x, prob.nf))   # pure data movement: bit-exact


def test_two_ticks_with_shift_match_oracle(problems, native):
    import torch
    prob, ora = problems("quadrotor")
    B = 4
    frames, refs = prob.sample_inputs(B, 0xB200 + 21)
    prob.set_schedule(10, 0.1); ora.set_schedule(10, 0.1)
    ora.set_qp_settings(_oracle.settings_from_b200(prob.get_settings()))
    loop = native.MpcLoop(prob, B)
    d_fr, d_p = torch.from_numpy(frames).cuda(), torch.from_numpy(refs).cuda()
    x1, f1, st1 = (t.cpu().numpy().copy() for t in loop.tick(d_fr, d_p, shift=True))
    ox1, of1, _ = ora.solve_batch(frames, refs)
    assert rel_err(x1, ox1) < REL_SOLUTION
    # tick 2: the plant moved to the predicted second frame; warm start = shifted tick-1 trajectory
    frames2 = np.ascontiguousarray(x1[:, prob.nf:2 * prob.nf])
    x2, f2, st2 = (t.cpu().numpy().copy() for t in loop.tick(torch.from_numpy(frames2).cuda(), d_p, shift=True))
    ox2, of2, _ = ora.solve_batch(frames2, refs, x0=shift_numpy(ox1, prob.nf))
    assert rel_err(x2, ox2) < REL_SOLUTION
    assert np.allclose(f2, of2, rtol=1e-6)
    assert (st2[:, native.STAT["sqp_steps"]] == 10).all()
    # without the shift the same tick starts from the unshifted trajectory (the reference's behaviour)
    loop.reset()
    loop.tick(d_fr, d_p)
    x2n = loop.tick(torch.from_numpy(frames2).cuda(), d_p, shift=False)[0].cpu().numpy()
    ox2n, _, _ = ora.solve_batch(frames2, refs, x0=ox1)
    assert rel_err(x2n, ox2n) < REL_SOLUTION


def test_closed_loop_quadrotors_on_device(problems, native):
    """64 quadrotors, 60 ticks of 5 ms against an independent numpy RK4 plant.  The reference pins the
    WHOLE first frame (state and control, OptimalControlProblem.cpp:93-96), so the control applied
    during a tick is the previous tick's stage-1 control and the plant must land on the solver's
    predicted stage-1 state; body rates and attitude errors must shrink; nothing leaves the device
    except the one frame per tick the plant needs."""
    import torch
    prob, _ = problems("quadrotor")
    B, ticks, dt = 64, 60, 0.005
    prob.set_schedule(10, 0.5)
    try:
        frames, refs = prob.sample_inputs(B, 0xB200 + 33)
        state, u_prev = frames[:, :12].copy(), frames[:, 12:].copy()
        att0 = np.linalg.norm(state[:, 3:6], axis=1).mean()
        om0 = np.linalg.norm(state[:, 9:12], axis=1).mean()
        loop = native.MpcLoop(prob, B)
        d_p = torch.from_numpy(refs).cuda()
        launches0 = prob.solver.launch_count()
        for t in range(ticks):
            fr = np.concatenate([state, u_prev], axis=1)
            x, f, st = loop.tick(torch.from_numpy(fr).cuda(), d_p, shift=True)
            nxt = loop.frame(1).cpu().numpy()
            assert np.isfinite(nxt).all()
            assert (nxt[:, 12:] >= prob.lbx[12:16] - 1e-6).all() and (nxt[:, 12:] <= prob.ubx[12:16] + 1e-6).all()
            state = rk4(state, u_prev, dt)
            assert np.abs(nxt[:, :12] - state).max() < (5e-3 if t == 0 else 1e-4)
            u_prev = nxt[:, 12:].copy()
        assert (st.cpu().numpy()[:, native.STAT["qp_status"]] == native.QP_SOLVED).all()
        assert np.linalg.norm(state[:, 9:12], axis=1).mean() < 0.7 * om0
        assert np.linalg.norm(state[:, 3:6], axis=1).mean() < 0.95 * att0
        # two kernels per SQP step + objective + stats store per tick, one shift per tick after the first
        assert prob.solver.launch_count() - launches0 == ticks * (2 * 10 + 2) + (ticks - 1)
    finally:
        prob.set_schedule(10, 0.1)
