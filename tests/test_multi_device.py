"""ocp_b200_create_multi / ocp_b200_solve_batch_multi (include/ocp_b200.h): one process, several GPUs.
CPU: the block partition (pure arithmetic) and the error path without a device.  GPU: the multi entry point with the
same device listed twice (exercises the per-device threads, slicing and error collection on a 1-GPU box) and, with
at least two GPUs, with real devices; results must be bit-identical to the single-handle path."""
import numpy as np
import pytest


def test_partition_is_a_contiguous_cover(native):
    for B in (0, 1, 7, 8, 9, 4096, 65536):
        for ndev in (1, 2, 3, 8):
            off = native.multi_partition(B, ndev)
            assert off[0] == 0 and off[-1] == B and (np.diff(off) >= 0).all()
            per = -(-B // ndev)
            assert (np.diff(off) <= per).all()
            full = B // per if per else 0
            assert (np.diff(off)[:full] == per).all()           # leading devices get ceil(B / ndev)
    with pytest.raises(native.OcpB200Error):
        native.multi_partition(5, 0)


def test_create_multi_without_a_device_fails_loudly(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    prob = native.Problem("cartpole", horizon=5)
    with pytest.raises(native.OcpB200Error, match="no CUDA device"):
        native.MultiSolver(prob, [0, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], "all"])
def test_multi_device_batch_equals_single_handle(native, devices):
    import torch
    if devices == "all":
        if torch.cuda.device_count() < 2:
            pytest.skip("needs at least two GPUs")
        devices = list(range(torch.cuda.device_count()))
    prob = native.Problem("quadrotor")
    B = 301                                      # not a multiple of the device count: ragged last block
    frames, refs = prob.sample_inputs(B, 0xB200 + 31)
    s = prob.get_settings()
    s.sqp_alpha, s.sqp_step_num = 0.5, 3
    prob.solver.update_settings(s)
    x1 = np.zeros((B, prob.N)); f1 = np.zeros(B); st1 = np.zeros((B, native.NSTATS))
    prob.solver.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x1, f1, st1)
    multi = native.MultiSolver(prob, devices, settings=s)
    x2 = np.zeros((B, prob.N)); f2 = np.zeros(B); st2 = np.zeros((B, native.NSTATS))
    multi.solve_batch(frames, refs, x2, f2, st2)
    # blocks smaller than the SM count run the latency plan, the single handle ran the throughput plan:
    # same arithmetic (tests/test_gpu_parity.py::test_throughput_plan_matches_oracle_and_latency_plan)
    assert np.array_equal(x1, x2) and np.array_equal(f1, f2)
    assert np.array_equal(st1[:, native.STAT["admm_iters"]], st2[:, native.STAT["admm_iters"]])
    multi.close()


@pytest.mark.gpu
def test_class_api_batch_over_device_list(native):
    """OptimalControlProblem::computeOptimalTrajectoryBatch with SQPOptimizationSolver::setDevices."""
    prob = native.Problem("cartpole", horizon=12, alpha=0.5, step_num=2)
    frames, refs = prob.sample_inputs(10, 5)
    x1, f1, st1 = prob.compute_optimal_trajectory_batch(frames, refs)
    prob.reset()
    prob.set_devices([0, 0])
    x2, f2, st2 = prob.compute_optimal_trajectory_batch(frames, refs)
    assert np.array_equal(x1, x2) and np.array_equal(f1, f2)
