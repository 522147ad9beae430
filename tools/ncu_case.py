"""Small fixed workload for ncu: one batched solve of B quadrotor instances (default 148 = one
per SM, one wave) with step_num SQP steps.  usage: python tools/ncu_case.py [B] [step_num] [problem]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import optimal_control_problem_b200 as ocp  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
name = sys.argv[3] if len(sys.argv) > 3 else "quadrotor"
prob = ocp.Problem(name, alpha=0.1, step_num=steps)
frames, refs = prob.sample_inputs(B, 0xB202)
x = np.zeros((B, prob.N)); st = np.zeros((B, ocp.NSTATS))
prob.solver.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x, None, st)
print("ok", name, B, steps, "admm iters", st[:, 2].sum(), "launches", prob.solver.launch_count())
