"""Times the generated assembly kernel alone (export_qp path is host-bound; use profiling of solve with 1 step).
usage: OCP_B200_ASSEMBLE_MIN_BLOCKS=k python tools/assemble_bench.py [B]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import optimal_control_problem_b200 as ocp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prob = ocp.Problem("quadrotor")
frames, refs = prob.sample_inputs(B, 3)
sol = prob.solver
x = np.zeros((B, prob.N))
sol.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x)
sol.set_profiling(True); sol.get_profile()
x[:] = 0
sol.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x)
pr = sol.get_profile()
print(prob.model_library.split("/")[-1], "assemble ms/launch", pr["assemble"]["ms"] / pr["assemble"]["launches"], "admm ms/launch", pr["admm"]["ms"] / pr["admm"]["launches"])
