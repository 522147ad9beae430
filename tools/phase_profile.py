"""Per-phase SM-cycle breakdown of the direct ADMM kernel (CTA 0), for DESIGN.md / profiles/.
The solve_* / factor_* sub-phases are only counted when libocp_b200.so is built with -DOCP_B200_FINE_PHASES
(OCP_B200_NVCC_EXTRA="-DOCP_B200_FINE_PHASES" python -c "from optimal_control_problem_b200 import _build; _build.build_cuda(force=True)").
usage: python tools/phase_profile.py [problem] [B]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import optimal_control_problem_b200 as ocp  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
prob = ocp.Problem(name)
frames, refs = prob.sample_inputs(B, 1)
sol = prob.solver
x = np.zeros((B, prob.N)); st = np.zeros((B, ocp.NSTATS))
sol.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x, None, st)
sol.set_profiling(True, phases=True)
x[:] = 0
t = time.time()
sol.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x, None, st)
dt = time.time() - t
print(name, "B", B, "wall ms", round(dt * 1e3, 2), sol.get_profile(), sol.device_dims())
ph = sol.get_phase_cycles()
tot = max(1, sum(v for k, v in ph.items() if not k.startswith(("solve_", "factor_"))))
print({k: (v, round(100 * v / tot, 1)) for k, v in ph.items()}, "total cycles", tot, "admm iters inst0", st[0, 2],
      "qps", st[0, 1])
