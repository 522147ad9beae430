"""Negative control for the guard doubles of -DOCP_B200_CANARY builds: with -DOCP_B200_CANARY_SELFTEST the direct kernel
sizes its w array as it did before the round-2 fix (m entries, although the Ruiz pass writes n column norms into it), and
a QP with m < n must then trip the guard behind that array: expect "OCP_B200 CANARY overwritten ... array 8" on stdout."""
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import optimal_control_problem_b200 as ocp  # noqa: E402

rng = np.random.default_rng(77)
n, m = 14, 5
M = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.4)
S = sp.csc_matrix(M @ M.T + n * np.eye(n)); S.sort_indices()
G = sp.csc_matrix(rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.6)); G.sort_indices()
xf = rng.standard_normal(n)
x, y, info = ocp.cucaqp_solve(n, m, S.indptr, S.indices, S.data, rng.standard_normal(n), G.indptr, G.indices, G.data,
                              G @ xf - 1.0, G @ xf + 1.0, eps_abs=1e-7, eps_rel=1e-7)
print("status", info[0], "iters", info[1], flush=True)
