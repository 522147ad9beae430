"""How does the throughput of admm_direct_kernel scale with the number of resident CTAs per SM?
Runs short-horizon quadrotor batches (their per-CTA shared memory is small enough for 3-4 CTAs/SM) with
each compiled variant of the throughput plan and each residency cap, one subprocess per configuration
(the variant and the cap are read from the environment when the solver handle is created).

  python tools/probe_occupancy.py [--batch 4096] [--horizons 10,6]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def child(horizon: int, batch: int) -> None:
    import numpy as np
    sys.path.insert(0, str(ROOT))
    import optimal_control_problem_b200 as ocp
    prob = ocp.Problem("quadrotor", horizon=horizon, alpha=0.1, step_num=10)
    frames, refs = prob.sample_inputs(batch, 0xB200)
    sol = prob.solver
    x = np.zeros((batch, prob.N)); st = np.zeros((batch, ocp.NSTATS))
    sol.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x, None, st)   # warm-up
    sol.set_profiling(True)
    for _ in range(3):
        x[:] = 0
        sol.solve_batch(frames, refs, prob.lbx, prob.ubx, prob.lbg, prob.ubg, x, None, st)
    prof = sol.get_profile()
    plan = sol.launch_plan()
    print(json.dumps({"horizon": horizon, "variant": os.environ.get("OCP_B200_MULTI_VARIANT", "192x2"),
                      "cap": os.environ.get("OCP_B200_MAX_CTAS_PER_SM", ""),
                      "admm_ms_per_launch": round(prof["admm"]["ms"] / prof["admm"]["launches"], 4),
                      "assemble_ms_per_launch": round(prof["assemble"]["ms"] / prof["assemble"]["launches"], 4),
                      "admm_iters_per_solve": float(st[:, 2].mean()), "wide": plan["wide"], "tri": [plan["tri_bs"], plan["tri_nb"]]}))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--horizons", default="10,6")
    ap.add_argument("--child", type=int, default=0)
    a = ap.parse_args()
    if a.child:
        child(a.child, a.batch)
        return
    for h in [int(v) for v in a.horizons.split(",")]:
        for variant in ("192x2", "192x3", "128x4"):
            for cap in ("1", "2", "3", "4"):
                if int(cap) > int(variant[-1]):
                    continue
                env = dict(os.environ, OCP_B200_MULTI_VARIANT=variant, OCP_B200_MAX_CTAS_PER_SM=cap, OCP_B200_PLAN="multi")
                r = subprocess.run([sys.executable, __file__, "--child", str(h), "--batch", str(a.batch)], env=env,
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-400:]
                print(line, flush=True)


if __name__ == "__main__":
    main()
