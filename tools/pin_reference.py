#!/usr/bin/env python3
"""Pins the oracle against the REAL CasADi and OSQP -- to be run on any machine that has them
(`pip install casadi osqp scipy numpy`; neither exists in the build image of this repository).

It rebuilds, with the real libraries, what the reference computes on its CUDA_SQP path and stores it as fixtures
that tests/test_reference_pins.py compares the oracle (and through it the CUDA path) with:

  tests/golden/ref_<problem>_h<H>.npz   for quadrotor / cartpole / centroidal at a small and at the benchmark horizon
      h_colptr, h_rowidx, a_colptr, a_rowidx      CCS patterns of SX::hessian / SX::jacobian of the augmented system
                                                  (src/sqp_solver/SQPOptimizationSolver.cpp:50-62, AutoDifferentiator.cpp:16-27)
      frames, refs, x                              seeded inputs (stored, so the test needs no generator)
      hv, q, av, l, u                              localSystemFunction values at those points (SQPOptimizationSolver.cpp:100-120)
      qp_<eps>_{x, y, iter, status, rho_updates, prim_res, dual_res}
                                                  OSQP on those QPs: eps 1e-3 (what the reference runs, :83-85) and 1e-8,
                                                  adaptive_rho_interval = 100 (the deterministic rule the oracle restates;
                                                  OSQP's default picks the interval from wall-clock time)
      sqp_x, sqp_f, sqp_iters                      the reference's SQP loop (alpha 0.1, 10 steps, cold OSQP set-up per step,
                                                  x += alpha d, SQPOptimizationSolver.cpp:137-181) from x = 0
  tests/golden/ref_kat.npz                       the same for the seven known-answer problems of test/test.cpp (one full step)

The model definitions are the ones of tests/indep_models.py (plain Python over a scalar interface), evaluated here
with casadi.SX.  `--dry-run DIR` exercises the whole script without CasADi / OSQP: patterns from the structural
propagator, values from complex-step differentiation, no QP solves -- that is what the repository's own CPU test runs.

usage: python tools/pin_reference.py [--out tests/golden] [--dry-run DIR] [--problems quadrotor,cartpole,centroidal]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
import indep_models as W  # noqa: E402

HORIZONS = {"quadrotor": (4, 20), "cartpole": (6, 200), "centroidal": (3, 50)}
OSQP_SETTINGS = dict(eps_abs=1e-3, eps_rel=1e-3, max_iter=10000, verbose=False, warm_starting=True, polishing=False,
                     adaptive_rho_interval=100)   # everything else: library defaults, as in the reference


def build_casadi(f_and_g, n_p, N):
    """localSystemFunction of SQPOptimizationSolver.cpp:74-77 with the real CasADi; f_and_g(X, P, M) -> (f, [g])."""
    import casadi as ca
    P = ca.SX.sym("p", n_p)
    Xs = ca.SX.sym("X", N)
    X = [Xs[i] for i in range(N)]
    Pl = [P[i] for i in range(n_p)]
    f, g = f_and_g(X, Pl, ca)
    g = ca.vertcat(*g) if g else ca.SX(0, 1)
    w = ca.vertcat(P, Xs)
    c = ca.vertcat(P, Xs, g)
    H, grad = ca.hessian(f, w)
    J = ca.jacobian(c, w)
    m = c.shape[0]
    l = ca.SX.sym("l", m); u = ca.SX.sym("u", m)
    fn = ca.Function("localSystemFunction", [P, Xs, l, u], [H, grad, J, l + (-c), u + (-c)])
    obj = ca.Function("objective", [P, Xs], [f])
    sp_h, sp_a = fn.sparsity_out(0), fn.sparsity_out(2)
    pats = tuple(np.array(v, np.int32) for v in (sp_h.colind(), sp_h.row(), sp_a.colind(), sp_a.row()))

    def evaluate(p, x, lfull, ufull):
        Hv, q, Av, lo, up = fn(p, x, lfull, ufull)
        return (np.array(Hv.nonzeros()), np.array(q).ravel(), np.array(Av.nonzeros()), np.array(lo).ravel(),
                np.array(up).ravel())

    return pats, evaluate, lambda p, x: float(obj(p, x))


def build_witness(f_and_g, n_p, N):
    """Same interface without CasADi (dry run): structural propagator + complex step through indep_models."""
    P = [W.Dep.var(i) for i in range(n_p)]
    X = [W.Dep.var(n_p + i) for i in range(N)]
    f, g = f_and_g(X, P, W.DepMath)
    pats = W.patterns_from_expressions(f, g, n_p, N)
    return pats, None, None


def osqp_solve(n, m, pats, hv, q, av, l, u, **over):
    """CuCaQP::setSystem/initSolver/solve (CuCaQP.cpp:271-288, 183-211): OsqpEigen hands OSQP the upper triangle."""
    import osqp
    import scipy.sparse as sp
    hc, hr, ac, ar = pats
    Pm = sp.triu(sp.csc_matrix((hv, hr, hc), shape=(n, n)), format="csc")
    Am = sp.csc_matrix((av, ar, ac), shape=(m, n))
    lo, up = np.asarray(l, float), np.asarray(u, float)     # +-inf is accepted (clipped to OSQP_INFTY inside)
    s = osqp.OSQP()
    kw = dict(OSQP_SETTINGS); kw.update(over)
    s.setup(P=Pm, q=q, A=Am, l=lo, u=up, **kw)
    r = s.solve()
    return dict(x=np.array(r.x), y=np.array(r.y), iter=int(r.info.iter), status=int(r.info.status_val),
                rho_updates=int(r.info.rho_updates), prim_res=float(getattr(r.info, "prim_res", getattr(r.info, "pri_res", np.nan))),
                dual_res=float(getattr(r.info, "dual_res", getattr(r.info, "dua_res", np.nan))))


def pin_model(name: str, H: int, out: Path, dry: bool) -> None:
    model = W.MODELS[name]
    n_p, nf, N = model.nx, model.nf, H * model.nf

    def f_and_g(X, P, M):
        g, _, _ = W.constraints(model, H, X, M)
        return W.objective(model, H, X, P), g

    pats, evaluate, objective = (build_witness if dry else build_casadi)(f_and_g, n_p, N)
    lbx, ubx, lbg, ubg = W.model_bounds(name, H)
    n, m = n_p + N, n_p + N + lbg.size
    rng = np.random.default_rng(0xB200 + H)
    B = 2
    # inputs in the spirit of problems.cpp:sample_inputs: a perturbed hover / upright / stance frame held over the horizon
    base = {"quadrotor": np.r_[np.zeros(12), np.full(4, W.QUAD_MASS * W.GRAVITY / 4.0)],
            "cartpole": np.r_[0.0, np.pi, 0.0, 0.0, 0.0],
            "centroidal": np.r_[0.0, 0.0, 0.35, np.zeros(9), 0.25, 0.15, 0.0, 0.25, -0.15, 0.0, -0.25, 0.15, 0.0, -0.25, -0.15, 0.0,
                                np.tile([0.0, 0.0, W.LEG_MASS * W.GRAVITY / 4.0], 4)]}[name]
    frames = base[None, :] + 0.1 * rng.standard_normal((B, nf)) * (np.arange(nf) < model.nx)
    refs = np.zeros((B, n_p)) if name != "centroidal" else np.tile(base[:n_p], (B, 1))
    x = np.tile(frames, (1, H)) + 0.05 * rng.standard_normal((B, N))
    data = dict(h_colptr=pats[0], h_rowidx=pats[1], a_colptr=pats[2], a_rowidx=pats[3], frames=frames, refs=refs, x=x,
                lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, horizon=H, backend="witness" if dry else "casadi+osqp")
    if not dry:
        import casadi
        import osqp
        data["versions"] = np.array([f"casadi {casadi.__version__}", f"osqp {osqp.__version__}"])
        vals = {k: [] for k in ("hv", "q", "av", "l", "u")}
        qp = {}
        for b in range(B):
            lb = lbx.copy(); ub = ubx.copy()
            lb[:nf] = frames[b]; ub[:nf] = frames[b]                      # OptimalControlProblem.cpp:93-96
            lfull = np.r_[refs[b], lb, lbg]; ufull = np.r_[refs[b], ub, ubg]
            hv, q, av, l, u = evaluate(refs[b], x[b], lfull, ufull)
            for k, v in zip(("hv", "q", "av", "l", "u"), (hv, q, av, l, u)):
                vals[k].append(v)
            for tag, eps in (("1e-3", 1e-3), ("1e-8", 1e-8)):
                r = osqp_solve(n, m, pats, hv, q, av, l, u, eps_abs=eps, eps_rel=eps)
                for k, v in r.items():
                    qp.setdefault(f"qp_{tag}_{k}", []).append(v)
        data.update({k: np.array(v) for k, v in vals.items()})
        data.update({k: np.array(v) for k, v in qp.items()})
        # the reference's SQP loop from x = 0 (SQPOptimizationSolver.cpp:137-181), alpha 0.1, 10 steps
        sx, sf, sit = [], [], []
        for b in range(B):
            lb = lbx.copy(); ub = ubx.copy()
            lb[:nf] = frames[b]; ub[:nf] = frames[b]
            lfull = np.r_[refs[b], lb, lbg]; ufull = np.r_[refs[b], ub, ubg]
            xi = np.zeros(N); iters = 0
            for _ in range(10):
                hv, q, av, l, u = evaluate(refs[b], xi, lfull, ufull)
                r = osqp_solve(n, m, pats, hv, q, av, l, u)
                xi = xi + 0.1 * r["x"][n_p:]
                iters += r["iter"]
            sx.append(xi); sf.append(objective(refs[b], xi)); sit.append(iters)
        data.update(sqp_x=np.array(sx), sqp_f=np.array(sf), sqp_iters=np.array(sit))
    np.savez_compressed(out / f"ref_{name}_h{H}.npz", **data)
    print(f"wrote {out / f'ref_{name}_h{H}.npz'}: n={n} m={m} nnz_h={pats[1].size} nnz_a={pats[3].size}")


def pin_kat(out: Path, dry: bool) -> None:
    INF = float(np.float32(np.inf))
    args = {1: dict(lbx=[-50, -100], ubx=[50, 100], lbg=[0.0], ubg=[0.0], p=[]),
            2: dict(lbx=[-50, -100], ubx=[50, 100], lbg=[], ubg=[], p=[]),
            3: dict(lbx=[-100, -100], ubx=[100, 100], lbg=[1.0], ubg=[INF], p=[]),
            4: dict(lbx=[-100, -100], ubx=[100, 100], lbg=[1.0, 2.0], ubg=[INF, INF], p=[]),
            5: dict(lbx=[0, 0, 0], ubx=[INF, INF, INF], lbg=[0.0], ubg=[0.0], p=[]),
            6: dict(lbx=[-100, -100], ubx=[100, 100], lbg=[], ubg=[], p=[5.0]),
            7: dict(lbx=[0, 0], ubx=[2, 3], lbg=[], ubg=[], p=[])}            # test/test.cpp:13-185
    data = {}
    for case in range(1, 8):
        nx, n_p = W.KAT_DIMS[case]
        pats, evaluate, objective = (build_witness if dry else build_casadi)(lambda X, P, M: W.kat_problem(case, X, P), n_p, nx)
        for k, v in zip(("h_colptr", "h_rowidx", "a_colptr", "a_rowidx"), pats):
            data[f"k{case}_{k}"] = v
        if dry:
            continue
        a = {k: np.array(v, float) for k, v in args[case].items()}
        lfull = np.r_[a["p"], a["lbx"], a["lbg"]]; ufull = np.r_[a["p"], a["ubx"], a["ubg"]]
        hv, q, av, l, u = evaluate(a["p"], np.zeros(nx), lfull, ufull)
        r = osqp_solve(n_p + nx, lfull.size, pats, hv, q, av, l, u, eps_abs=1e-8, eps_rel=1e-8)
        data[f"k{case}_x"] = r["x"][n_p:]          # one full step from x = 0 solves these QPs
        data[f"k{case}_iter"] = r["iter"]; data[f"k{case}_status"] = r["status"]
    np.savez_compressed(out / "ref_kat.npz", backend="witness" if dry else "casadi+osqp", **data)
    print("wrote", out / "ref_kat.npz")


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    ap.add_argument("--dry-run", default=None, metavar="DIR", help="no CasADi / OSQP: structural patterns only, written to DIR")
    ap.add_argument("--problems", default="quadrotor,cartpole,centroidal")
    ap.add_argument("--small-only", action="store_true", help="skip the benchmark horizons (minutes of CasADi set-up)")
    a = ap.parse_args()
    out = Path(a.dry_run or a.out)
    out.mkdir(parents=True, exist_ok=True)
    dry = a.dry_run is not None
    pin_kat(out, dry)
    for name in a.problems.split(","):
        for H in HORIZONS[name][:1 if a.small_only else 2]:
            pin_model(name, H, out, dry)


if __name__ == "__main__":
    main()
