"""INDEPENDENT WITNESS -- test infrastructure.  Shares no code with casadi-lite, the oracle or the product.

The three benchmark OCPs (optimal_control_problem_b200/problems/problems.cpp) and the seven known-answer
problems of the reference's test/test.cpp, written once more in plain Python against a minimal scalar
interface (+ - * /, sin, cos, tan).  The same model code is run by three interpreters:

  Dep      structural (boolean) dependency propagation.  Gives the CCS patterns of J = dc/dw and of
           H = hess_w f for the augmented system w = [p; x], c = [p; x; g] that the reference builds at
           src/sqp_solver/SQPOptimizationSolver.cpp:50-62 through AutoDifferentiator.cpp:16-27.  CasADi's
           patterns are structural: an entry exists when the output depends on the input in the expression
           graph after construction-time simplification (x * 0 -> 0, x + 0 -> x, x - 0 -> x, 0 / x -> 0,
           constant folding); rows strictly increasing inside a column.  The Hessian rule is the one
           dependency propagation through the gradient expressions amounts to: linear operations pass
           second-order structure on, a product adds D(a) x D(b), a non-linear function adds D(a) x D(a).
  float / complex numbers
           values; first derivatives by complex-step differentiation (exact to rounding, no AD code);
           the cost is a weighted sum of squares, its gradient and Hessian are written out by hand.
  casadi.SX (tools/pin_reference.py, on a machine that has CasADi)
           the real thing, for the fixtures tests/golden/ref_*.npz.
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass, field

import numpy as np

GRAVITY = 9.81


# ---------------------------------------------------------------------------------------------
# interpreter 1: structural dependencies
# ---------------------------------------------------------------------------------------------
class Dep:
    """A scalar expression reduced to: constant value (or None), first-order dependency set d, second-order
    structure h (set of index pairs (i, j) with i <= j)."""
    __slots__ = ("c", "d", "h")

    def __init__(self, c=None, d=frozenset(), h=frozenset()):
        self.c, self.d, self.h = c, d, h

    @staticmethod
    def var(i: int) -> "Dep":
        return Dep(None, frozenset([i]), frozenset())

    @staticmethod
    def lift(v) -> "Dep":
        return v if isinstance(v, Dep) else Dep(float(v))

    @staticmethod
    def _cross(a: frozenset, b: frozenset) -> frozenset:
        return frozenset((min(i, j), max(i, j)) for i in a for j in b)

    def _lin(self, o: "Dep", sign: float) -> "Dep":
        o = Dep.lift(o)
        if self.c is not None and o.c is not None:
            return Dep(self.c + sign * o.c)
        if o.c is not None and o.c == 0.0:      # x + 0, x - 0
            return self
        if self.c is not None and self.c == 0.0 and sign > 0:   # 0 + x
            return o
        return Dep(None, self.d | o.d, self.h | o.h)   # (0 - x keeps x's structure as well)

    def __add__(self, o): return self._lin(o, 1.0)
    def __radd__(self, o): return Dep.lift(o)._lin(self, 1.0)
    def __sub__(self, o): return self._lin(o, -1.0)
    def __rsub__(self, o): return Dep.lift(o)._lin(self, -1.0)
    def __neg__(self): return Dep(-self.c) if self.c is not None else Dep(None, self.d, self.h)

    def __mul__(self, o):
        o = Dep.lift(o)
        if self.c is not None and o.c is not None:
            return Dep(self.c * o.c)
        if (self.c is not None and self.c == 0.0) or (o.c is not None and o.c == 0.0):
            return Dep(0.0)                      # 0 * x -> 0
        if self.c is not None:
            return Dep(None, o.d, o.h)
        if o.c is not None:
            return Dep(None, self.d, self.h)
        return Dep(None, self.d | o.d, self.h | o.h | Dep._cross(self.d, o.d))

    def __rmul__(self, o): return Dep.lift(o).__mul__(self)

    def __truediv__(self, o):
        o = Dep.lift(o)
        if self.c is not None and o.c is not None:
            return Dep(self.c / o.c)
        if self.c is not None and self.c == 0.0:
            return Dep(0.0)                      # 0 / x -> 0
        if o.c is not None:
            return Dep(None, self.d, self.h)     # division by a constant is linear
        return Dep(None, self.d | o.d, self.h | o.h | Dep._cross(self.d, o.d) | Dep._cross(o.d, o.d))

    def __rtruediv__(self, o): return Dep.lift(o).__truediv__(self)

    def nonlinear(self, fn) -> "Dep":
        if self.c is not None:
            return Dep(fn(self.c))
        return Dep(None, self.d, self.h | Dep._cross(self.d, self.d))


class DepMath:
    sin = staticmethod(lambda a: Dep.lift(a).nonlinear(math.sin))
    cos = staticmethod(lambda a: Dep.lift(a).nonlinear(math.cos))
    tan = staticmethod(lambda a: Dep.lift(a).nonlinear(math.tan))


class NumMath:
    """floats and complex numbers (complex-step differentiation)"""
    sin = staticmethod(lambda a: cmath.sin(a) if isinstance(a, complex) else math.sin(a))
    cos = staticmethod(lambda a: cmath.cos(a) if isinstance(a, complex) else math.cos(a))
    tan = staticmethod(lambda a: cmath.tan(a) if isinstance(a, complex) else math.tan(a))


# ---------------------------------------------------------------------------------------------
# the models: lists of scalars in, lists of scalars out; M supplies sin / cos / tan
# ---------------------------------------------------------------------------------------------
def rk4(ode, x, u, dt, M):
    def axpy(a, xs, ys):          # ys + a * xs, written as problems.cpp writes it: x + (a) * k
        return [y + a * k for y, k in zip(ys, xs)]
    k1 = ode(x, u, M)
    k2 = ode(axpy(0.5 * dt, k1, x), u, M)
    k3 = ode(axpy(0.5 * dt, k2, x), u, M)
    k4 = ode(axpy(dt, k3, x), u, M)
    return [xi + (dt / 6.0) * (a + 2.0 * b + 2.0 * c + d) for xi, a, b, c, d in zip(x, k1, k2, k3, k4)]


def euler_rates(rpy, w, M):
    sr, cr, tp, cp = M.sin(rpy[0]), M.cos(rpy[0]), M.tan(rpy[1]), M.cos(rpy[1])
    return [w[0] + sr * tp * w[1] + cr * tp * w[2], cr * w[1] - sr * w[2], (sr / cp) * w[1] + (cr / cp) * w[2]]


QUAD_MASS, QUAD_ARM, QUAD_KAPPA = 1.0, 0.17, 0.016
QUAD_J = (0.01, 0.01, 0.02)


def quadrotor_ode(x, u, M):
    rpy, v, w = x[3:6], x[6:9], x[9:12]
    sr, cr, sp, cp, sy, cy = M.sin(rpy[0]), M.cos(rpy[0]), M.sin(rpy[1]), M.cos(rpy[1]), M.sin(rpy[2]), M.cos(rpy[2])
    T = u[0] + u[1] + u[2] + u[3]
    ax = (cy * sp * cr + sy * sr) * T / QUAD_MASS
    ay = (sy * sp * cr - cy * sr) * T / QUAD_MASS
    az = (cp * cr) * T / QUAD_MASS - GRAVITY
    tx, ty, tz = QUAD_ARM * (u[1] - u[3]), QUAD_ARM * (u[2] - u[0]), QUAD_KAPPA * (u[0] - u[1] + u[2] - u[3])
    wx, wy, wz = w
    dwx = (tx - (QUAD_J[2] - QUAD_J[1]) * wy * wz) / QUAD_J[0]
    dwy = (ty - (QUAD_J[0] - QUAD_J[2]) * wz * wx) / QUAD_J[1]
    dwz = (tz - (QUAD_J[1] - QUAD_J[0]) * wx * wy) / QUAD_J[2]     # J[1] == J[0]: the gyroscopic term is 0 * wx * wy
    return list(v) + euler_rates(rpy, w, M) + [ax, ay, az, dwx, dwy, dwz]


CART_M, POLE_M, POLE_L = 1.0, 0.1, 0.5


def cartpole_ode(x, u, M):
    th, ds, dth = x[1], x[2], x[3]
    s, c = M.sin(th), M.cos(th)
    total = CART_M + POLE_M
    temp = (u[0] + POLE_M * POLE_L * dth * dth * s) / total
    ddth = (GRAVITY * s - c * temp) / (POLE_L * (4.0 / 3.0 - POLE_M * c * c / total))
    dds = temp - POLE_M * POLE_L * ddth * c / total
    return [ds, dth, dds, ddth]


LEG_MASS, LEG_MU = 12.0, 0.6
LEG_IINV = (1.0 / 0.25, 1.0 / 0.5, 1.0 / 0.6)


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def centroidal_ode(x, u, M):
    com, rpy, lin, ang = x[0:3], x[3:6], x[6:9], x[9:12]
    omega = [LEG_IINV[i] * ang[i] for i in range(3)]
    dlin = [0.0, 0.0, -LEG_MASS * GRAVITY]
    dang = [0.0, 0.0, 0.0]
    for i in range(4):
        f = u[3 * i:3 * i + 3]
        r = [x[12 + 3 * i + a] - com[a] for a in range(3)]
        dlin = [d + fi for d, fi in zip(dlin, f)]
        dang = [d + ci for d, ci in zip(dang, cross(r, f))]
    return [li / LEG_MASS for li in lin] + euler_rates(rpy, omega, M) + dlin + dang + [0.0] * 12


@dataclass
class Model:
    name: str
    nx: int
    nu: int
    dt: float
    horizon: int
    ode: object
    Q: list
    R: list
    unom: list
    lb: list            # frame bounds [x; u]
    ub: list
    friction: bool = False
    extra: dict = field(default_factory=dict)

    @property
    def nf(self): return self.nx + self.nu


INF = math.inf
MODELS = {
    "quadrotor": Model("quadrotor", 12, 4, 0.005, 20, quadrotor_ode,
                       [10, 10, 10, 5, 5, 5, 1, 1, 1, 0.5, 0.5, 0.5], [0.1] * 4, [QUAD_MASS * GRAVITY / 4.0] * 4,
                       [-INF] * 3 + [-0.8, -0.8, -INF] + [-INF] * 6 + [0.0] * 4,
                       [INF] * 3 + [0.8, 0.8, INF] + [INF] * 6 + [4.905] * 4),
    "cartpole": Model("cartpole", 4, 1, 0.01, 200, cartpole_ode, [1.0, 10.0, 0.1, 0.1], [0.01], [0.0],
                      [-2.4, -INF, -INF, -INF, -20.0], [2.4, INF, INF, INF, 20.0]),
    "centroidal": Model("centroidal", 24, 12, 0.01, 50, centroidal_ode,
                        [50, 50, 100, 20, 20, 10, 1, 1, 1, 2, 2, 2] + [100.0] * 12, [1e-3] * 12,
                        [0.0, 0.0, LEG_MASS * GRAVITY / 4.0] * 4,
                        [-INF] * 36, [INF] * 24 + [INF, INF, 200.0] * 4, friction=True),
}


def constraints(model: Model, H: int, X, M):
    """g as the reference assembles it: x_{k+1} - F(x_k, u_k) for k = 0..H-2 (addEquationConstraint: lhs - rhs
    in [0, 0], OptimalControlProblem.cpp:476-481), then -- centroidal -- 5 friction rows per foot and stage."""
    nf, nx = model.nf, model.nx
    g, lbg, ubg = [], [], []
    for k in range(H - 1):
        xk, uk = X[k * nf:k * nf + nx], X[k * nf + nx:(k + 1) * nf]
        F = rk4(model.ode, xk, uk, model.dt, M)
        g += [X[(k + 1) * nf + i] - F[i] for i in range(nx)]
        lbg += [0.0] * nx
        ubg += [0.0] * nx
    if model.friction:
        for k in range(H):
            uk = X[k * nf + nx:(k + 1) * nf]
            for i in range(4):
                fx, fy, fz = uk[3 * i], uk[3 * i + 1], uk[3 * i + 2]
                g += [fx - LEG_MU * fz, -fx - LEG_MU * fz, fy - LEG_MU * fz, -fy - LEG_MU * fz, fz]
                lbg += [-INF, -INF, -INF, -INF, 0.0]
                ubg += [0.0, 0.0, 0.0, 0.0, INF]
    return g, lbg, ubg


def objective(model: Model, H: int, X, P):
    """sum_k sum_i Q_i (x_ki - p_i)^2 + sum_j R_j (u_kj - unom_j)^2 (addVectorCost: sum w_i e_i^2,
    OptimalControlProblem.cpp:580-585)."""
    nf, nx = model.nf, model.nx
    f = 0.0
    for k in range(H):
        for i in range(nx):
            e = X[k * nf + i] - P[i]
            f = f + model.Q[i] * e * e
        for j in range(model.nu):
            e = X[k * nf + nx + j] - model.unom[j]
            f = f + model.R[j] * e * e
    return f


# ---------------------------------------------------------------------------------------------
# patterns of the augmented local system
# ---------------------------------------------------------------------------------------------
def _ccs_from_columns(cols, ncol):
    colptr, rowidx = [0], []
    for j in range(ncol):
        rowidx += sorted(cols[j])
        colptr.append(len(rowidx))
    return np.array(colptr, np.int32), np.array(rowidx, np.int32)


def patterns_from_expressions(f: Dep, g: list, n_p: int, N: int):
    """CCS of H = hess_w f (full symmetric) and J = d[p; x; g]/dw with w = [p; x] (variable ids 0..n-1)."""
    n = n_p + N
    hcols = [set() for _ in range(n)]
    for (i, j) in Dep.lift(f).h:
        hcols[j].add(i)
        hcols[i].add(j)
    acols = [set([j]) for j in range(n)]           # identity rows of c = [p; x; ...]
    for r, gr in enumerate(g):
        for j in Dep.lift(gr).d:
            acols[j].add(n + r)
    return _ccs_from_columns(hcols, n) + _ccs_from_columns(acols, n)


def model_patterns(name: str, H: int | None = None):
    model = MODELS[name]
    H = H or model.horizon
    n_p, N = model.nx, H * model.nf
    P = [Dep.var(i) for i in range(n_p)]
    X = [Dep.var(n_p + i) for i in range(N)]
    g, _, _ = constraints(model, H, X, DepMath)
    f = objective(model, H, X, P)
    return patterns_from_expressions(f, g, n_p, N)


# ---------------------------------------------------------------------------------------------
# values of the augmented local system (SQPOptimizationSolver.cpp:100-120)
# ---------------------------------------------------------------------------------------------
def model_bounds(name: str, H: int | None = None):
    model = MODELS[name]
    H = H or model.horizon
    _, lbg, ubg = constraints(model, H, [0.0] * (H * model.nf), NumMath)
    return (np.tile(np.array(model.lb, float), H), np.tile(np.array(model.ub, float), H), np.array(lbg, float),
            np.array(ubg, float))


def local_system_dense(name: str, H: int, p, x, frame=None):
    """(H, grad, J, l', u') as dense numpy arrays for w = [p; x]: gradient / Hessian of the weighted sum of squares by
    hand, dg/dw by complex-step differentiation of the stage defects, l' = l - c, u' = u - c with l = [p; lbx; lbg],
    u = [p; ubx; ubg] and the first frame pinned (OptimalControlProblem.cpp:93-96) when `frame` is given."""
    model = MODELS[name]
    nf, nx, nu = model.nf, model.nx, model.nu
    n_p, N = nx, H * nf
    n = n_p + N
    p = np.asarray(p, float); x = np.asarray(x, float)
    g, lbg, ubg = constraints(model, H, list(x), NumMath)
    g = np.array(g, float)
    m = n + g.size
    Hm = np.zeros((n, n)); grad = np.zeros(n); J = np.zeros((m, n))
    J[:n, :n] = np.eye(n)
    for k in range(H):
        for i in range(nx):
            xi = n_p + k * nf + i
            e = x[k * nf + i] - p[i]
            grad[xi] += 2.0 * model.Q[i] * e
            grad[i] -= 2.0 * model.Q[i] * e
            Hm[xi, xi] += 2.0 * model.Q[i]; Hm[i, i] += 2.0 * model.Q[i]
            Hm[xi, i] -= 2.0 * model.Q[i]; Hm[i, xi] -= 2.0 * model.Q[i]
        for j in range(nu):
            uj = n_p + k * nf + nx + j
            grad[uj] += 2.0 * model.R[j] * (x[k * nf + nx + j] - model.unom[j])
            Hm[uj, uj] += 2.0 * model.R[j]
    h = 1e-30
    for k in range(H - 1):                           # defect k = x_{k+1} - F(x_k, u_k)
        base = list(x[k * nf:(k + 1) * nf])
        for c in range(nf):
            z = [complex(v) for v in base]
            z[c] += 1j * h
            F = rk4(model.ode, z[:nx], z[nx:], model.dt, NumMath)
            for i in range(nx):
                J[n + k * nx + i, n_p + k * nf + c] = -(F[i].imag / h)
        for i in range(nx):
            J[n + k * nx + i, n_p + (k + 1) * nf + i] = 1.0
    if model.friction:
        r0 = n + (H - 1) * nx
        for k in range(H):
            for i in range(4):
                c0 = n_p + k * nf + nx + 3 * i
                rows = r0 + (k * 4 + i) * 5
                J[rows + 0, c0] = 1.0; J[rows + 0, c0 + 2] = -LEG_MU
                J[rows + 1, c0] = -1.0; J[rows + 1, c0 + 2] = -LEG_MU
                J[rows + 2, c0 + 1] = 1.0; J[rows + 2, c0 + 2] = -LEG_MU
                J[rows + 3, c0 + 1] = -1.0; J[rows + 3, c0 + 2] = -LEG_MU
                J[rows + 4, c0 + 2] = 1.0
    lbx = np.tile(np.array(model.lb, float), H); ubx = np.tile(np.array(model.ub, float), H)
    if frame is not None:
        lbx[:nf] = frame; ubx[:nf] = frame
    c = np.concatenate([p, x, g])
    lfull = np.concatenate([p, lbx, np.array(lbg, float)]); ufull = np.concatenate([p, ubx, np.array(ubg, float)])
    return Hm, grad, J, lfull - c, ufull - c


def objective_value(name: str, H: int, p, x) -> float:
    return float(objective(MODELS[name], H, list(np.asarray(x, float)), list(np.asarray(p, float))))


# ---------------------------------------------------------------------------------------------
# the reference's known-answer problems (test/test.cpp:13-185): f and g over x (and p)
# ---------------------------------------------------------------------------------------------
def kat_problem(case: int, X, P):
    sq = lambda a: a * a
    if case == 1: return sq(X[0]) + sq(X[1]), [X[0] + X[1] - 1]
    if case == 2: return sq(X[0] - 3) + sq(X[1] + 2), []
    if case == 3: return sq(X[0] - 2) + sq(X[1] - 3), [X[0] + X[1] - 1]
    if case == 4: return sq(X[0]) + sq(X[1]), [X[0], X[1]]
    if case == 5: return sq(X[0] - 1) + sq(X[1] - 2) + sq(X[2] - 3), [X[0] + X[1] + X[2] - 5]
    if case == 6: return sq(X[0] - P[0]) + sq(X[1]), []
    if case == 7: return sq(X[0] - 3) + sq(X[1] - 4), []
    raise ValueError(case)


KAT_DIMS = {1: (2, 0), 2: (2, 0), 3: (2, 0), 4: (2, 0), 5: (3, 0), 6: (2, 1), 7: (2, 0)}   # (|x|, |p|)


def kat_patterns(case: int):
    nx, n_p = KAT_DIMS[case]
    P = [Dep.var(i) for i in range(n_p)]
    X = [Dep.var(n_p + i) for i in range(nx)]
    f, g = kat_problem(case, X, P)
    return patterns_from_expressions(f, g, n_p, nx)
