"""HAND-DERIVED sparsity patterns -- literal CCS arrays written down from CasADi's structural rule, NOT produced by
any code of this repository (neither casadi-lite, nor the oracle, nor tests/indep_models.py).

Rule (CasADi, `SX::jacobian` / `SX::hessian`): an entry (i, j) exists when output i depends on input j in the
expression graph after construction-time simplification; CCS with strictly increasing row indices in a column.
The reference differentiates the AUGMENTED system (src/sqp_solver/SQPOptimizationSolver.cpp:50-62):
    w = [p; x],   c = [p; x; g],   H = hess_w f (full symmetric),   J = dc/dw
so J starts with an n x n identity (n = |p| + |x|) and the rows of g follow.

Known-answer problems of the reference (test/test.cpp:13-185); (h_colptr, h_rowidx, a_colptr, a_rowidx):
"""

KAT = {
    # 1: f = x0^2 + x1^2, g = x0 + x1 - 1.  H = diag.  J = [I2; 1 1] -> column j holds rows {j, 2}.
    1: ([0, 1, 2], [0, 1], [0, 2, 4], [0, 2, 1, 2]),
    # 2: f = (x0-3)^2 + (x1+2)^2, no g.  J = I2.
    2: ([0, 1, 2], [0, 1], [0, 1, 2], [0, 1]),
    # 3: f = (x0-2)^2 + (x1-3)^2, g = x0 + x1 - 1: same structure as case 1.
    3: ([0, 1, 2], [0, 1], [0, 2, 4], [0, 2, 1, 2]),
    # 4: f = x0^2 + x1^2, g = (x0, x1).  J = [I2; I2] -> column j holds rows {j, 2 + j}.
    4: ([0, 1, 2], [0, 1], [0, 2, 4], [0, 2, 1, 3]),
    # 5: three variables, g = x0 + x1 + x2 - 5.  J = [I3; 1 1 1] -> column j holds rows {j, 3}.
    5: ([0, 1, 2, 3], [0, 1, 2], [0, 2, 4, 6], [0, 3, 1, 3, 2, 3]),
    # 6: w = [p, x0, x1], f = (x0 - p)^2 + x1^2: (x0 - p)^2 couples p and x0 in all four combinations.
    #    H columns: p -> {p, x0}, x0 -> {p, x0}, x1 -> {x1}.  No g: J = I3.
    6: ([0, 2, 4, 5], [0, 1, 0, 1, 2], [0, 1, 2, 3], [0, 1, 2]),
    # 7: f = (x0-3)^2 + (x1-4)^2, box only.
    7: ([0, 1, 2], [0, 1], [0, 1, 2], [0, 1]),
}

"""Cart-pole OCP with horizon 2 (problems.cpp: state (s, th, ds, dth), force u, one RK4 step per stage):
    w = [p0..p3 | s0 th0 ds0 dth0 u0 | s1 th1 ds1 dth1 u1]        indices 0-3 | 4-8 | 9-13,   n = 14
    f = sum_k sum_i Q_i (x_ki - p_i)^2 + R u_k^2
    g = x_1 - F(x_0, u_0)  (4 rows, indices 14..17),   m = 18

H: (x_ki - p_i)^2 gives (p_i, p_i), (p_i, x_ki), (x_ki, p_i), (x_ki, x_ki); u_k^2 gives (u_k, u_k).
    column p_i   -> rows {i, 4 + i, 9 + i}
    column x_0i  -> rows {i, 4 + i}          column u_0 -> {8}
    column x_1i  -> rows {i, 9 + i}          column u_1 -> {13}

J: the ODE is  d/dt (s, th, ds, dth) = (ds, dth, dds(th, dth, u), ddth(th, dth, u)):  s never appears on the right-hand
side, ddth depends on (th, dth, u), dds on (th, dth, u) as well (through ddth and temp).  Following the four RK4 stages:
    F_s   = s  + dt/6 (ds + 2 (ds + dt/2 dds) + ...)   depends on  s, ds, th, dth, u
    F_th  = th + dt/6 (dth + 2 (dth + dt/2 ddth) + ...) depends on  th, dth, u          (never on s or ds)
    F_ds  = ds + dt/6 (dds + ...)                       depends on  ds, th, dth, u
    F_dth = dth + dt/6 (ddth + ...)                     depends on  dth, th, u
    row 14 (s):   s0 th0 ds0 dth0 u0 | s1            row 15 (th):  th0 dth0 u0 | th1
    row 16 (ds):  th0 ds0 dth0 u0 | ds1              row 17 (dth): th0 dth0 u0 | dth1
"""
CARTPOLE_H2 = (
    # h_colptr, h_rowidx
    [0, 3, 6, 9, 12, 14, 16, 18, 20, 21, 23, 25, 27, 29, 30],
    [0, 4, 9, 1, 5, 10, 2, 6, 11, 3, 7, 12,
     0, 4, 1, 5, 2, 6, 3, 7, 8,
     0, 9, 1, 10, 2, 11, 3, 12, 13],
    # a_colptr, a_rowidx
    [0, 1, 2, 3, 4, 6, 11, 14, 19, 24, 26, 28, 30, 32, 33],
    [0, 1, 2, 3,
     4, 14,                    # s0
     5, 14, 15, 16, 17,        # th0
     6, 14, 16,                # ds0
     7, 14, 15, 16, 17,        # dth0
     8, 14, 15, 16, 17,        # u0
     9, 14, 10, 15, 11, 16, 12, 17,   # s1, th1, ds1, dth1
     13],                      # u1
)
