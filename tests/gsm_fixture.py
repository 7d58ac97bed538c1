"""A generalised survival model for the tests, built the way rstpm2 builds the list the reference's gsm class reads
(src/gsm.cpp:60-92): natural cubic splines in log time, q_const from the QR decomposition of the second derivatives
at the boundary knots (the construction the reference keeps, commented out, at inst/include/splines.h:56-70)."""
import ctypes as C

import numpy as np


def make_model(oracle, two_terms=True, intercept=0):
    from microsimulation_b200.gsm import GsmModel
    B = np.log([0.1, 50.0])
    knots = np.log([1.0, 5.0, 15.0])
    df = intercept + 3 + len(knots)
    # first pass with a placeholder q_const to get bs::basis(boundary_knots, 2) from the oracle's own spline code
    tmp = GsmModel(0.1, 50.0, [0.0], [dict(gamma=np.zeros(df - 2), knots=knots, Boundary_knots=B, intercept=intercept,
                                           q_const=np.eye(df)[:df - 2], x=[1.0])])
    const_basis = oracle.spline_eval(tmp.cstruct(), 0, B, der=2, natural=False, width=df)  # 2 x df
    q, _ = np.linalg.qr(const_basis.T, mode="complete")
    q_const = np.ascontiguousarray(q[:, 2:].T)  # (df-2) x df
    # gamma: least-squares fit of a smooth increasing log cumulative hazard
    tmp = GsmModel(0.1, 50.0, [0.0], [dict(gamma=np.zeros(df - 2), knots=knots, Boundary_knots=B, intercept=intercept,
                                           q_const=q_const, x=[1.0])])
    yy = np.linspace(B[0], B[1], 200)
    X = oracle.spline_eval(tmp.cstruct(), 0, yy, der=0, natural=True, width=df - 2)
    target = -3.0 + 1.3 * yy + 0.15 * np.sin(yy)
    A = np.c_[np.ones_like(yy), X]
    coef = np.linalg.lstsq(A, target, rcond=None)[0]
    terms = [dict(gamma=coef[1:], knots=knots, Boundary_knots=B, intercept=intercept, q_const=q_const, x=[1.0, 1.0, 1.0])]
    if two_terms:  # a time-varying covariate effect, present for covariate patterns 1 and 2 only
        terms.append(dict(gamma=0.05 * coef[1:], knots=knots, Boundary_knots=B, intercept=intercept, q_const=q_const.T.copy(),
                          x=[0.0, 1.0, -0.5]))
    etap = coef[0] + np.array([0.0, 0.4, -0.3])
    return GsmModel(0.1, 50.0, etap, terms, etap0=etap + 0.2, inflate=1.0, log_time=True, t0=20.0), q_const, knots, B
