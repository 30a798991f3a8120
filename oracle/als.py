"""CPU oracle: one weighted-ALS half-sweep in dense float64 (tiny shapes only).  TEST INFRASTRUCTURE.

The reference has no ALS (its wrmf.py is minibatch Adagrad); this restates the normal equations of SURVEY.md
Appendix A ("WRMF (ALS, new)") -- the objective the reference's commented-out weighting (wrmf.py:61-62) and the paper its
README cites (README.md:29) describe:  minimise sum_ui c_ui (r_ui - x_u.y_i)^2 + reg (|X|^2 + |Y|^2), c = weight on
observed pairs, 1 elsewhere.  PARITY UNPINNED (no reference implementation exists); checked for self-consistency: each
half-sweep must not increase the objective."""
import numpy as np


def half_sweep(Y, csr_rows, weight, reg):
    """Returns X [n_x, d]: csr_rows[u] = observed column ids (rows of Y) of row u."""
    Y = np.asarray(Y, dtype=np.float64)
    d = Y.shape[1]
    G = Y.T @ Y
    X = np.zeros((len(csr_rows), d))
    for u, cols in enumerate(csr_rows):
        Yp = Y[np.asarray(list(cols), dtype=np.int64)] if len(cols) else np.zeros((0, d))
        A = G + (weight - 1.0) * (Yp.T @ Yp) + reg * np.eye(d)
        b = weight * Yp.sum(0)
        X[u] = np.linalg.solve(A, b)
    return X


def objective(X, Y, R_dense, weight, reg):
    X, Y = np.asarray(X, np.float64), np.asarray(Y, np.float64)
    P = X @ Y.T
    C = np.where(R_dense > 0, weight, 1.0)
    return float((C * (R_dense - P) ** 2).sum() + reg * ((X ** 2).sum() + (Y ** 2).sum()))
