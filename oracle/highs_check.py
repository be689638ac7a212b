"""Independent second opinion on the oracle's QP solve: the HiGHS active-set QP solver that scipy bundles
(private module ``scipy.optimize._highspy._core``; SURVEY.md 8c "independent cross-check available here").

TEST INFRASTRUCTURE ONLY.  Not the oracle (HiGHS stops at its own tolerances, ~1e-6..1e-4 relative here) - a third-party
check that the certified optimum of ``qp_exact`` is the optimum of the QP the reference hands to cvxopt
(MPC.py:288-297): min 1/2 z'Hz + f'z  s.t.  G z <= h, A z = b."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def available() -> bool:
    try:
        from scipy.optimize._highspy import _core  # noqa: F401
        return hasattr(_core, "_Highs") and hasattr(_core, "HighsHessian")
    except Exception:
        return False


def solve_qp(H, f, G, hv, A, b):
    """Solve the dense QP with HiGHS; returns (z, objective, model_status_string)."""
    from scipy.optimize._highspy import _core as hs
    n = H.shape[0]
    rows = sp.vstack([sp.csr_matrix(G), sp.csr_matrix(A)]).tocsc()
    inf = hs.kHighsInf
    model = hs.HighsModel()
    lp = model.lp_
    lp.num_col_, lp.num_row_ = n, rows.shape[0]
    lp.col_cost_ = np.asarray(f, dtype=np.float64)
    lp.col_lower_ = np.full(n, -inf)
    lp.col_upper_ = np.full(n, inf)
    lp.row_lower_ = np.concatenate([np.full(G.shape[0], -inf), np.asarray(b, dtype=np.float64)])
    lp.row_upper_ = np.concatenate([np.asarray(hv, dtype=np.float64), np.asarray(b, dtype=np.float64)])
    lp.a_matrix_.format_ = hs.MatrixFormat.kColwise
    lp.a_matrix_.num_col_, lp.a_matrix_.num_row_ = n, rows.shape[0]
    lp.a_matrix_.start_ = rows.indptr.astype(np.int32)
    lp.a_matrix_.index_ = rows.indices.astype(np.int32)
    lp.a_matrix_.value_ = rows.data.astype(np.float64)
    q = sp.tril(sp.csc_matrix(H)).tocsc()
    hess = model.hessian_
    hess.dim_ = n
    hess.format_ = hs.HessianFormat.kTriangular
    hess.start_ = q.indptr.astype(np.int32)
    hess.index_ = q.indices.astype(np.int32)
    hess.value_ = q.data.astype(np.float64)
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    h.passModel(model)
    h.run()
    sol = h.getSolution()
    return np.array(sol.col_value), float(h.getObjectiveValue()), h.modelStatusToString(h.getModelStatus())
