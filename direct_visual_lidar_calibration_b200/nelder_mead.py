"""dfo::NelderMead<N> mirror (reference: include/dfo/nelder_mead.hpp) on top of vlcal_nelder_mead_batched."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class NelderMeadParams:
    def __init__(self, init_step=0.1, alpha=1.0, gamma=2.0, rho=0.5, sigma=0.5, max_iterations=1024, convergence_var_thresh=1e-5):  # nelder_mead.hpp:12
        self.init_step, self.alpha, self.gamma, self.rho, self.sigma = init_step, alpha, gamma, rho, sigma
        self.max_iterations, self.convergence_var_thresh = max_iterations, convergence_var_thresh

    def to_c(self) -> _lib.NMParams:
        return _lib.NMParams(self.init_step, self.alpha, self.gamma, self.rho, self.sigma, self.max_iterations, self.convergence_var_thresh)


class NelderMead:
    """optimize(function, x0) with `function` scoring ONE point (reference surface, optimizer.hpp:31-33), or
    optimize_batched(batch_function, x0) with `batch_function(X[count,n]) -> y[count]`.  Both follow the serial
    reference's trajectory exactly; `observed` lists the evaluations the reference would have made, in order."""

    def __init__(self, params: NelderMeadParams | None = None):
        self.params = params or NelderMeadParams()
        self.observed = []

    def optimize(self, function, x0):
        return self.optimize_batched(lambda X: np.array([function(x) for x in X]), x0)

    def optimize_batched(self, batch_function, x0):
        L = _lib.load_library()
        x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64)).reshape(-1)
        n = int(x0.size)
        self.observed = []
        err = []

        def _batch(xs, count, nn, ys, _user):
            try:
                X = np.ctypeslib.as_array(xs, shape=(count, nn)).copy()
                Y = np.asarray(batch_function(X), dtype=np.float64).reshape(count)
                for i in range(count):
                    ys[i] = Y[i]
            except Exception as e:  # never unwind through C
                err.append(e)
                for i in range(count):
                    ys[i] = float("nan")

        def _observe(x, nn, y, _user):
            self.observed.append((np.ctypeslib.as_array(x, shape=(nn,)).copy(), float(y)))

        res = _lib.NMResult()
        p = self.params.to_c()
        _lib.check(L.vlcal_nelder_mead_batched(n, _lib.NM_BATCH_FN(_batch), _lib.NM_OBSERVE_FN(_observe), None, x0.ctypes.data_as(C.POINTER(C.c_double)), C.byref(p), C.byref(res)))
        if err:
            raise err[0]
        return {
            "converged": bool(res.converged), "num_iterations": int(res.num_iterations), "x": np.array(res.x[:n]), "y": float(res.y),
            "num_evaluations": int(res.num_evaluations), "num_batches": int(res.num_batches), "num_evaluations_computed": int(res.num_evaluations_computed),
        }
