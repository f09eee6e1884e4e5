"""The reference's NID_BFGS branch (VisualCameraCalibration::estimate_pose_bfgs, visual_camera_calibration.cpp:187-238):
min over SE(3) of sum_bags NIDCost, solved with a line-search BFGS on the manifold.  The reference uses
ceres::GradientProblemSolver; this is the Ceres-free solver of csrc/bfgs.cu (same problem, Ceres' documented defaults,
iterates not claimed identical)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .camera import _dp
from .cost import T_to_colmajor, colmajor_to_T

TERMINATION = {0: "no_convergence", 1: "gradient_tolerance", 2: "function_tolerance", 3: "parameter_tolerance", 4: "line_search_failed", 5: "failure"}


def default_bfgs_params() -> _lib.BfgsParams:
    p = _lib.BfgsParams()
    _lib.load_library().vlcal_bfgs_default_params(C.byref(p))
    return p


def _result_dict(r: _lib.BfgsResult):
    return {
        "iterations": int(r.iterations), "evaluations": int(r.evaluations), "termination": TERMINATION.get(int(r.termination), "?"),
        "line_search_restarts": int(r.line_search_restarts), "initial_cost": float(r.initial_cost), "final_cost": float(r.final_cost),
        "gradient_max_norm": float(r.gradient_max_norm),
    }


def _pose_callback(callback):
    def _cb(Tp, cost, _user):
        if callback:
            callback(colmajor_to_T(np.ctypeslib.as_array(Tp, shape=(16,)))[0], float(cost))

    return _lib.POSE_CALLBACK(_cb)


def minimize_se3(objective, init_T, params: _lib.BfgsParams | None = None, callback=None):
    """BFGS over SE(3) of a Python objective: objective(x7) -> (ok, cost, grad7) with x7 = [qx qy qz qw tx ty tz] and grad7
    the ambient gradient.  Host only (no GPU needed).  Returns (T[4, 4], result dict)."""
    L = _lib.load_library()

    def _obj(xp, costp, gradp, _user):
        ok, cost, grad = objective(np.array([xp[i] for i in range(7)]))
        costp[0] = float(cost)
        for i in range(7):
            gradp[i] = float(grad[i])
        return 1 if ok else 0

    p = params or default_bfgs_params()
    out = np.empty(16)
    res = _lib.BfgsResult()
    _lib.check(L.vlcal_bfgs_minimize_se3(_lib.SE3_OBJECTIVE(_obj), None, C.byref(p), _dp(T_to_colmajor(init_T)), _pose_callback(callback), None, _dp(out), C.byref(res)))
    return colmajor_to_T(out)[0], _result_dict(res)


def estimate_pose_bfgs_on_costs(costs, init_T_camera_lidar, params: _lib.BfgsParams | None = None, callback=None, allreduce=None):
    """BFGS over already-built mode-B cost objects (NIDCost, one per bag; the `nid_costs` vector of :198-206).
    allreduce: optional callable(np.ndarray[9]) summing in place over all ranks (bags sharded, one process per GPU)."""
    L = _lib.load_library()

    def _allreduce(vals, count, _user):
        allreduce(np.ctypeslib.as_array(vals, shape=(count,)))

    ar = _lib.ALLREDUCE_FN(_allreduce) if allreduce else _lib.ALLREDUCE_FN()
    handles = (C.c_void_p * len(costs))(*[c.handle for c in costs])
    p = params or default_bfgs_params()
    out = np.empty(16)
    res = _lib.BfgsResult()
    _lib.check(L.vlcal_estimate_pose_bfgs_ctx(handles, len(costs), C.byref(p), _dp(T_to_colmajor(init_T_camera_lidar)), _pose_callback(callback), ar, None, _dp(out), C.byref(res)))
    return colmajor_to_T(out)[0], _result_dict(res)
