"""camera::create_camera / GenericCamera mirror (reference: src/camera/create_camera.cpp, include/camera/*.hpp)."""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a.size else C.POINTER(C.c_double)()


class GenericCamera:
    """camera::GenericCamera<Projection>: holds the model id and the (zero-padded) parameter vectors.

    Unlike the reference's GenericCameraBase (include/camera/generic_camera_base.hpp:29-40) the parameters are
    readable, because the GPU path needs them (SURVEY 8b "camera opacity")."""

    def __init__(self, model: str, model_id: int, intrinsics: np.ndarray, distortion: np.ndarray):
        self.model = model
        self.model_id = model_id
        self.intrinsics = intrinsics
        self.distortion = distortion

    def project(self, point_3d) -> np.ndarray:
        """GenericCameraBase::project (double). Host helper for single points / small arrays."""
        L = _lib.load_library()
        pts = np.ascontiguousarray(np.asarray(point_3d, dtype=np.float64)).reshape(-1, 3)
        out = np.empty((pts.shape[0], 2))
        for i in range(pts.shape[0]):
            _lib.check(L.vlcal_camera_project(self.model_id, _dp(self.intrinsics), self.intrinsics.size, _dp(self.distortion), self.distortion.size, _dp(pts[i]), _dp(out[i])))
        return out if np.asarray(point_3d).ndim > 1 else out[0]

    __call__ = project


def create_camera(camera_model: str, intrinsics, distortion_coeffs):
    """camera::create_camera(model, intrinsics, distortion). Returns None where the reference returns nullptr
    (unknown model: create_camera.cpp:49-50; intrinsic count mismatch: :19-22), printing the same messages."""
    L = _lib.load_library()
    model_id = L.vlcal_camera_model_id(camera_model.encode())
    if model_id < 0:
        print(f"error: unknown camera model {camera_model}", file=sys.stderr)
        return None
    ni, nd = C.c_int(), C.c_int()
    _lib.check(L.vlcal_camera_num_params(model_id, C.byref(ni), C.byref(nd)))
    intr = np.ascontiguousarray(np.asarray(intrinsics, dtype=np.float64)).reshape(-1)
    if intr.size != ni.value:
        print("error: num of intrinsic parameters mismatch!!", file=sys.stderr)
        return None
    src = np.asarray(distortion_coeffs, dtype=np.float64).reshape(-1)
    dist = np.zeros(nd.value)
    k = min(nd.value, src.size)
    dist[:k] = src[:k]  # create_camera.cpp:24-27
    return GenericCamera(camera_model, model_id, intr, dist)
