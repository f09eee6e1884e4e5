"""compute-sanitizer workload for the mode-B kernels (K2 value, K3 value + gradient) and the BFGS driver: small inputs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import direct_visual_lidar_calibration_b200 as V
import util
from scipy.spatial.transform import Rotation

for model in util.MODELS:
    pr = util.random_problem(model, n=3000, seed=5)
    T = pr["T"]
    tp = np.concatenate([Rotation.from_matrix(T[:3, :3]).as_quat(), T[:3, 3]])[None]
    cam = V.create_camera(model, pr["intrinsics"], pr["distortion"])
    for bins in (16, 8):
        cost = V.NIDCost(cam, V.VisualLiDARData(pr["image"], pr["points"], pr["intensities"]), bins)
        ok, nid = cost.evaluate(tp)
        ok2, nid2, grad = cost.evaluate_with_gradient(tp)
        assert np.isfinite(nid).all() and np.isfinite(grad).all() and abs(nid[0] - nid2[0]) < 1e-9
        cost.close()
pr = util.random_problem("plumb_bob", n=3000, seed=6)
cam = V.create_camera("plumb_bob", pr["intrinsics"], pr["distortion"])
params = V.VisualCameraCalibrationParams()
params.registration_type = V.RegistrationType.NID_BFGS
calib = V.VisualCameraCalibration(cam, [V.VisualLiDARData(pr["image"], pr["points"], pr["intensities"])], params)
T, r = calib.estimate_pose_bfgs(pr["T"])
assert np.isfinite(T).all()
print("SANITIZE_MODEB_OK", r["iterations"], r["termination"])
