"""The two inner solves of the reference on the C2 workload, side by side: NID_NELDER_MEAD (mode A, trajectory-exact) and
NID_BFGS (mode B value + gradient on K3, Ceres-free BFGS).  Reports time, evaluations and the pose error against the ground
truth of the synthetic scene."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import direct_visual_lidar_calibration_b200 as V
from direct_visual_lidar_calibration_b200 import synthetic as S


def pose_error(A, B):
    d = np.linalg.inv(A) @ B
    return float(np.linalg.norm(d[:3, 3])), float(np.degrees(np.arccos(np.clip(0.5 * (np.trace(d[:3, :3]) - 1.0), -1.0, 1.0))))


n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
bag = S.make_bag("pinhole_1920x1080", "os1_64", n, config_index=1, bag_index=0)
T_gt = S.gt_T_camera_lidar()
T0 = S.perturb(T_gt, (0.5, 0.5, 0.5), (0.02, 0.02, 0.02))
cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
data = [V.VisualLiDARData(bag["image"], bag["points"], bag["intensities"])]
out = {"points": n, "start_error_m_deg": pose_error(T_gt, T0)}
for name, rtype in (("nelder_mead", V.RegistrationType.NID_NELDER_MEAD), ("bfgs", V.RegistrationType.NID_BFGS)):
    p = V.VisualCameraCalibrationParams()
    p.registration_type = rtype
    calib = V.VisualCameraCalibration(cam, data, p)
    solve = calib.estimate_pose_nelder_mead if name == "nelder_mead" else calib.estimate_pose_bfgs
    solve(T0)  # warm (allocations, first launches)
    times = []
    for _ in range(5):
        t0 = time.perf_counter()
        T, r = solve(T0)
        times.append(time.perf_counter() - t0)
    out[name] = {"ms_per_inner_solve_from_host_buffers": round(1e3 * float(np.median(times)), 3), "iterations": r.get("num_iterations", r.get("iterations")),
                 "evaluations": r.get("num_evaluations", r.get("evaluations")), "final_cost": r.get("y", r.get("final_cost")), "termination": r.get("termination", "max_iterations/converged"),
                 "error_m_deg": pose_error(T_gt, T)}
print(json.dumps(out))
