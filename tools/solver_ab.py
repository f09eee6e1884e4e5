"""A/B of the inner-solve loops on the C2 workload, profiling off: wall time per solve, host loop vs device-resident loop."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import direct_visual_lidar_calibration_b200 as V
from direct_visual_lidar_calibration_b200 import calibration as VC
from direct_visual_lidar_calibration_b200 import synthetic as S

bag = S.make_bag("pinhole_1920x1080", "os1_64", 1_000_000, config_index=1)
T_init = S.perturb(S.gt_T_camera_lidar(), (0.5, 0.5, 0.5), (0.02, 0.02, 0.02))
cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
idx = V.ViewCulling(cam, (bag["width"], bag["height"])).cull_indices(bag["points"], T_init)
cost = V.CostCalculatorNID(cam, V.VisualLiDARData(bag["image"], bag["points"][idx], bag["intensities"][idx]))
cost.reorder_for_pose(T_init)
params = V.VisualCameraCalibrationParams()
for mode, name in ((1, "host"), (2, "device"), (1, "host"), (2, "device")):
    V.set_solver_mode(mode)
    for _ in range(3):
        VC.estimate_pose_on_costs([cost], T_init, params)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        T, r = VC.estimate_pose_on_costs([cost], T_init, params)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps({"solver": name, "ms_per_solve": round(dt * 1e3, 3), "batches": r["num_batches"], "us_per_batch": round(dt * 1e6 / r["num_batches"], 2), "evals_per_s": round(r["num_evaluations"] / dt), "y": r["y"]}), flush=True)
