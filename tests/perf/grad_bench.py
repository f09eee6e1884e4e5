"""Mode-B value+gradient kernel (K3): per-evaluation time on the C2 cloud, beside the value-only kernel and the CPU oracle."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import direct_visual_lidar_calibration_b200 as V
from direct_visual_lidar_calibration_b200 import synthetic as S
from oracle import oracle as O
from scipy.spatial.transform import Rotation

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
bag = S.make_bag("pinhole_1920x1080", "os1_64", n, config_index=1, bag_index=0)
T0 = S.perturb(S.gt_T_camera_lidar(), (0.5, 0.5, 0.5), (0.02, 0.02, 0.02))
cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
idx = V.ViewCulling(cam, (bag["width"], bag["height"])).cull_indices(bag["points"], T0)
pts, ins = bag["points"][idx], bag["intensities"][idx]
cost = V.NIDCost(cam, V.VisualLiDARData(bag["image"], pts, ins), 16)
tp = np.concatenate([Rotation.from_matrix(T0[:3, :3]).as_quat(), T0[:3, 3]])[None]
for _ in range(3):
    cost.evaluate_with_gradient(tp)
    cost.evaluate(tp)
reps = 30
t0 = time.perf_counter()
for _ in range(reps):
    ok, nid, grad = cost.evaluate_with_gradient(tp)
t_g = (time.perf_counter() - t0) / reps
t0 = time.perf_counter()
for _ in range(reps):
    cost.evaluate(tp)
t_v = (time.perf_counter() - t0) / reps
ocam = O.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
m = min(len(pts), 100000)
t0 = time.perf_counter()
rok, rnid, rgrad = O.nid_cost_bspline_grad(ocam, bag["image"], pts[:m], ins[:m], 16, tp[0])
t_cpu = (time.perf_counter() - t0) * len(pts) / m
sub = V.NIDCost(cam, V.VisualLiDARData(bag["image"], pts[:m], ins[:m]), 16)
_, snid, sgrad = sub.evaluate_with_gradient(tp)
print(json.dumps({"points": int(len(pts)), "grad_call_ms": 1e3 * t_g, "value_call_ms": 1e3 * t_v, "cpu_oracle_grad_ms_extrapolated": 1e3 * t_cpu, "speedup_vs_cpu_1core": t_cpu / t_g,
                  "check_points": m, "d_nid": abs(float(snid[0]) - rnid), "d_grad_rel": float(np.abs(sgrad[0] - rgrad).max() / max(1.0, np.abs(rgrad).max())), "grad": [float(g) for g in grad[0]]}))
