"""Per-launch time of the NID histogram kernel vs cloud size N and poses-per-launch P, for each kernel variant.
A linear fit t = t0 + c * N * P separates the fixed cost (launch + serial finalize tail) from the per-point-pose cost.
Prints one JSON line per measurement; used for profiles/."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import direct_visual_lidar_calibration_b200 as V
from direct_visual_lidar_calibration_b200 import synthetic as S

bag = S.make_bag("pinhole_1920x1080", "os1_64", 1_000_000, config_index=1)
cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
idx = V.ViewCulling(cam, (bag["width"], bag["height"])).cull_indices(bag["points"], bag["T_gt"])
pts, inten = bag["points"][idx], bag["intensities"][idx]
rng = np.random.default_rng(0)
poses = np.stack([S.perturb(bag["T_gt"], rng.uniform(-0.05, 0.05, 3), rng.uniform(-0.002, 0.002, 3)) for _ in range(8)])
for n in (1024, 65536, 262144, len(pts)):
    cost = V.CostCalculatorNID(cam, V.VisualLiDARData(bag["image"], pts[:n], inten[:n]))
    for variant, name in ((3, "filter_kpt4"), (2, "filter_kpt2"), (1, "exact_fp64")):
        cost.set_kernel_variant(variant)
        for P in (1, 4, 8):
            for _ in range(5):
                cost.calculate_batch(poses[:P])
            cost.set_profiling(True)
            cost.reset_profile()
            for _ in range(50):
                cost.calculate_batch(poses[:P])
            pr = cost.profile()
            cost.set_profiling(False)
            us = 1e3 * pr["kernel_ms_total"] / pr["kernel_launches"]
            print(json.dumps({"kernel": name, "n_points": n, "poses": P, "us_per_launch": round(us, 3), "gpointposes_per_s": round(n * P / us * 1e-3, 2)}), flush=True)

# where the time of one launch goes (globaltimer stamps inside the kernel + host clock)
for n in (1024, len(pts)):
    cost = V.CostCalculatorNID(cam, V.VisualLiDARData(bag["image"], pts[:n], inten[:n]))
    for P in (1, 4, 8):
        for _ in range(5):
            cost.calculate_batch(poses[:P])
        tl = [cost.debug_timeline(poses[:P]) for _ in range(20)]
        med = {k: round(float(np.median([t[k] for t in tl])), 2) for k in tl[0]}
        print(json.dumps({"timeline_us": med, "n_points": n, "poses": P}), flush=True)
