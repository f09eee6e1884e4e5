"""BASELINE.json configs C1 / C3 / C5 on one GPU: evaluations/s and Mpoints/s of the batched NID kernel, HBM-roofline
fraction (algorithmic bytes 16 N + W H per launch / measured HBM peak), the CPU oracle timed on the same inputs, and a
full-size parity check (integer histogram identical to the oracle's for one pose).  One JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import direct_visual_lidar_calibration_b200 as V
from direct_visual_lidar_calibration_b200 import synthetic as S
from oracle import oracle as O

PEAK = 6576.1
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass


def measure(name, bag, poses, launches=30, cpu_evals=3, variants=((0, "filter"), (2, "filter_kpt2"), (3, "filter_kpt4"), (1, "exact_fp64")), tile_order=True):
    cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
    ocam = O.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
    data = V.VisualLiDARData(bag["image"], bag["points"], bag["intensities"])
    t0 = time.perf_counter()
    cost = V.CostCalculatorNID(cam, data)
    t_create = time.perf_counter() - t0
    if tile_order:
        cost.reorder_for_pose(poses[0])  # what the calibrate path gets from its culling pass
    n, (H, W) = data.size(), bag["image"].shape
    fov = cost.max_fov
    # parity at full size, one pose
    nid_g, hist_g = cost.calculate_batch(poses[:1], return_hist=True)
    t0 = time.perf_counter()
    nid_o, hist_o = O.nid_calculate(ocam, bag["image"], bag["points"], bag["intensities"], 16, fov, poses[0])
    t_cpu1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    for k in range(cpu_evals):
        O.nid_calculate(ocam, bag["image"], bag["points"], bag["intensities"], 16, fov, poses[k % len(poses)])
    t_cpu = (time.perf_counter() - t0) / cpu_evals
    t0 = time.perf_counter()
    O.nid_calculate(ocam, bag["image"], bag["points"], bag["intensities"], 16, fov, poses[0], omp=True)
    t_cpu_omp = time.perf_counter() - t0
    parity = {"hist_identical": bool(np.array_equal(hist_g[0], hist_o)), "differing_counts": int(np.abs(hist_g[0] - hist_o).sum()), "abs_dnid": float(abs(nid_g[0] - nid_o)), "inliers": int(hist_o.sum())}
    alg_bytes = 16 * n + W * H
    for variant, vname in variants:
        cost.set_kernel_variant(variant)
        if variant == 0 and not cost.filter_enabled:
            continue
        for P in (1, 4, 8):
            if P > len(poses):
                continue
            batch = poses[:P]
            for _ in range(3):
                cost.calculate_batch(batch)
            cost.set_profiling(True)
            cost.reset_profile()
            t0 = time.perf_counter()
            for _ in range(launches):
                cost.calculate_batch(batch)
            wall = (time.perf_counter() - t0) / launches
            pr = cost.profile()
            cost.set_profiling(False)
            k_us = 1e3 * pr["kernel_ms_total"] / pr["kernel_launches"]
            gbs = alg_bytes / (k_us * 1e-6) * 1e-9
            print(json.dumps({
                "config": name, "kernel": vname, "camera": bag["camera_model"], "points": n, "image": f"{W}x{H}", "poses_per_launch": P, "tile_ordered": tile_order,
                "kernel_us": round(k_us, 2), "call_us": round(wall * 1e6, 2), "evals_per_s_kernel": round(P / (k_us * 1e-6)), "evals_per_s_call": round(P / wall),
                "mpoints_per_s": round(n * P / (k_us * 1e-6) * 1e-6), "achieved_GBps": round(gbs, 1), "hbm_frac_of_measured": round(gbs / PEAK, 4),
                "cpu_oracle_evals_per_s_1core": round(1.0 / t_cpu, 2), "cpu_oracle_omp_evals_per_s": round(1.0 / t_cpu_omp, 2), "cpu_cores": os.cpu_count(),
                "create_ms": round(1e3 * t_create, 2), "parity": parity,
            }), flush=True)


def main():
    which = sys.argv[1:] or ["C1", "C3", "C3f", "TILE", "C5"]
    rng = np.random.default_rng(5)
    if "C1" in which:
        bag = S.config_c1()
        poses = np.stack([S.perturb(bag["T_gt"], rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.01, 0.01, 3)) for _ in range(8)])
        measure("C1", bag, poses, launches=100, cpu_evals=10)
    if "C3" in which:
        bag = S.config_c3(5_000_000, "equirect_3840x1920")
        poses = np.stack([S.perturb(bag["T_gt"], rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.01, 0.01, 3)) for _ in range(8)])
        measure("C3-equirectangular", bag, poses)
    if "C3f" in which:
        bag = S.config_c3(5_000_000, "fisheye_1920x1080")
        poses = np.stack([S.perturb(bag["T_gt"], rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.01, 0.01, 3)) for _ in range(8)])
        measure("C3-fisheye", bag, poses)
    if "TILE" in which:  # A/B of the per-model tile-size default on the models C3 does not cover
        for key in ("atan_1920x1080", "omnidir_1920x1080", "rational_1920x1080"):
            bag = S.config_c3(2_000_000, "pinhole_1920x1080")  # the scene renderer has no inverse for these models: timing + parity only
            bag["camera_model"], bag["intrinsics"], bag["distortion"] = S.CAMERAS[key][0], S.CAMERAS[key][1], S.CAMERAS[key][2]
            poses = np.stack([S.perturb(bag["T_gt"], rng.uniform(-0.3, 0.3, 3), rng.uniform(-0.01, 0.01, 3)) for _ in range(8)])
            measure(f"TILE-{key}", bag, poses, launches=20, cpu_evals=1, variants=((2, "filter_kpt2"), (3, "filter_kpt4")))
    if "C5" in which:
        bag = S.config_c3(5_000_000, "pinhole_1920x1080")
        grid = S.pose_grid(bag["T_gt"])  # 16384 poses
        sub = grid[:: len(grid) // 256][:256]
        cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
        cost = V.CostCalculatorNID(cam, V.VisualLiDARData(bag["image"], bag["points"], bag["intensities"]))
        cost.calculate_batch(sub[:16])
        t0 = time.perf_counter()
        nid = cost.calculate_batch(sub)
        dt = time.perf_counter() - t0
        ocam = O.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
        t0 = time.perf_counter()
        ref = [O.nid_calculate(ocam, bag["image"], bag["points"], bag["intensities"], 16, cost.max_fov, T)[0] for T in sub[:4]]
        t_cpu = (time.perf_counter() - t0) / 4
        best = int(np.nanargmin(nid))
        print(json.dumps({
            "config": "C5-pose-grid", "points": 5_000_000, "poses_scored": len(sub), "grid_poses": int(len(grid)), "seconds": round(dt, 4), "evals_per_s": round(len(sub) / dt),
            "mpoints_per_s": round(5.0 * len(sub) / dt), "extrapolated_seconds_16384_poses_1gpu": round(dt * len(grid) / len(sub), 2),
            "cpu_oracle_evals_per_s_1core": round(1.0 / t_cpu, 3), "extrapolated_cpu_seconds_16384_poses": round(t_cpu * len(grid)),
            "max_abs_dnid_vs_oracle_4poses": float(np.max(np.abs(np.array(ref) - nid[:4]))), "argmin_pose": best, "min_nid": float(nid[best]),
        }), flush=True)
        measure("C5-cloud-pinhole", bag, sub[:8], launches=20, cpu_evals=2)


if __name__ == "__main__":
    main()
