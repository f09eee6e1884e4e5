"""`calibrate` on the preprocessed-dataset contract of the reference (src/calibrate.cpp:165-202).

    python -m direct_visual_lidar_calibration_b200.calibrate <data_path> [--registration_type nid_bfgs|nid_nelder_mead]

Reads <data_path>/calib.json (+ <bag>.png / <bag>.ply), takes the initial guess from results.init_T_lidar_camera (manual)
or results.init_T_lidar_camera_auto, runs VisualCameraCalibration on the GPU and writes results.T_lidar_camera back
(calibrate.cpp:57-76,128-140).  No viewer; flags keep the reference's names and defaults.  nid_nelder_mead follows the
reference evaluation for evaluation; nid_bfgs (the reference's default) runs the mode-B value + gradient kernel under a
Ceres-free BFGS (bfgs.py) -- same problem, not Ceres' iterates."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import RegistrationType, VisualCameraCalibration, VisualCameraCalibrationParams, create_camera
from . import io as vio


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="calibrate")
    ap.add_argument("data_path")
    ap.add_argument("--first_n_bags", type=int, default=None, help="use only the first N bags (just for evaluation)")
    ap.add_argument("--disable_culling", action="store_true", help="disable depth buffer-based hidden points removal")
    ap.add_argument("--nid_bins", type=int, default=16, help="Number of histogram bins for NID")
    ap.add_argument("--registration_type", default="nid_bfgs", help="nid_bfgs or nid_nelder_mead")
    ap.add_argument("--nelder_mead_init_step", type=float, default=1e-3)
    ap.add_argument("--nelder_mead_convergence_criteria", type=float, default=1e-8)
    ap.add_argument("--device", type=int, default=-1)
    args = ap.parse_args(argv)

    if args.registration_type not in ("nid_nelder_mead", "nid_bfgs"):
        print(f"error: unknown registration type {args.registration_type}", file=sys.stderr)  # calibrate.cpp:105-108
        return 1
    config = vio.load_calib_json(args.data_path)
    cam = config["camera"]
    proj = create_camera(cam["camera_model"], cam["intrinsics"], cam["distortion_coeffs"])
    if proj is None:
        return 1
    bag_names = list(config["meta"]["bag_names"])
    if args.first_n_bags is not None:
        bag_names = bag_names[: args.first_n_bags]
        print(f"use only the first {args.first_n_bags} bags")
    dataset = []
    for name in bag_names:
        print(f"loading {args.data_path}/{name}.(png|ply)")
        dataset.append(vio.load_visual_lidar_data(args.data_path, name))

    results = config.get("results", {})
    if "init_T_lidar_camera" in results:  # calibrate.cpp:57-60
        print("use manually estimated initial guess")
        init_values = results["init_T_lidar_camera"]
    elif "init_T_lidar_camera_auto" in results:  # :61-65
        print("use automatically estimated initial guess")
        init_values = results["init_T_lidar_camera_auto"]
    else:
        print("error: initial guess of T_lidar_camera must be computed before calibration!!", file=sys.stderr)  # :67-70
        return 1
    init_T_camera_lidar = vio.invert_isometry(vio.tum_to_T(init_values))  # :72-76

    params = VisualCameraCalibrationParams()  # :95-110
    params.disable_z_buffer_culling = args.disable_culling
    params.nid_bins = args.nid_bins
    params.nelder_mead_init_step = args.nelder_mead_init_step
    params.nelder_mead_convergence_criteria = args.nelder_mead_convergence_criteria
    params.registration_type = RegistrationType.NID_NELDER_MEAD if args.registration_type == "nid_nelder_mead" else RegistrationType.NID_BFGS
    calib = VisualCameraCalibration(proj, dataset, params, device=args.device)
    T_camera_lidar = calib.calibrate(init_T_camera_lidar)
    for _, cost in calib.trace:
        print(f"cost:{cost}")  # visual_camera_calibration.cpp:115

    T_lidar_camera = vio.invert_isometry(T_camera_lidar)  # :128-133
    config.setdefault("results", {})["T_lidar_camera"] = vio.T_to_tum(T_lidar_camera)
    vio.save_calib_json(args.data_path, config)
    print("--- T_lidar_camera ---")
    print(np.array2string(T_lidar_camera, precision=6, suppress_small=True))
    print(f"saved to {args.data_path}/calib.json")
    return 0


if __name__ == "__main__":
    sys.exit(main())
