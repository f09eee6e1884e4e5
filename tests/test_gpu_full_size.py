"""GPU suite (-m gpu), BASELINE.json's full sizes against the CPU oracle: the CUDA path, through the C ABI, must produce the
oracle's joint histograms bit for bit on the 1 M-point C2 cloud (plumb_bob 1920x1080), the 5 M-point C3 cloud
(equirectangular 3840x1920) and the C5 cloud (5 M points under the 1920x1080 pinhole camera) -- one or two poses each (the
oracle needs 0.1-0.6 s per evaluation at these sizes) -- and the pose-list kernel must agree with the per-pose calls."""
import numpy as np
import pytest

import util
from direct_visual_lidar_calibration_b200 import synthetic as S

pytestmark = pytest.mark.gpu
NID_TOL = 1e-12


def _check(V, O, model, intr, dist, image, pts, inten, Ts):
    cam = V.create_camera(model, intr, dist)
    cost = V.CostCalculatorNID(cam, V.VisualLiDARData(image, pts, inten))
    nid, hist = cost.calculate_batch(Ts, return_hist=True)
    ocam = O.create_camera(model, intr, dist)
    fov = O.estimate_camera_fov(ocam, image.shape[1], image.shape[0])
    assert cost.max_fov == fov
    for p, T in enumerate(Ts):
        ref_nid, ref_hist = O.nid_calculate(ocam, image, pts, inten, 16, fov, T)
        assert int(np.abs(hist[p] - ref_hist).sum()) == 0, f"{model}: {int(np.abs(hist[p] - ref_hist).sum())} counts differ at pose {p}"
        assert hist[p].sum() > 0.05 * pts.shape[0]
        assert abs(nid[p] - ref_nid) < NID_TOL
    # the same poses as a list through vlcal_nid_score_poses, and with the tile-ordered cloud: identical bits
    assert np.array_equal(V.score_poses([cost], Ts), nid)
    cost.reorder_for_pose(Ts[0])
    nid2, hist2 = cost.calculate_batch(Ts, return_hist=True)
    assert np.array_equal(hist2, hist) and np.array_equal(nid2, nid)
    return cost


def test_c2_full_size_histograms_equal_the_oracle(gpu, oracle):
    bag = S.config_c2(1_000_000)
    Ts = [bag["T_init"], S.perturb(bag["T_gt"], (-0.4, 0.3, 0.2), (0.01, -0.02, 0.015))]
    _check(gpu, oracle, bag["camera_model"], bag["intrinsics"], bag["distortion"], bag["image"], bag["points"], bag["intensities"], Ts)


def test_c3_and_c5_full_size_histograms_equal_the_oracle(gpu, oracle):
    seed = S.SEED0 + 2000
    pts, inten = S.make_cloud("avia", 5_000_000, scene_seed=seed, seed=seed + 1)
    T_gt = S.gt_T_camera_lidar()
    for key in ("equirect_3840x1920", "pinhole_1920x1080"):  # C3, then the same cloud under the C5 camera
        model, intr, dist, w, h = S.CAMERAS[key]
        image = S.render_image(model, intr, dist, w, h, T_gt, scene_seed=seed, noise_seed=seed + 2)
        Ts = [S.perturb(T_gt, (0.5, 0.5, 0.5), (0.02, 0.02, 0.02))]
        cost = _check(gpu, oracle, model, intr, dist, image, pts, inten, Ts)
        if key == "pinhole_1920x1080":
            # a slice of the C5 pose grid in one launch: every score equals the per-pose evaluation bit for bit
            grid = S.pose_grid(T_gt)[:: 16384 // 48][:48]
            listed = gpu.score_poses([cost], grid)
            cost.set_kernel_variant(4)  # round-1 kernels, one launch per 8 poses
            assert np.array_equal(cost.calculate_batch(grid), listed, equal_nan=True)
