"""On-disk formats row: PLY / PNG / calib.json contract of the reference's preprocess -> calibrate pipeline."""
import json
import os

import numpy as np
import pytest

from direct_visual_lidar_calibration_b200 import io as vio
from direct_visual_lidar_calibration_b200 import synthetic as S


def test_ply_roundtrip_binary_and_ascii(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(1000, 3)).astype(np.float32)
    inten = (rng.integers(0, 256, 1000) / 256.0).astype(np.float32)
    p = str(tmp_path / "bag.ply")
    vio.save_ply_binary(p, pts, inten)
    head = open(p, "rb").read(200).decode("ascii", errors="replace")
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 1000\nproperty float x\nproperty float y\nproperty float z\nproperty float intensity\nend_header\n")
    q, i = vio.load_ply(p)
    assert q.dtype == np.float64 and np.array_equal(q, pts.astype(np.float64)) and np.array_equal(i, inten.astype(np.float64))
    # a foreign layout: doubles, extra properties, other order, ascii
    with open(tmp_path / "a.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment hi\nelement vertex 2\nproperty double intensity\nproperty float z\nproperty float x\nproperty uchar tag\nproperty float y\nend_header\n")
        f.write("0.5 3 1 7 2\n0.25 6 4 9 5\n")
    q, i = vio.load_ply(str(tmp_path / "a.ply"))
    assert np.array_equal(q, [[1, 2, 3], [4, 5, 6]]) and np.array_equal(i, [0.5, 0.25])


def test_tum_pose_convention_roundtrip():
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(1)
    for _ in range(20):
        R = Rotation.random(random_state=rng.integers(1 << 30)).as_matrix()
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, rng.normal(size=3)
        v = vio.T_to_tum(T)
        assert np.allclose(vio.tum_to_T(v), T, atol=1e-12)
        assert np.allclose(np.abs(v[3:]), np.abs(Rotation.from_matrix(R).as_quat()), atol=1e-12)  # x y z w order
        assert np.allclose(vio.invert_isometry(T) @ T, np.eye(4), atol=1e-12)
    # docs/programs.md:143-152 example: [x y z qx qy qz qw] with an un-normalised quaternion is normalised on load
    T = vio.tum_to_T([0.029965, 0.001851, 0.108368, -0.502097 * 2, 0.492510 * 2, -0.500947 * 2, 0.504366 * 2])
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)


def _write_dataset(tmp_path, n_points=30000):
    import cv2

    bag = S.make_bag("pinhole_640x480", "frustum", n_points, config_index=11, scale=0.5)
    d = str(tmp_path)
    cv2.imwrite(os.path.join(d, "bag0.png"), bag["image"])
    vio.save_ply_binary(os.path.join(d, "bag0.ply"), bag["points"][:, :3], bag["intensities"])
    T_init = S.perturb(bag["T_gt"], (0.3, -0.3, 0.3), (0.01, -0.01, 0.01))
    config = {
        "meta": {"data_path": d, "bag_names": ["bag0"]},
        "camera": {"camera_model": bag["camera_model"], "intrinsics": bag["intrinsics"], "distortion_coeffs": bag["distortion"]},
        "results": {"init_T_lidar_camera_auto": vio.T_to_tum(vio.invert_isometry(T_init))},
    }
    vio.save_calib_json(d, config)
    return bag, T_init


def test_dataset_files_load_like_the_reference(tmp_path):
    bag, _ = _write_dataset(tmp_path, 5000)
    data = vio.load_visual_lidar_data(str(tmp_path), "bag0")
    assert np.array_equal(data.image, bag["image"])
    assert np.array_equal(data.points[:, :3], bag["points"][:, :3]) and np.all(data.points[:, 3] == 1.0)
    assert np.array_equal(data.intensities, bag["intensities"])
    cfg = vio.load_calib_json(str(tmp_path))
    assert cfg["camera"]["camera_model"] == "plumb_bob" and cfg["meta"]["bag_names"] == ["bag0"]
    assert open(os.path.join(tmp_path, "calib.json")).read().endswith("}\n")


def test_cli_rejects_unknown_registration_type(tmp_path, capsys):
    from direct_visual_lidar_calibration_b200 import calibrate as cli

    _write_dataset(tmp_path, 1000)
    assert cli.main([str(tmp_path), "--registration_type", "nid_newton"]) == 1  # calibrate.cpp:105-108
    assert "unknown registration type" in capsys.readouterr().err


@pytest.mark.gpu
def test_cli_default_is_the_bfgs_branch(gpu, oracle, tmp_path):
    """The reference's default registration_type is nid_bfgs (calibrate.cpp:176): runs on K3 + the Ceres-free BFGS."""
    from direct_visual_lidar_calibration_b200 import calibrate as cli

    bag, T_init = _write_dataset(tmp_path)
    assert cli.main([str(tmp_path)]) == 0
    cfg = json.load(open(os.path.join(tmp_path, "calib.json")))
    T = vio.invert_isometry(vio.tum_to_T(cfg["results"]["T_lidar_camera"]))
    init_T = vio.invert_isometry(vio.tum_to_T(cfg["results"]["init_T_lidar_camera_auto"]))
    ocam = oracle.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
    from scipy.spatial.transform import Rotation

    def tp(M):
        return np.concatenate([Rotation.from_matrix(M[:3, :3]).as_quat(), M[:3, 3]])

    c0 = oracle.nid_cost_bspline(ocam, bag["image"], bag["points"], bag["intensities"], 16, tp(init_T))[1]
    c1 = oracle.nid_cost_bspline(ocam, bag["image"], bag["points"], bag["intensities"], 16, tp(T))[1]
    assert c1 <= c0 + 1e-6  # mode-B NID over the whole cloud did not get worse


@pytest.mark.gpu
def test_cli_end_to_end_writes_T_lidar_camera(gpu, oracle, tmp_path):
    from direct_visual_lidar_calibration_b200 import calibrate as cli

    bag, T_init = _write_dataset(tmp_path)
    rc = cli.main([str(tmp_path), "--registration_type", "nid_nelder_mead"])
    assert rc == 0
    cfg = json.load(open(os.path.join(tmp_path, "calib.json")))
    assert "init_T_lidar_camera_auto" in cfg["results"] and len(cfg["results"]["T_lidar_camera"]) == 7
    T = vio.invert_isometry(vio.tum_to_T(cfg["results"]["T_lidar_camera"]))
    # same answer as the oracle's calibrate on the same files' contents (through the TUM round trip)
    ocam = oracle.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
    init_T = vio.invert_isometry(vio.tum_to_T(cfg["results"]["init_T_lidar_camera_auto"]))
    ref = oracle.calibrate(ocam, [(bag["image"], bag["points"], bag["intensities"])], init_T)
    assert np.abs(T - ref["T"]).max() < 1e-9
    fov = oracle.estimate_camera_fov(ocam, bag["width"], bag["height"])
    nid0 = oracle.nid_calculate(ocam, bag["image"], bag["points"], bag["intensities"], 16, fov, init_T)[0]
    nid1 = oracle.nid_calculate(ocam, bag["image"], bag["points"], bag["intensities"], 16, fov, T)[0]
    assert nid1 < nid0
