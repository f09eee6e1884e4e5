"""CPU suite, part 1: pins for the oracle (oracle/vlcal_oracle.c).

The reference has no tests / golden vectors: these pins are ours -- hand-derived known answers,
OpenCV cross-checks for the models the reference declares OpenCV-compatible (pinhole.hpp:9, fisheye.hpp:9-10,
rational_polynomial.hpp:7), closed forms, and properties of the algorithm.
"""
import json
import math
import os

import numpy as np
import pytest

import util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_nid_known_answers(oracle):
    with open(os.path.join(GOLDEN, "nid_kat.json")) as f:
        kats = json.load(f)
    assert len(kats) == 5
    for k in kats:
        nid, (Hr, Hs, Hrs, MI) = oracle.nid_from_hist(np.array(k["hist"], dtype=np.int32))
        assert abs(nid - k["NID"]) < 1e-13, k["name"]
        assert abs(Hr - k["Hr"]) < 1e-13 and abs(Hs - k["Hs"]) < 1e-13 and abs(Hrs - k["Hrs"]) < 1e-13 and abs(MI - k["MI"]) < 1e-13
    # SURVEY.md 8c literal values
    assert abs(oracle.nid_from_hist(np.array([[3, 1], [0, 4]]))[0] - 0.609576026642764) < 1e-14
    assert abs(oracle.nid_from_hist(np.array([[1, 1], [1, 1]]))[0] - 1.00000000000289) < 1e-13  # > 1 !
    assert abs(oracle.nid_from_hist(np.ones((16, 16), dtype=np.int32))[0] - 0.999959608536476) < 1e-13
    assert math.isnan(oracle.nid_from_hist(np.zeros((16, 16), dtype=np.int32))[0])  # 0 inliers -> NaN (no guard in the reference)


def test_image_bin_lut_has_no_rounding_ambiguity():
    # int(v/255.0*16) is an exact integer only at v in {0, 255} (255 -> 16 -> clamped to 15)
    for v in range(256):
        x = v / 255.0 * 16
        if x == int(x):
            assert v in (0, 255)


def test_create_camera_rules(oracle):
    assert oracle.create_camera("nope", [1, 2, 3, 4], []) is None  # create_camera.cpp:49-50
    assert oracle.create_camera("plumb_bob", [1, 2, 3], []) is None  # :19-22
    cam = oracle.create_camera("plumb_bob", [1, 2, 3, 4], [0.1, 0.2])  # :24-27 zero padding
    assert list(cam.dist[:5]) == [0.1, 0.2, 0.0, 0.0, 0.0]
    cam = oracle.create_camera("fisheye", [1, 2, 3, 4], [1, 2, 3, 4, 5, 6])  # truncation
    assert cam.n_dist == 4 and list(cam.dist[:4]) == [1, 2, 3, 4]
    assert oracle.create_camera("equidistant", [1, 2, 3, 4], []).model == oracle.create_camera("fisheye", [1, 2, 3, 4], []).model
    assert oracle.create_camera("equirectangular", [640, 320], []).n_dist == 0
    assert oracle.create_camera("omnidir", [1, 2, 3, 4, 5], []).n_intr == 5


def test_cameras_match_opencv(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 1, (500, 3))
    pts[:, 2] = rng.uniform(0.5, 5, 500)
    intr = [400.0, 410.0, 320.0, 240.0]
    K = np.array([[400.0, 0, 320.0], [0, 410.0, 240.0], [0, 0, 1.0]])
    z3 = np.zeros(3)
    d5 = [-0.04, 0.08, 1e-4, -3e-4, -0.04]
    ref, _ = cv2.projectPoints(pts, z3, z3, K, np.array(d5))
    assert np.abs(oracle.project(oracle.create_camera("plumb_bob", intr, d5), pts) - ref[:, 0]).max() < 1e-9
    d8 = d5 + [0.01, 0.02, -0.005]
    ref, _ = cv2.projectPoints(pts, z3, z3, K, np.array(d8))
    assert np.abs(oracle.project(oracle.create_camera("rational_polynomial", intr, d8), pts) - ref[:, 0]).max() < 1e-9
    d4 = [0.01, -0.02, 0.003, -0.001]
    ref, _ = cv2.fisheye.projectPoints(pts.reshape(-1, 1, 3), z3, z3, K, np.array(d4))
    assert np.abs(oracle.project(oracle.create_camera("fisheye", intr, d4), pts) - ref[:, 0]).max() < 1e-9


def test_camera_closed_forms(oracle):
    # equirectangular: forward axis -> image centre; +x -> 3/4 width; up (-y) -> above centre
    cam = oracle.create_camera("equirectangular", [640.0, 320.0], [])
    assert np.allclose(oracle.project(cam, [0, 0, 2.0]), [320.0, 160.0])
    assert np.allclose(oracle.project(cam, [3.0, 0, 0]), [480.0, 160.0])
    assert oracle.project(cam, [0, -1.0, 1.0])[1] < 160.0
    assert np.allclose(oracle.project(cam, [0.01, 0.01, 0.01]), [320.0, 160.0])  # |p|^2 < 1e-3 -> centre (equirectangular.hpp:15)
    # omnidir with xi = 0 and no distortion degenerates to the pinhole
    p = np.array([0.3, -0.2, 1.5])
    o = oracle.project(oracle.create_camera("omnidir", [400.0, 410.0, 320.0, 240.0, 0.0], []), p)
    q = oracle.project(oracle.create_camera("plumb_bob", [400.0, 410.0, 320.0, 240.0], []), p)
    assert np.allclose(o, q, atol=1e-10)
    # atan: tiny d0 or tiny r -> identity distortion (atan.hpp:17)
    a = oracle.project(oracle.create_camera("atan", [400.0, 410.0, 320.0, 240.0], [1e-8]), p)
    assert np.allclose(a, q, atol=1e-12)
    a = oracle.project(oracle.create_camera("atan", [400.0, 400.0, 320.0, 240.0], [0.9]), p)
    r = math.hypot(0.2, -0.2 / 1.5 * 1.5 / 1.5)  # noqa: F841  (documented formula below)
    x, y = p[0] / p[2], p[1] / p[2]
    rr = math.hypot(x, y)
    fac = (1 / 0.9) * math.atan(rr * 2 * math.tan(0.45)) / rr
    assert np.allclose(a, [400 * fac * x + 320, 400 * fac * y + 240], atol=1e-10)
    # fisheye uses abs(z): a point behind the camera mirrors to the front (fisheye.hpp:16), r = 0 -> NaN
    f = oracle.create_camera("fisheye", [300.0, 300.0, 320.0, 240.0], [])
    assert np.allclose(oracle.project(f, [0.2, 0.1, -1.0]), oracle.project(f, [0.2, 0.1, 1.0]))
    assert np.isnan(oracle.project(f, [0.0, 0.0, 1.0])).all()


def test_se3_expmap_matches_matrix_exponential(oracle):
    scipy_linalg = pytest.importorskip("scipy.linalg")
    rng = np.random.default_rng(1)
    for scale in (1e-9, 1e-3, 0.3, 2.0):
        for _ in range(5):
            xi = rng.normal(size=6) * scale
            w, v = xi[:3], xi[3:]
            A = np.zeros((4, 4))
            A[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
            A[:3, 3] = v
            assert np.abs(scipy_linalg.expm(A) - oracle.se3_expmap(xi)).max() < 1e-12
    # GTSAM tangent order is (omega, v): a pure translation lives in the last three entries
    T = oracle.se3_expmap([0, 0, 0, 1, 2, 3])
    assert np.allclose(T[:3, 3], [1, 2, 3]) and np.allclose(T[:3, :3], np.eye(3))


def test_rotation_angle_and_isometry(oracle):
    T = oracle.se3_expmap([0.1, -0.2, 0.3, 0.5, 0.1, -0.3])
    assert abs(oracle.rotation_angle(T) - math.sqrt(0.01 + 0.04 + 0.09)) < 1e-12
    assert np.allclose(oracle.isometry_mul(T, np.linalg.inv(T)), np.eye(4), atol=1e-14)


def test_estimate_camera_fov(oracle):
    # undistorted pinhole: the widest of the three probed pixels is the corner (0,0): acos(z) of the ray through it
    cam = oracle.create_camera("plumb_bob", [400.0, 400.0, 320.0, 240.0], [])
    fov = oracle.estimate_camera_fov(cam, 640, 480)
    assert abs(fov - math.atan(math.hypot(320.0, 240.0) / 400.0)) < 2e-3  # NelderMead<2> tolerance 1e-5 on the simplex variance
    # equirectangular: pixel (0,0) is lon=-pi, lat=+pi/2 -> the model reaches beyond 90 degrees
    cam = oracle.create_camera("equirectangular", [640.0, 320.0], [])
    assert oracle.estimate_camera_fov(cam, 640, 320) > 1.5


def test_nelder_mead_reference_quirks(oracle):
    # evaluation bookkeeping: N+1 initial evaluations, then xo (unused) and xr every iteration (nelder_mead.hpp:36,44,58,61)
    r = oracle.nelder_mead(lambda x: float((x[0] - 1) ** 2 + 3 * (x[1] + 2) ** 2), [0.0, 0.0])
    assert r["converged"] and r["num_evaluations"] == len(r["calls"])
    assert abs(r["x"][0] - 1) < 5e-2 and abs(r["x"][1] + 2) < 5e-2
    calls = r["calls"]
    assert np.allclose(calls[0][0], [0, 0]) and np.allclose(calls[1][0], [0.1, 0]) and np.allclose(calls[2][0], [0, 0.1])
    # first iteration: simplex sorted by value, xo = mean of the best N, xr = xo + (xo - worst)
    f = lambda x: float((x[0] - 1) ** 2 + 3 * (x[1] + 2) ** 2)  # noqa: E731
    pts = sorted([np.array([0, 0.0]), np.array([0.1, 0]), np.array([0, 0.1])], key=f)
    xo = (pts[0] + pts[1]) / 2
    assert np.allclose(calls[3][0], xo) and np.allclose(calls[4][0], xo + (xo - pts[2]))
    # max_iterations = 0 -> result is the un-sorted x[0] = x0 (the loop body never runs)
    r0 = oracle.nelder_mead(f, [0.3, 0.4], max_iterations=0)
    assert np.allclose(r0["x"], [0.3, 0.4]) and not r0["converged"] and r0["num_iterations"] == 0
    # a NaN objective never satisfies a comparison -> contraction branch, still terminates
    rn = oracle.nelder_mead(lambda x: float("nan"), [0.0, 0.0], max_iterations=20)
    assert rn["num_iterations"] == 19 or rn["converged"]


@pytest.mark.parametrize("model", util.MODELS)
def test_oracle_matches_golden(oracle, model):
    g = np.load(os.path.join(GOLDEN, f"mode_a_{model}.npz"))
    H, W = g["image"].shape
    cam = oracle.create_camera(model, g["intrinsics"], g["distortion"])
    fov = oracle.estimate_camera_fov(cam, W, H)
    assert fov == float(g["max_fov"])
    pts = g["points"].astype(np.float64)
    ins = g["intensities"].astype(np.float64)
    assert g["hist"].sum() > 300  # the fixture exercises real inliers
    for p, T in enumerate(g["poses"]):
        nid, h = oracle.nid_calculate(cam, g["image"], pts, ins, 16, fov, T)
        assert np.array_equal(h, g["hist"][p])
        assert nid == g["nid"][p] or (math.isnan(nid) and math.isnan(g["nid"][p]))
    idx = oracle.view_cull(cam, W, H, fov, True, pts, g["poses"][0])
    assert np.array_equal(idx, g["cull_indices"])


@pytest.mark.parametrize("model", util.MODELS)
def test_nid_properties(oracle, model):
    pr = util.random_problem(model, n=5000, seed=3)
    cam = oracle.create_camera(model, pr["intrinsics"], pr["distortion"])
    fov = oracle.estimate_camera_fov(cam, pr["W"], pr["H"])
    nid, h = oracle.nid_calculate(cam, pr["image"], pr["points"], pr["intensities"], 16, fov, pr["T"])
    assert 0 < h.sum() <= 5000 and 0.0 <= nid <= 1.0 + 1e-9
    # permutation invariance of the integer histogram
    perm = np.random.default_rng(5).permutation(5000)
    nid2, h2 = oracle.nid_calculate(cam, pr["image"], pr["points"][perm], pr["intensities"][perm], 16, fov, pr["T"])
    assert np.array_equal(h, h2) and nid == nid2
    # the OpenMP "best effort" variant computes the same histogram
    nid3, h3 = oracle.nid_calculate(cam, pr["image"], pr["points"], pr["intensities"], 16, fov, pr["T"], omp=True)
    assert np.array_equal(h, h3) and nid == nid3
    # other bin counts
    for bins in (4, 32):
        _, hb = oracle.nid_calculate(cam, pr["image"], pr["points"], pr["intensities"], bins, fov, pr["T"])
        assert hb.shape == (bins, bins) and hb.sum() == h.sum()
    # marginal consistency with an independent numpy count of the lidar bins of the inliers is covered by hist.sum()


def test_truncation_and_bounds_edges(oracle):
    # pinhole without distortion, identity pose: u = 100*x/z + 2, v = 100*y/z + 2 on a 4x4 image
    cam = oracle.create_camera("plumb_bob", [100.0, 100.0, 2.0, 2.0], [])
    img = np.arange(16, dtype=np.uint8).reshape(4, 4) * 16
    T = np.eye(4)

    def inliers(x, y):
        pts = np.array([[x, y, 1.0, 1.0]])
        return oracle.nid_calculate(cam, img, pts, np.array([0.5]), 16, 1.4, T)[1].sum()

    assert inliers(-0.025, 0.0) == 1  # u = -0.5 -> cast<int> truncates toward zero -> column 0, ACCEPTED (cost_calculator_nid.cpp:37-38)
    assert inliers(-0.0301, 0.0) == 0  # u = -1.01 -> -1 -> rejected
    assert inliers(0.0199, 0.0) == 1  # u = 3.99 -> 3 accepted
    assert inliers(0.02, 0.0) == 0  # u = 4.0 -> rejected (>= W)
    assert inliers(0.0, -0.025) == 1 and inliers(0.0, 0.0201) == 0
    # behind the camera: rejected by the FoV test although the pinhole would project it
    pts = np.array([[0.0, 0.0, -1.0, 1.0]])
    assert oracle.nid_calculate(cam, img, pts, np.array([0.5]), 16, 1.4, T)[1].sum() == 0
    # NaN coordinates are rejected (cvttsd2si -> INT_MIN)
    pts = np.array([[np.nan, 0.0, 1.0, 1.0]])
    assert oracle.nid_calculate(cam, img, pts, np.array([0.5]), 16, 1.4, T)[1].sum() == 0
    # intensity 1.0 -> bin 16 -> clamped to 15; negative -> 0
    h = oracle.nid_calculate(cam, img, np.array([[0, 0, 1.0, 1.0]] * 2), np.array([1.0, -0.3]), 16, 1.4, T)[1]
    assert h[:, 15].sum() == 1 and h[:, 0].sum() == 1


def test_view_culling_semantics(oracle):
    cam = oracle.create_camera("plumb_bob", [100.0, 100.0, 2.0, 2.0], [])
    T = np.eye(4)
    fov = 1.4
    # three points on the same pixel at ranges 2.0, 2.05, 2.2: the nearest and the one within +0.1 m survive
    pts = np.array([[0.001, 0.001, 2.2, 1.0], [0.001, 0.001, 2.0, 1.0], [0.001, 0.001, 2.05, 1.0], [5.0, 0.0, 1.0, 1.0]])
    assert list(oracle.view_cull(cam, 4, 4, fov, True, pts, T)) == [1, 2]
    assert list(oracle.view_cull(cam, 4, 4, fov, False, pts, T)) == [0, 1, 2]  # disable_culling: FoV + bounds only
    # the FoV test normalises the homogeneous 4-vector (view_culling.cpp:45): looser than the cost's test
    p = np.array([[0.0, 0.0, 0.5, 1.0]])  # z/|(x,y,z,1)| = 0.447 ; z/|xyz| = 1
    assert list(oracle.view_cull(cam, 4, 4, math.acos(0.5), True, p, T)) == []
    assert oracle.nid_calculate(cam, np.zeros((4, 4), np.uint8), p, np.array([0.1]), 16, math.acos(0.5), T)[1].sum() == 1


def test_nid_bspline_weights_sum_to_one(oracle):
    pr = util.random_problem("plumb_bob", n=2000, seed=9)
    cam = oracle.create_camera("plumb_bob", pr["intrinsics"], pr["distortion"])
    T = pr["T"]
    # Sophus parameterisation [qx qy qz qw tx ty tz]
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(T[:3, :3]).as_quat()
    ok, nid, hist = oracle.nid_cost_bspline(cam, pr["image"], pr["points"], pr["intensities"], 16, list(q) + list(T[:3, 3]))
    assert ok and 0 < nid <= 1.0 + 1e-9
    # every inlier contributes total weight 1 (sum of B-spline weights), so the soft histogram sums to the inlier count
    assert abs(hist.sum() - round(hist.sum())) < 1e-6 and hist.sum() > 100


def test_calibrate_recovers_ground_truth(oracle):
    from direct_visual_lidar_calibration_b200 import synthetic as S

    bag = S.make_bag("pinhole_640x480", "frustum", 30000, config_index=9, scale=0.5)
    cam = oracle.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
    T0 = S.perturb(bag["T_gt"], (0.3, -0.3, 0.3), (0.01, -0.01, 0.01))
    p = oracle.default_calib_params()
    p.max_inner_iterations = 60
    p.max_outer_iterations = 2
    fov = oracle.estimate_camera_fov(cam, bag["width"], bag["height"])
    c0 = oracle.nid_calculate(cam, bag["image"], bag["points"], bag["intensities"], 16, fov, T0)[0]
    r = oracle.calibrate(cam, [(bag["image"], bag["points"], bag["intensities"])], T0, p)
    c1 = oracle.nid_calculate(cam, bag["image"], bag["points"], bag["intensities"], 16, fov, r["T"])[0]
    assert c1 < c0  # the solve improves the NID
    assert r["outer_iterations"] >= 1 and r["total_evaluations"] == r["trace"].shape[0]


def _numpy_nid_plumb_bob(image, pts, inten, intr, dist, bins, max_fov, T):
    """Independent restatement of cost_calculator_nid.cpp:21-67 + pinhole.hpp in vectorised numpy (same operation order)."""
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    pc = [((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] for r in range(3)]
    n2 = (pc[0] * pc[0] + pc[1] * pc[1]) + pc[2] * pc[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        nz = np.where(n2 > 0, pc[2] / np.sqrt(n2), pc[2])
        keep = ~(nz < math.cos(max_fov))
        px, py = pc[0] / pc[2], pc[1] / pc[2]
        k1, k2, p1, p2, k3 = dist
        x2, y2 = px * px, py * py
        r2 = x2 + y2
        r4 = r2 * r2
        r6 = r2 * r4
        rc = 1.0 + k1 * r2 + k2 * r4 + k3 * r6
        t1, t2, t3 = 2.0 * px * py, r2 + 2.0 * x2, r2 + 2.0 * y2
        u = intr[0] * (rc * px + p1 * t1 + p2 * t2) + intr[2]
        v = intr[1] * (rc * py + p1 * t3 + p2 * t1) + intr[3]
        ok = np.isfinite(u) & np.isfinite(v) & (np.abs(u) < 2e9) & (np.abs(v) < 2e9)
        ix = np.where(ok, np.trunc(np.where(ok, u, 0)), -1).astype(np.int64)
        iy = np.where(ok, np.trunc(np.where(ok, v, 0)), -1).astype(np.int64)
    H, W = image.shape
    keep &= ok & (ix >= 0) & (iy >= 0) & (ix < W) & (iy < H)
    pixel = image[iy[keep], ix[keep]] / 255.0
    ib = np.clip((pixel * bins).astype(np.int64), 0, bins - 1)
    lb = np.clip((inten[keep] * bins).astype(np.int64), 0, bins - 1)
    hist = np.zeros((bins, bins), dtype=np.int64)
    np.add.at(hist, (ib, lb), 1)
    return hist


def test_oracle_agrees_with_an_independent_numpy_restatement(oracle):
    """Guards the C oracle against transcription slips: a second, vectorised restatement of the same reference lines."""
    for seed, bins in ((1, 16), (2, 8), (3, 32)):
        pr = util.random_problem("plumb_bob", n=40000, seed=seed)
        cam = oracle.create_camera("plumb_bob", pr["intrinsics"], pr["distortion"])
        fov = oracle.estimate_camera_fov(cam, pr["W"], pr["H"])
        for T in util.random_poses(pr["T"], 3, seed=seed):
            _, h = oracle.nid_calculate(cam, pr["image"], pr["points"], pr["intensities"], bins, fov, T)
            h_np = _numpy_nid_plumb_bob(pr["image"], pr["points"], pr["intensities"], pr["intrinsics"], pr["distortion"], bins, fov, T)
            assert np.array_equal(h, h_np)
            # entropy tail, independently
            s = h_np.sum()
            Hf = lambda p: -(p * np.log(p + 1e-6)).sum()  # noqa: E731
            Hr, Hs, Hrs = Hf(h_np.sum(1) / s), Hf(h_np.sum(0) / s), Hf(h_np / s)
            nid_np = (Hrs - (Hr + Hs - Hrs)) / Hrs
            assert abs(nid_np - oracle.nid_from_hist(h)[0]) < 1e-13


def test_mode_b_golden_fixtures(oracle):
    """Committed mode-B vectors (value of the double functor, value + 7 partials of the Jet functor; written from the
    reference functor by tests/golden/make_golden.py).  Runs wherever the oracle builds -- no reference tree needed."""
    import glob

    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mode_b_*.npz")))
    assert len(files) == 6
    for path in files:
        model = os.path.basename(path)[len("mode_b_"):-len(".npz")]
        g = np.load(path)
        cam = oracle.create_camera(model, g["intrinsics"], g["distortion"])
        pts, ins = g["points"].astype(np.float64), g["intensities"].astype(np.float64)
        for k, tp in enumerate(g["T_params"]):
            ok, nid, grad = oracle.nid_cost_bspline_grad(cam, g["image"], pts, ins, 16, tp)
            assert ok and nid == g["nid_jet_functor"][k] and np.array_equal(grad, g["grad"][k])
            assert oracle.nid_cost_bspline(cam, g["image"], pts, ins, 16, tp)[1] == g["nid_double_functor"][k]
