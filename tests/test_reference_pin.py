"""Pins the oracle (oracle/vlcal_oracle.c) against the REFERENCE's own sources of the path, compiled here from
/root/reference (oracle/ref_shim.cpp + oracle/ref_standin/ -> oracle/_ref/libvlcal_ref.so).

What runs on the reference side is the reference's code: create_camera.cpp + camera/*.hpp, dfo/nelder_mead.hpp,
estimate_fov.cpp, cost_calculator_nid.cpp, view_culling.cpp.  Third-party headers (Eigen, cv::Mat, ...) are stand-ins,
so Eigen's own reduction orders are restated, not pinned (oracle/ref_standin/Eigen/Core lists the conventions); the
comparisons below are therefore bit-exact, and the NID tolerance of the GPU tests (1e-12 on the entropy tail) covers
what the stand-in cannot pin.  CPU-only; skipped where neither /root/reference nor a prebuilt library exists."""
import glob
import os

import numpy as np
import pytest

import util as U
from oracle import oracle as O
from oracle import reference as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ref():
    if R.build() is None:
        pytest.skip("oracle/_ref/libvlcal_ref.so not built and /root/reference not present")
    return R


def cameras(model):
    intr, dist, (W, H) = U.CAMERAS[model]
    return O.create_camera(model, intr, dist), R.Camera(model, intr, dist), W, H


def special_points():
    e = 1e-300
    return np.array(
        [[0, 0, 0], [0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [0, -1, 0], [1e-4, 0, 1e-4], [0, 1e-2, 1e-2], [3, 4, 1e-9], [1e-3, 1e-3, 1.0], [e, e, e], [1e200, 1e200, 1e200],
         [np.nan, 0, 1], [np.inf, 0, 1], [0.5, -0.25, 2.0], [-7.0, 3.0, 0.1], [2.0, 2.0, -0.5], [0.0, 0.0, 1e-320]], dtype=np.float64)


@pytest.mark.parametrize("model", U.MODELS)
def test_projection_bit_exact(ref, model):
    oc, rc, W, H = cameras(model)
    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.normal(size=(30000, 3)) * rng.uniform(0.01, 40.0, (30000, 1)), special_points()])
    uv_o = np.array([O.project(oc, p) for p in pts])
    uv_r = ref.project(rc, pts)
    assert np.array_equal(uv_o.view(np.uint64), uv_r.view(np.uint64)) or np.array_equal(uv_o, uv_r, equal_nan=True)
    finite = np.isfinite(uv_o).all(axis=1)
    assert np.array_equal(uv_o[finite].view(np.uint64), uv_r[finite].view(np.uint64))  # signed zeros included


def test_create_camera_rejections(ref):
    assert ref.Camera("no_such_model", [1, 1, 1, 1], []).handle is None  # create_camera.cpp:49-50
    assert ref.Camera("plumb_bob", [1, 1, 1], []).handle is None  # :19-22
    assert O.create_camera("no_such_model", [1, 1, 1, 1], []) is None
    assert O.create_camera("plumb_bob", [1, 1, 1], []) is None
    # short distortion lists are zero-padded, long ones truncated (:24-27)
    for dist in ([], [-0.04], [-0.04, 0.08, 1e-4, -3e-4, -0.04, 9.0, 9.0]):
        oc, rc = O.create_camera("plumb_bob", [400, 410, 320, 240], dist), ref.Camera("plumb_bob", [400, 410, 320, 240], dist)
        p = np.array([[0.3, -0.2, 1.5]])
        assert np.array_equal(np.array([O.project(oc, p[0])]), ref.project(rc, p))


@pytest.mark.parametrize("model", U.MODELS)
def test_estimate_camera_fov_bit_exact(ref, model):
    oc, rc, W, H = cameras(model)
    assert O.estimate_camera_fov(oc, W, H) == ref.estimate_camera_fov(rc, W, H)
    assert O.estimate_camera_fov(oc, W // 2 + 1, H // 3) == ref.estimate_camera_fov(rc, W // 2 + 1, H // 3)  # odd sizes: integer halves (:37)


def _objectives():
    def rosenbrock(x):
        return float(100.0 * (x[1] - x[0] ** 2) ** 2 + (1.0 - x[0]) ** 2)

    def bowl(x):
        return float(np.sum((x - np.arange(len(x)) * 0.1) ** 2))

    def plateau(x):  # many exact ties: the sort / comparison order decides the trajectory
        return float(np.floor(4.0 * np.abs(x).sum()) / 4.0)

    def with_nan(x):  # NaN scores take the reference's comparison path
        return float("nan") if x[0] > 0.15 else float(np.sum(x * x) + np.sin(5.0 * x[-1]))

    def ridge(x):
        return float(abs(x[0] - x[1]) + 0.01 * np.sum(x * x))

    return {"rosenbrock": rosenbrock, "bowl": bowl, "plateau": plateau, "with_nan": with_nan, "ridge": ridge}


@pytest.mark.parametrize("n", [2, 3, 6])
@pytest.mark.parametrize("name", list(_objectives().keys()))
def test_nelder_mead_trajectory_identical(ref, n, name):
    f = _objectives()[name]
    rng = np.random.default_rng(5 + n)
    for trial, kw in enumerate([dict(), dict(init_step=1e-3, convergence_var_thresh=1e-10, max_iterations=60), dict(init_step=0.5, max_iterations=7)]):
        x0 = rng.uniform(-0.3, 0.3, n) if trial else np.zeros(n)
        a = O.nelder_mead(f, x0, **kw)
        b = ref.nelder_mead(f, x0, **kw)
        assert len(a["calls"]) == len(b["calls"]), (name, n, trial)
        for (xa, ya), (xb, yb) in zip(a["calls"], b["calls"]):
            assert np.array_equal(xa, xb) and (ya == yb or (np.isnan(ya) and np.isnan(yb)))
        assert a["converged"] == b["converged"] and a["num_iterations"] == b["num_iterations"]
        assert np.array_equal(a["x"], b["x"]) and (a["y"] == b["y"] or (np.isnan(a["y"]) and np.isnan(b["y"])))


@pytest.mark.parametrize("model", U.MODELS)
@pytest.mark.parametrize("bins", [16, 8])
def test_nid_calculate_bit_exact(ref, model, bins):
    oc, rc, W, H = cameras(model)
    fov = O.estimate_camera_fov(oc, W, H)
    pr = U.random_problem(model, n=30000, seed=21, f32=(bins == 16))
    Ts = U.random_poses(pr["T"], 4, seed=3)
    got = ref.nid_calculate(rc, pr["image"], pr["points"], pr["intensities"], bins, Ts)
    want = np.array([O.nid_calculate(oc, pr["image"], pr["points"], pr["intensities"], bins, fov, T)[0] for T in Ts])
    assert np.array_equal(got, want), (got, want)


def test_nid_calculate_edge_cases(ref):
    oc, rc, W, H = cameras("plumb_bob")
    fov = O.estimate_camera_fov(oc, W, H)
    pr = U.random_problem("plumb_bob", n=2000, seed=2)
    T = pr["T"]
    # no inliers -> 0/0 (cost_calculator_nid.cpp:54-57): NaN on both sides
    behind = pr["points"].copy()
    behind[:, 0] = -np.abs(behind[:, 0]) - 1.0
    a = ref.nid_calculate(rc, pr["image"], behind, pr["intensities"], 16, [T])[0]
    b = O.nid_calculate(oc, pr["image"], behind, pr["intensities"], 16, fov, T)[0]
    assert np.isnan(a) and np.isnan(b)
    # intensities outside [0, 1), NaN intensity, NaN / huge coordinates
    ins = pr["intensities"].copy()
    ins[::7] = 1.0
    ins[1::7] = -0.5
    ins[2::7] = 3.0
    ins[3::97] = np.nan
    pts = pr["points"].copy()
    pts[5::211, 1] = np.nan
    pts[6::211, 2] = 1e30
    a = ref.nid_calculate(rc, pr["image"], pts, ins, 16, [T])[0]
    b = O.nid_calculate(oc, pr["image"], pts, ins, 16, fov, T)[0]
    assert a == b
    # a single point
    a = ref.nid_calculate(rc, pr["image"], pr["points"][:1], pr["intensities"][:1], 16, [T])[0]
    b = O.nid_calculate(oc, pr["image"], pr["points"][:1], pr["intensities"][:1], 16, fov, T)[0]
    assert a == b or (np.isnan(a) and np.isnan(b))


@pytest.mark.parametrize("model", U.MODELS)
@pytest.mark.parametrize("depth", [True, False])
def test_view_culling_indices_identical(ref, model, depth):
    oc, rc, W, H = cameras(model)
    fov = O.estimate_camera_fov(oc, W, H)
    pr = U.random_problem(model, n=40000, seed=33)
    for T in U.random_poses(pr["T"], 2, seed=8):
        assert np.array_equal(O.view_cull(oc, W, H, fov, depth, pr["points"], T), ref.view_cull(rc, W, H, depth, pr["points"], T))


def test_golden_fixtures_match_the_reference(ref):
    """The committed fixtures (made from the oracle, tests/golden/make_golden.py) are what the reference's code returns."""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "mode_a_*.npz")))
    assert len(files) == len(U.MODELS)
    for path in files:
        model = os.path.basename(path)[len("mode_a_"):-len(".npz")]
        g = np.load(path)
        H, W = g["image"].shape
        rc = ref.Camera(model, g["intrinsics"], g["distortion"])
        assert ref.estimate_camera_fov(rc, W, H) == float(g["max_fov"])
        pts, ins = g["points"].astype(np.float64), g["intensities"].astype(np.float64)
        assert np.array_equal(ref.nid_calculate(rc, g["image"], pts, ins, 16, g["poses"]), g["nid"])
        assert np.array_equal(ref.view_cull(rc, W, H, True, pts, g["poses"][0]), g["cull_indices"])


def _sophus_params(T):
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(T[:3, :3]).as_quat()  # x y z w
    return np.concatenate([q, T[:3, 3]])


@pytest.mark.parametrize("model", U.MODELS)
def test_bspline_nid_value_bit_exact(ref, model):
    """NIDCost::operator()<double> (include/vlcal/costs/nid_cost.hpp:36-107), the value half of the BFGS branch."""
    oc, rc, W, H = cameras(model)
    pr = U.random_problem(model, n=8000, seed=71)
    for bins, T in zip((16, 16, 8), U.random_poses(pr["T"], 3, seed=5)):
        tp = _sophus_params(T)
        ok_r, nid_r = ref.nid_cost_bspline(rc, pr["image"], pr["points"], pr["intensities"], bins, tp)
        ok_o, nid_o, _ = O.nid_cost_bspline(oc, pr["image"], pr["points"], pr["intensities"], bins, tp)
        assert ok_r and ok_o and nid_r == nid_o, (nid_r, nid_o)


def test_bspline_failure_flag(ref):
    oc, rc, W, H = cameras("plumb_bob")
    pr = U.random_problem("plumb_bob", n=100, seed=72)
    far = np.array([[1.0, 5e3, 0.0, 1.0]] * 10)  # far off to the side: every projection lands outside the image
    tp = _sophus_params(pr["T"])
    ok_r, _ = ref.nid_cost_bspline(rc, pr["image"], far, np.full(10, 0.5), 16, tp)
    ok_o, _, _ = O.nid_cost_bspline(oc, pr["image"], far, np.full(10, 0.5), 16, tp)
    assert ok_r is False and ok_o is False  # no inliers -> NaN -> the functor returns false (nid_cost.hpp:98-102)


def _best_cost_poses(init_T, trace):
    """Poses the reference hands to params.callback: evaluations that improve on the best cost so far (:112-116)."""
    best, out = np.finfo(np.float64).max, []
    for row in trace:
        if row[6] < best:
            best = row[6]
            out.append(O.isometry_mul(init_T, O.se3_expmap(row[:6])))
    return out


@pytest.mark.parametrize("model,n_bags", [("plumb_bob", 1), ("plumb_bob", 2), ("fisheye", 1), ("equirectangular", 2)])
def test_estimate_pose_nelder_mead_identical(ref, model, n_bags):
    """VisualCameraCalibration::estimate_pose_nelder_mead (visual_camera_calibration.cpp:70-139) through calibrate() with
    one outer iteration: culling at the start pose, one cost object per bag, objective sum in bag order, best-cost
    callbacks, result init_T * Expmap(x).  (GTSAM's Expmap and Eigen's Isometry product are stand-ins on the reference
    side -- restated like the oracle's; everything else is the reference's code.)"""
    oc, rc, W, H = cameras(model)
    bags = []
    for b in range(n_bags):
        pr = U.random_problem(model, n=6000 + 500 * b, seed=90 + b)
        bags.append((pr["image"], pr["points"], pr["intensities"]))
    T0 = U.random_poses(pr["T"], 1, seed=12, rot_deg=0.5, trans=0.02)[0]
    p = O.default_calib_params()
    p.max_inner_iterations = 40
    p.max_outer_iterations = 1
    a = O.estimate_pose_nelder_mead(oc, bags, T0, p)
    b = ref.calibrate_nelder_mead(rc, bags, T0, max_outer_iterations=1, max_inner_iterations=40)
    assert np.array_equal(a["T"], b["T"])
    want = _best_cost_poses(T0, a["trace"])
    assert b["num_callbacks"] == len(want)
    for Tw, Tg in zip(want, b["callback_T"]):
        assert np.array_equal(Tw, Tg)


def test_calibrate_outer_loop_identical(ref):
    """VisualCameraCalibration::calibrate (visual_camera_calibration.cpp:35-68): re-culling at every outer iteration and
    the delta_t / delta_r termination test."""
    oc, rc, W, H = cameras("plumb_bob")
    pr = U.random_problem("plumb_bob", n=8000, seed=95)
    bags = [(pr["image"], pr["points"], pr["intensities"])]
    T0 = U.random_poses(pr["T"], 1, seed=13, rot_deg=0.5, trans=0.02)[0]
    for kw in (dict(max_outer_iterations=3, max_inner_iterations=25, delta_trans_thresh=1e-9, delta_rot_thresh=1e-9),  # never converges: 3 outer iterations
               dict(max_outer_iterations=5, max_inner_iterations=25),  # default thresholds: stops after the first
               dict(max_outer_iterations=2, max_inner_iterations=30, disable_z_buffer_culling=True, nelder_mead_init_step=5e-3)):
        p = O.default_calib_params()
        p.max_outer_iterations, p.max_inner_iterations = kw["max_outer_iterations"], kw["max_inner_iterations"]
        p.delta_trans_thresh = kw.get("delta_trans_thresh", p.delta_trans_thresh)
        p.delta_rot_thresh = kw.get("delta_rot_thresh", p.delta_rot_thresh)
        p.disable_z_buffer_culling = int(kw.get("disable_z_buffer_culling", False))
        p.nelder_mead_init_step = kw.get("nelder_mead_init_step", p.nelder_mead_init_step)
        a = O.calibrate(oc, bags, T0, p)
        b = ref.calibrate_nelder_mead(rc, bags, T0, **kw)
        assert np.array_equal(a["T"], b["T"]), kw


@pytest.mark.parametrize("model", U.MODELS)
def test_bspline_nid_gradient_bit_exact(ref, model):
    """NIDCost::operator()<ceres::Jet<double, 7>> (what AutoDiffFirstOrderFunction evaluates in the BFGS branch,
    visual_camera_calibration.cpp:211): residual and its 7 partials, reference functor + stand-in Jet vs the oracle."""
    oc, rc, W, H = cameras(model)
    pr = U.random_problem(model, n=6000, seed=75)
    for bins, T in zip((16, 8), U.random_poses(pr["T"], 2, seed=6)):
        tp = _sophus_params(T)
        ok_r, nid_r, g_r = ref.nid_cost_bspline_jet(rc, pr["image"], pr["points"], pr["intensities"], bins, tp)
        ok_o, nid_o, g_o = O.nid_cost_bspline_grad(oc, pr["image"], pr["points"], pr["intensities"], bins, tp)
        assert ok_r and ok_o and nid_r == nid_o and np.array_equal(g_r, g_o), (nid_r, nid_o, g_r, g_o)
        assert np.abs(g_o).max() > 1e-4


def test_bspline_gradient_is_the_derivative_of_the_value(ref):
    """Independent of any Jet arithmetic: central differences of the double functor on a smooth image."""
    oc, rc, W, H = cameras("fisheye")
    pr = U.random_problem("fisheye", n=6000, seed=76)
    yy, xx = np.mgrid[0:H, 0:W]
    img = (127 + 100 * np.sin(xx / 37.0) * np.cos(yy / 23.0)).astype(np.uint8)
    tp = _sophus_params(pr["T"])
    _, _, g = O.nid_cost_bspline_grad(oc, img, pr["points"], pr["intensities"], 16, tp)
    fd = np.zeros(7)
    for k in range(7):
        a, b = tp.copy(), tp.copy()
        a[k] += 1e-6
        b[k] -= 1e-6
        fd[k] = (ref.nid_cost_bspline(rc, img, pr["points"], pr["intensities"], 16, a)[1] - ref.nid_cost_bspline(rc, img, pr["points"], pr["intensities"], 16, b)[1]) / 2e-6
    assert np.abs(g - fd).max() < 1e-5 * max(1.0, np.abs(g).max()), (g, fd)


def _random_camera(model, rng):
    W, H = int(rng.integers(64, 400)), int(rng.integers(48, 300))
    f = rng.uniform(0.4, 1.6) * W
    intr = [f, f * rng.uniform(0.9, 1.1), W * rng.uniform(0.4, 0.6), H * rng.uniform(0.4, 0.6)]
    if model == "plumb_bob":
        dist = list(rng.normal(0, [0.05, 0.05, 1e-3, 1e-3, 0.02]))
    elif model == "fisheye":
        dist = list(rng.normal(0, [0.02, 0.01, 0.005, 0.002]))
    elif model == "atan":
        dist = [float(rng.choice([0.0, 1e-8, rng.uniform(0.2, 1.2)]))]  # includes the d0 < 1e-7 branch (atan.hpp:17)
    elif model == "omnidir":
        intr = intr + [rng.uniform(0.5, 1.5)]
        dist = list(rng.normal(0, [0.05, 0.02, 1e-3, 1e-3]))
    elif model == "equirectangular":
        intr, dist = [float(W), float(H)], []
    else:
        dist = list(rng.normal(0, [0.05, 0.05, 1e-3, 1e-3, 0.02, 0.05, 0.03, 0.01]))
    return intr, dist, W, H


@pytest.mark.parametrize("model", U.MODELS)
def test_fuzz_random_cameras_projection_fov_nid_culling(ref, model):
    """Random intrinsics / distortions / image sizes (the fixed presets above could hide a branch): projection, FoV, NID and
    culling of the oracle against the reference's code, bit for bit."""
    rng = np.random.default_rng(1000 + U.MODELS.index(model))
    for trial in range(8):
        intr, dist, W, H = _random_camera(model, rng)
        oc, rc = O.create_camera(model, intr, dist), ref.Camera(model, intr, dist)
        pts = rng.normal(size=(3000, 3)) * rng.uniform(0.05, 30.0, (3000, 1))
        assert np.array_equal(np.array([O.project(oc, p) for p in pts]), ref.project(rc, pts), equal_nan=True), (model, trial)
        fov = O.estimate_camera_fov(oc, W, H)
        assert fov == ref.estimate_camera_fov(rc, W, H), (model, trial, intr, dist)
        pr = U.random_problem(model, n=4000, seed=2000 + trial, size=(W, H))
        Ts = U.random_poses(pr["T"], 2, seed=trial, rot_deg=4.0, trans=0.3)
        bins = int(rng.choice([4, 16, 64]))
        got = ref.nid_calculate(rc, pr["image"], pr["points"], pr["intensities"], bins, Ts)
        want = np.array([O.nid_calculate(oc, pr["image"], pr["points"], pr["intensities"], bins, fov, T)[0] for T in Ts])
        assert np.array_equal(got, want, equal_nan=True), (model, trial, got, want)
        assert np.array_equal(O.view_cull(oc, W, H, fov, True, pr["points"], Ts[0]), ref.view_cull(rc, W, H, True, pr["points"], Ts[0]))


def test_mode_b_golden_fixtures_match_reference_and_oracle(ref):
    """tests/golden/mode_b_*.npz were written from the reference functor (make_golden.py); both the reference build and
    the oracle must reproduce them exactly."""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "mode_b_*.npz")))
    assert len(files) == len(U.MODELS)
    for path in files:
        model = os.path.basename(path)[len("mode_b_"):-len(".npz")]
        g = np.load(path)
        rc, oc = ref.Camera(model, g["intrinsics"], g["distortion"]), O.create_camera(model, g["intrinsics"], g["distortion"])
        pts, ins = g["points"].astype(np.float64), g["intensities"].astype(np.float64)
        for k, tp in enumerate(g["T_params"]):
            ok_r, nid_r, grad_r = ref.nid_cost_bspline_jet(rc, g["image"], pts, ins, 16, tp)
            ok_o, nid_o, grad_o = O.nid_cost_bspline_grad(oc, g["image"], pts, ins, 16, tp)
            assert ok_r and ok_o and nid_r == nid_o == g["nid_jet_functor"][k]
            assert np.array_equal(grad_r, g["grad"][k]) and np.array_equal(grad_o, g["grad"][k])
            assert ref.nid_cost_bspline(rc, g["image"], pts, ins, 16, tp)[1] == g["nid_double_functor"][k] == O.nid_cost_bspline(oc, g["image"], pts, ins, 16, tp)[1]


@pytest.mark.parametrize("model", U.MODELS)
def test_generate_lidar_image_equals_the_reference(model):
    """generate_lidar_image (src/vlcal/preprocess/generate_lidar_image.cpp:8-41): the oracle's intensity image and index map
    equal the reference's own code bit for bit, including the tie rule (of equal squared ranges the last point wins)."""
    pr = U.random_problem(model, n=30000, seed=5)
    pts = np.concatenate([pr["points"], pr["points"][:2000]])  # exact duplicates: equal squared ranges
    ins = np.concatenate([pr["intensities"], (pr["intensities"][:2000] + 0.5) % 1.0])
    cam = O.create_camera(model, pr["intrinsics"], pr["distortion"])
    rcam = R.Camera(model, pr["intrinsics"], pr["distortion"])
    a = O.generate_lidar_image(cam, pr["W"], pr["H"], pr["T"], pts, ins)
    b = R.generate_lidar_image(rcam, pr["W"], pr["H"], pr["T"], pts, ins)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (a[1] >= 30000).sum() > 100  # duplicates won their pixels
