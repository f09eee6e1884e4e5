"""CPU check of the lean fp32 classifier of the persistent kernel (csrc/lean_filter.cuh) against the exact double path
(csrc/exact_classify.cuh; reference decisions: src/vlcal/calib/cost_calculator_nid.cpp:31-38): tests/cpp/lean_check.cu
is the SAME source the kernel compiles, built host-only.  Every verdict the filter keeps must be the exact path's --
including the reference's truncation toward zero for u, v in (-1, 0) -- on random and on edge-adversarial points
(pixel coordinates nudged to within 1e-3 .. 1e-8 px of an integer).  The GPU suite repeats the check with the device
intrinsics (tests/test_gpu_parity.py::test_filter_error_bound_is_sound)."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "direct_visual_lidar_calibration_b200", "csrc")


@pytest.fixture(scope="module")
def lean_check(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("lean") / "lean_check")
    subprocess.run(
        [nvcc, "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-ffp-contract=off", f"-I{CSRC}", f"-I{os.path.join(ROOT, 'include')}", "-o", exe,
         os.path.join(ROOT, "tests", "cpp", "lean_check.cu")],
        check=True,
    )
    return exe


@pytest.mark.parametrize("adversarial", [1, 0])
def test_lean_classifier_never_contradicts_the_exact_path(lean_check, adversarial):
    out = subprocess.run([lean_check, "60000", str(adversarial)], check=True, capture_output=True, text=True).stdout
    cases = json.loads(out)
    assert len(cases) == 9
    for c in cases:
        assert c["enabled"] == 1, c
        assert c["point_poses"] == 240000
        assert c["mismatches"] == 0, c
        assert c["accepted"] > 50000, c
        # the filter must actually decide most point-poses (natural deferral rate of the C2 camera: ~3 %)
        assert c["uncertain"] / c["point_poses"] < (0.15 if adversarial else 0.08), c
    # the reference truncates toward zero: u in (-1, 0) is column 0 (cost_calculator_nid.cpp:37) -- the strip must be hit
    assert sum(c["accepted_in_minus_one_strip"] for c in cases if c["model"] == 0) > 100
