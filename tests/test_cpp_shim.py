"""The reference-side C++ binding (include/vlcal_b200/cost_calculator_nid_cuda.hpp) compiles against stand-in
reference headers, links to the C-ABI library, and behaves like a vlcal::CostCalculator."""
import os
import struct
import subprocess

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _build(vlcal, tmp_path):
    exe = str(tmp_path / "cpp_shim")
    libdir = os.path.dirname(vlcal.library_path())
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(HERE, "cpp_stubs"), "-I", os.path.join(ROOT, "include"),
           os.path.join(HERE, "cpp_shim_main.cpp"), "-o", exe, "-L", libdir, "-lvlcal_nid", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    return exe


def _write_problem(path, pr):
    with open(path, "wb") as f:
        f.write(struct.pack("iii", pr["W"], pr["H"], pr["points"].shape[0]))
        f.write(pr["image"].tobytes())
        f.write(np.ascontiguousarray(pr["points"]).tobytes())
        f.write(np.ascontiguousarray(pr["intensities"]).tobytes())
        f.write(np.ascontiguousarray(pr["T"].T).tobytes())  # column-major


def test_shim_compiles_and_fails_loudly_without_gpu(vlcal, tmp_path):
    exe = _build(vlcal, tmp_path)
    if vlcal.device_count() > 0:
        pytest.skip("a GPU is present; covered by the gpu test")
    pr = util.random_problem("plumb_bob", n=100, seed=1)
    _write_problem(tmp_path / "p.bin", pr)
    out = subprocess.run([exe, str(tmp_path / "p.bin")], capture_output=True, text=True)
    assert out.returncode == 3 and "no CUDA device" in out.stdout


@pytest.mark.gpu
def test_shim_matches_python_surface_and_oracle(gpu, oracle, tmp_path):
    exe = _build(gpu, tmp_path)
    pr = util.random_problem("plumb_bob", n=20000, seed=2)
    _write_problem(tmp_path / "p.bin", pr)
    out = subprocess.run([exe, str(tmp_path / "p.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    vals = [float(v) for v in out.stdout.split()[1:]]
    cam = gpu.create_camera("plumb_bob", pr["intrinsics"], pr["distortion"])
    ref = gpu.CostCalculatorNID(cam, gpu.VisualLiDARData(pr["image"], pr["points"], pr["intensities"])).calculate(pr["T"])
    assert vals[0] == ref and vals[1] == ref and vals[2] == ref
    assert vals[3] == ref  # CostCalculatorNIDCuda(proj, data, params): the reference's constructor shape, through CameraParamsView
    assert vals[4] == 1.0 and vals[5] == 1.0  # a camera without the view throws; unknown model -> nullptr like the reference
    ocam = oracle.create_camera("plumb_bob", pr["intrinsics"], pr["distortion"])
    onid = oracle.nid_calculate(ocam, pr["image"], pr["points"], pr["intensities"], 16, oracle.estimate_camera_fov(ocam, pr["W"], pr["H"]), pr["T"])[0]
    assert abs(vals[0] - onid) < 1e-12


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "include")), reason="needs the reference tree's headers (this container only)")
def test_bindings_compile_inside_the_reference_tree(vlcal, tmp_path):
    """Both reference-side bindings against the reference's OWN headers (third-party ones from oracle/ref_standin):
    cost_calculator_nid_cuda.hpp (Nelder-Mead branch) and nid_cost_cuda.hpp (BFGS branch, a ceres::FirstOrderFunction)."""
    exe = str(tmp_path / "cpp_bfgs")
    libdir = os.path.dirname(vlcal.library_path())
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(ROOT, "oracle", "ref_standin"), "-I", os.path.join(REFERENCE, "include"), "-I", os.path.join(ROOT, "include"),
           os.path.join(HERE, "cpp_bfgs_main.cpp"), os.path.join(REFERENCE, "src", "camera", "create_camera.cpp"), "-o", exe, "-L", libdir, "-lvlcal_nid", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    if vlcal.device_count() > 0:
        assert out.returncode == 0 and out.stdout.startswith("EVAL ok="), out.stdout + out.stderr
    else:  # no CPU fallback: the constructor throws with the library's message
        assert out.returncode == 3 and "EXCEPTION" in out.stdout, out.stdout + out.stderr
