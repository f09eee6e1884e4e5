"""On-disk formats either side of the hot path: what the reference's `preprocess` writes and what
`calibrate` reads / writes.

  <data_path>/<bag>.png   first camera image, grayscale uint8                       (preprocess.cpp:160-161, visual_lidar_data.cpp:13)
  <data_path>/<bag>.ply   binary PLY, float x y z + float intensity per vertex      (preprocess.cpp:163-169, visual_lidar_data.cpp:19-26)
  <data_path>/calib.json  camera.{camera_model,intrinsics,distortion_coeffs}, meta.bag_names,
                          results.{init_T_lidar_camera[_auto], T_lidar_camera} = [x y z qx qy qz qw] of T_lidar_camera
                          (preprocess.cpp:220-232, calibrate.cpp:38-45,57-76,128-140)
The PLY writer/reader of the reference is Iridescence's glk::save_ply_binary / glk::load_ply (not vendored); the reader here
parses the header generically (binary little endian or ascii, any property order, float/double/int types)."""
from __future__ import annotations

import json
import os

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


def save_ply_binary(path: str, points_xyz, intensities=None) -> None:
    """Binary little-endian PLY with float x, y, z (+ float intensity), the layout of preprocess.cpp:163-169."""
    pts = np.asarray(points_xyz, dtype=np.float32).reshape(-1, 3)
    cols = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    if intensities is not None:
        cols.append(("intensity", "<f4"))
    rec = np.empty(pts.shape[0], dtype=cols)
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if intensities is not None:
        rec["intensity"] = np.asarray(intensities, dtype=np.float32).reshape(-1)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {pts.shape[0]}"]
    header += [f"property float {name}" for name, _ in cols]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rec.tobytes())


def load_ply(path: str):
    """Returns (points float64 (N,3), intensities float64 (N,) or None).  Values are the file's floats widened to double,
    exactly what FrameCPU(ply->vertices) + add_intensities do (frame_cpu.cpp:79-86,127-132)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n_vertex, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", errors="replace").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n_vertex = int(tok[2])
                elif props and n_vertex:
                    pass  # elements after the vertices are ignored
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=n_vertex, ndmin=2) if n_vertex else np.zeros((0, len(props)))
            col = {n: data[:, i] for i, n in enumerate(names)}
        elif fmt in ("binary_little_endian", "binary_big_endian"):
            order = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(n, order + t) for n, t in props])
            rec = np.frombuffer(f.read(dt.itemsize * n_vertex), dtype=dt, count=n_vertex)
            col = {n: rec[n] for n in names}
        else:
            raise ValueError(f"{path}: unknown PLY format {fmt}")
    pts = np.stack([np.asarray(col[k], dtype=np.float64) for k in ("x", "y", "z")], axis=1)
    inten = np.asarray(col["intensity"], dtype=np.float64) if "intensity" in col else None
    return np.ascontiguousarray(pts), inten


def load_image_gray(path: str) -> np.ndarray:
    """cv::imread(path, 0)."""
    import cv2

    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(f"failed to load {path}")
    return img


def load_visual_lidar_data(data_path: str, bag_name: str):
    """vlcal::VisualLiDARData(data_path, bag_name) (visual_lidar_data.cpp:10-27)."""
    from .cost import VisualLiDARData

    image = load_image_gray(os.path.join(data_path, bag_name + ".png"))
    pts, inten = load_ply(os.path.join(data_path, bag_name + ".ply"))
    if inten is None:
        raise ValueError(f"{bag_name}.ply has no intensity property")
    return VisualLiDARData(image, pts, inten)


def load_calib_json(data_path: str) -> dict:
    with open(os.path.join(data_path, "calib.json")) as f:
        return json.load(f)


def save_calib_json(data_path: str, config: dict) -> None:
    with open(os.path.join(data_path, "calib.json"), "w") as f:
        f.write(json.dumps(config, indent=2) + "\n")  # config.dump(2) << std::endl  (calibrate.cpp:140)


# ---- pose <-> [x y z qx qy qz qw] (TUM order) of T_lidar_camera -------------------------------------------------------

def quat_to_matrix(qx, qy, qz, qw) -> np.ndarray:
    """Eigen::Quaterniond(w,x,y,z).normalized().toRotationMatrix() (calibrate.cpp:72)."""
    q = np.array([qx, qy, qz, qw], dtype=np.float64)
    q = q / np.linalg.norm(q)
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def matrix_to_quat(R) -> np.ndarray:
    """Eigen::Quaterniond(Matrix3d) -> (x, y, z, w)."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    q = np.empty(4)
    if t > 0:
        s = np.sqrt(t + 1.0)
        q[3] = 0.5 * s
        s = 0.5 / s
        q[0], q[1], q[2] = (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * s
        s = 0.5 / s
        q[3] = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
    return q


def tum_to_T(values) -> np.ndarray:
    T = np.eye(4)
    T[:3, 3] = values[0:3]
    T[:3, :3] = quat_to_matrix(values[3], values[4], values[5], values[6])
    return T


def T_to_tum(T) -> list:
    q = matrix_to_quat(np.asarray(T)[:3, :3])
    t = np.asarray(T)[:3, 3]
    return [float(t[0]), float(t[1]), float(t[2]), float(q[0]), float(q[1]), float(q[2]), float(q[3])]


def invert_isometry(T) -> np.ndarray:
    T = np.asarray(T, dtype=np.float64)
    out = np.eye(4)
    out[:3, :3] = T[:3, :3].T
    out[:3, 3] = -T[:3, :3].T @ T[:3, 3]
    return out
