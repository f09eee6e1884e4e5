"""Generates the golden fixtures in this directory FROM THE ORACLE (oracle/vlcal_oracle.c).

The reference ships no golden vectors, so these pins are ours: they freeze the oracle's answers on
seeded inputs so that (a) a later edit of the oracle that changes behaviour is caught by the CPU suite and (b) the
GPU suite can compare against committed numbers, not only against a checker built in the same run.
Where oracle/_ref (the reference's own sources compiled here, oracle/ref_shim.cpp) is available, every stored max_fov /
nid / cull_indices is checked against the reference's code before it is written; tests/test_reference_pin.py repeats
that check on the committed files.
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as O  # noqa: E402
from oracle import reference as R  # noqa: E402
import util  # noqa: E402


def main():
    # 1. hand-derived known-answer vectors for the entropy/NID tail (cost_calculator_nid.cpp:54-64), SURVEY.md 8c,
    #    computed independently in float64 with numpy (not with the oracle)
    def nid_np(h):
        h = np.asarray(h, dtype=np.float64)
        s = h.sum()
        H = lambda p: -(p * np.log(p + 1e-6)).sum()  # noqa: E731
        Hr, Hs, Hrs = H(h.sum(1) / s), H(h.sum(0) / s), H(h / s)
        MI = Hr + Hs - Hrs
        return {"Hr": Hr, "Hs": Hs, "Hrs": Hrs, "MI": MI, "NID": (Hrs - MI) / Hrs}

    kats = []
    for name, h in [("diag2", [[2, 0], [0, 2]]), ("ones2", [[1, 1], [1, 1]]), ("mixed2", [[3, 1], [0, 4]]), ("10I16", (10 * np.eye(16, dtype=int)).tolist()), ("ones16", np.ones((16, 16), dtype=int).tolist())]:
        kats.append({"name": name, "hist": h, **nid_np(h)})
    with open(os.path.join(HERE, "nid_kat.json"), "w") as f:
        json.dump(kats, f, indent=1)

    # 2. per-camera-model mode-A fixtures: small seeded problem, 3 poses, oracle histograms + NID
    for model in util.MODELS:
        pr = util.random_problem(model, n=3000, seed=100 + util.MODELS.index(model), size=(160, 120) if model != "equirectangular" else (160, 80))
        intr = list(pr["intrinsics"])
        if model == "equirectangular":
            intr = [160.0, 80.0]
        else:
            intr = [v * 0.25 for v in intr]
        cam = O.create_camera(model, intr, pr["distortion"])
        fov = O.estimate_camera_fov(cam, pr["W"], pr["H"])
        Ts = util.random_poses(pr["T"], 3, seed=7)
        nids, hists = [], []
        for T in Ts:
            nid, h = O.nid_calculate(cam, pr["image"], pr["points"], pr["intensities"], 16, fov, T)
            nids.append(nid)
            hists.append(h)
        idx = O.view_cull(cam, pr["W"], pr["H"], fov, True, pr["points"], Ts[0])
        if R.build() is not None:
            rcam = R.Camera(model, intr, pr["distortion"])
            pts32 = pr["points"].astype(np.float32).astype(np.float64)
            assert np.array_equal(pts32, pr["points"]) and R.estimate_camera_fov(rcam, pr["W"], pr["H"]) == fov
            assert np.array_equal(R.nid_calculate(rcam, pr["image"], pr["points"], pr["intensities"].astype(np.float32).astype(np.float64), 16, Ts), np.array(nids))
            assert np.array_equal(R.view_cull(rcam, pr["W"], pr["H"], True, pr["points"], Ts[0]), idx)
        np.savez_compressed(
            os.path.join(HERE, f"mode_a_{model}.npz"),
            intrinsics=np.array(intr), distortion=np.array(pr["distortion"], dtype=np.float64), image=pr["image"],
            points=pr["points"].astype(np.float32), intensities=pr["intensities"].astype(np.float32), poses=Ts, max_fov=fov,
            nid=np.array(nids), hist=np.stack(hists).astype(np.int32), cull_indices=idx.astype(np.int32),
        )
    # 3. mode-B fixtures (NIDCost value + gradient), FROM THE REFERENCE'S OWN FUNCTOR when oracle/_ref is available
    #    (include/vlcal/costs/nid_cost.hpp instantiated with double and with Jets, oracle/ref_shim.cpp); the oracle must agree
    #    bit for bit before anything is written
    from scipy.spatial.transform import Rotation

    for model in util.MODELS:
        pr = util.random_problem(model, n=2500, seed=300 + util.MODELS.index(model), size=(160, 120) if model != "equirectangular" else (160, 80))
        intr = [160.0, 80.0] if model == "equirectangular" else [v * 0.25 for v in pr["intrinsics"]]
        cam = O.create_camera(model, intr, pr["distortion"])
        Ts = util.random_poses(pr["T"], 2, seed=11)
        tps = np.stack([np.concatenate([Rotation.from_matrix(T[:3, :3]).as_quat(), T[:3, 3]]) for T in Ts])
        pts = pr["points"].astype(np.float32).astype(np.float64)
        ins = pr["intensities"].astype(np.float32).astype(np.float64)
        vals, grads = [], []
        for tp in tps:
            ok, nid, grad = O.nid_cost_bspline_grad(cam, pr["image"], pts, ins, 16, tp)
            ok_v, nid_v, _ = O.nid_cost_bspline(cam, pr["image"], pts, ins, 16, tp)
            assert ok and ok_v
            if R.build() is not None:
                rcam = R.Camera(model, intr, pr["distortion"])
                ok_r, nid_r, grad_r = R.nid_cost_bspline_jet(rcam, pr["image"], pts, ins, 16, tp)
                ok_d, nid_d = R.nid_cost_bspline(rcam, pr["image"], pts, ins, 16, tp)
                assert ok_r and ok_d and nid_r == nid and np.array_equal(grad_r, grad) and nid_d == nid_v
            vals.append((nid_v, nid))
            grads.append(grad)
        np.savez_compressed(
            os.path.join(HERE, f"mode_b_{model}.npz"),
            intrinsics=np.array(intr), distortion=np.array(pr["distortion"], dtype=np.float64), image=pr["image"], points=pts.astype(np.float32), intensities=ins.astype(np.float32),
            T_params=tps, nid_double_functor=np.array([v[0] for v in vals]), nid_jet_functor=np.array([v[1] for v in vals]), grad=np.stack(grads),
        )
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
