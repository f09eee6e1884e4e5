"""Coarse pose-grid initial guess (BASELINE.json config 5).

The reference's `initial_guess_auto` needs SuperGlue 2D-3D matches (src/initial_guess_auto.cpp:113-139); with the batched
NID kernel a brute-force alternative becomes cheap: score an SE(3) grid around a rough pose (8 poses per pass over the
cloud), keep the best few, optionally refine each with the Nelder-Mead inner solve.  There is no reference behaviour to
match here beyond the per-pose NID values, which are the ones of CostCalculatorNID::calculate.

Pose sharding over GPUs (axis ii of SURVEY 8e): rank r scores poses r, r+G, ... on its replica of the cloud; the scores
are gathered on the host side -- no collective on the data path."""
from __future__ import annotations

import numpy as np

from .cost import CostCalculatorNID


def pose_grid(T_center: np.ndarray, n_rot=(8, 8, 8), n_trans=(2, 4, 4), rot_half_deg: float = 4.0, trans_half: float = 0.10) -> np.ndarray:
    """T_center * Exp(rot, trans) over a regular grid: prod(n_rot) * prod(n_trans) poses, (P, 4, 4)."""
    from .synthetic import perturb

    axes_r = [np.linspace(-rot_half_deg, rot_half_deg, k) if k > 1 else np.zeros(1) for k in n_rot]
    axes_t = [np.linspace(-trans_half, trans_half, k) if k > 1 else np.zeros(1) for k in n_trans]
    out = []
    for rx in axes_r[0]:
        for ry in axes_r[1]:
            for rz in axes_r[2]:
                for tx in axes_t[0]:
                    for ty in axes_t[1]:
                        for tz in axes_t[2]:
                            out.append(perturb(T_center, (rx, ry, rz), (tx, ty, tz)))
    return np.stack(out)


def score_poses(cost: CostCalculatorNID, poses: np.ndarray, rank: int = 0, world: int = 1) -> np.ndarray:
    """NID of every pose.  One persistent launch scores this rank's whole share of the list (vlcal_nid_score_poses);
    with world > 1 rank r scores poses[r::world] and the full vector is assembled with one all_gather of the device
    tensors (NCCL) -- the only exchange of the whole search."""
    from .cost import score_poses as _score

    mine = poses[rank::world]
    vals = _score([cost], mine) if len(mine) else np.zeros(0)
    if world == 1:
        return vals
    import torch
    import torch.distributed as dist

    per = (len(poses) + world - 1) // world  # ranks at the tail may hold one pose less: pad to a common length
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    local = torch.full((per,), float("nan"), dtype=torch.float64, device=dev)
    local[: len(vals)] = torch.from_numpy(vals).to(dev)
    parts = torch.empty((world, per), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(parts, local)
    parts = parts.cpu().numpy()
    full = np.empty(len(poses))
    for r in range(world):
        k = len(poses[r::world])
        full[r::world] = parts[r, :k]
    return full


def grid_search(cost: CostCalculatorNID, T_center: np.ndarray, top_k: int = 5, rank: int = 0, world: int = 1, **grid_kwargs):
    """Returns (poses[top_k], nid[top_k]) of the best grid poses (NaN scores -- no inliers -- rank last)."""
    poses = pose_grid(T_center, **grid_kwargs)
    nid = score_poses(cost, poses, rank, world)
    order = np.argsort(np.where(np.isnan(nid), np.inf, nid), kind="stable")[:top_k]
    return poses[order], nid[order]
