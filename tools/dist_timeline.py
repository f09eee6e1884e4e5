"""torchrun worker: in-kernel timeline of one batched evaluation with the fused peer exchange attached (all ranks in lockstep)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import direct_visual_lidar_calibration_b200 as V
from direct_visual_lidar_calibration_b200 import synthetic as S
from direct_visual_lidar_calibration_b200.distributed import PeerExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bag = S.make_bag("pinhole_1920x1080", "os1_64", 1_000_000, config_index=1, bag_index=rank)
T0 = S.perturb(S.gt_T_camera_lidar(), (0.5, 0.5, 0.5), (0.02, 0.02, 0.02))
cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
idx = V.ViewCulling(cam, (bag["width"], bag["height"]), device=local).cull_indices(bag["points"], T0)
cost = V.CostCalculatorNID(cam, V.VisualLiDARData(bag["image"], bag["points"][idx], bag["intensities"][idx]), device=local)
cost.reorder_for_pose(T0)
px = PeerExchange(local, rank, world)
px.connect_with_torch()
cost.attach_peer_exchange(px)
rng = np.random.default_rng(0)
poses = np.stack([S.perturb(T0, rng.uniform(-0.05, 0.05, 3), rng.uniform(-0.002, 0.002, 3)) for _ in range(4)])
for _ in range(5):
    cost.calculate_batch(poses)
tl = []
for _ in range(30):
    dist.barrier()
    tl.append(cost.debug_timeline(poses))
med = {k: round(float(np.median([t[k] for t in tl])), 2) for k in tl[0]}
allm = [None] * world
dist.all_gather_object(allm, med)
if rank == 0:
    for r, m in enumerate(allm):
        print(json.dumps({"rank": r, "world": world, "timeline_us": m}))
dist.barrier()
px.close()
dist.destroy_process_group()
