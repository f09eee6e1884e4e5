"""torchrun worker: the persistent solve with its in-kernel score exchange over cudaIpc-shared mailboxes
(csrc/nid_persistent.cuh) == the same joint solve with every bag in one process.

   N GPUs :  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/dist_check_pk.py
   1 GPU  :  VLCAL_DIST_SINGLE_GPU=1 python -m torch.distributed.run ... --nproc-per-node 2 tools/dist_check_pk.py
             (both ranks on device 0, handles over gloo: the GPU time-slices the two spinning kernels)

Cases: (a) one bag per rank; (b) two bags per rank (the grid of each launch is partitioned over its local bags);
(c) the mixed BASELINE config-4 dataset: spinning (os1_64) and non-repetitive (avia) bags alternating."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import direct_visual_lidar_calibration_b200 as V
from direct_visual_lidar_calibration_b200 import calibration as VC
from direct_visual_lidar_calibration_b200 import synthetic as S
from direct_visual_lidar_calibration_b200.distributed import PeerExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
single_gpu = os.environ.get("VLCAL_DIST_SINGLE_GPU", "0") == "1"
device = 0 if single_gpu else local
torch.cuda.set_device(device)
if single_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", device))
n_base = int(os.environ.get("VLCAL_DIST_POINTS", "30000"))


def make(kind, b):
    pattern = "frustum" if kind == "small" else ("os1_64" if b % 2 == 0 else "avia")
    cam_key = "pinhole_640x480" if kind == "small" else "pinhole_1920x1080"
    return S.make_bag(cam_key, pattern, n_base + 4000 * b, config_index=9, bag_index=b, scale=0.5)


def gather_equal(vec):
    t = torch.from_numpy(np.asarray(vec, dtype=np.float64))
    if not single_gpu:
        t = t.cuda()
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return all(torch.equal(p, parts[0]) for p in parts)


def run_case(name, kind, bags_per_rank, px):
    n_bags = world * bags_per_rank
    bags = [make(kind, b) for b in range(n_bags)]
    cam = V.create_camera(bags[0]["camera_model"], bags[0]["intrinsics"], bags[0]["distortion"])
    T0 = S.perturb(S.gt_T_camera_lidar(), (0.3, -0.3, 0.3), (0.01, -0.01, 0.01))
    params = V.VisualCameraCalibrationParams()
    params.max_inner_iterations = 40

    def cost_of(b):
        return V.CostCalculatorNID(cam, V.VisualLiDARData(bags[b]["image"], bags[b]["points"], bags[b]["intensities"]), device=device)

    mine = [cost_of(rank * bags_per_rank + k) for k in range(bags_per_rank)]  # rank r owns bags [r*B, (r+1)*B): (rank, bag) order = bag order
    for c in mine:
        c.attach_peer_exchange(px)
    V.set_solver_mode(3)
    trace = []
    Tp, rp = VC.estimate_pose_on_costs(mine, T0, params, callback=lambda T, c: trace.append(c))
    for c in mine:
        c.attach_peer_exchange(None)
    same_ranks = gather_equal(np.concatenate([Tp.reshape(-1), rp["x"], [rp["y"], rp["num_iterations"], rp["num_evaluations"], len(trace)], trace[:8]]))
    ok = same_ranks
    if rank == 0:
        allc = [cost_of(b) for b in range(n_bags)]
        if n_bags > 8:
            V.set_solver_mode(0)  # one launch holds at most 8 bags: the library picks the host loop for this reference solve
        T1, r1 = VC.estimate_pose_on_costs(allc, T0, params)  # all bags in one launch, no exchange
        V.set_solver_mode(1)
        Th, rh = VC.estimate_pose_on_costs(allc, T0, params)  # round-1 host loop: one launch per bag and batch
        V.set_solver_mode(3)
        one_launch = bool(np.array_equal(T1, Tp) and r1["y"] == rp["y"] and r1["num_iterations"] == rp["num_iterations"] and r1["num_evaluations"] == rp["num_evaluations"])
        host_loop = bool(np.array_equal(Th, Tp) and rh["y"] == rp["y"] and rh["num_evaluations"] == rp["num_evaluations"])
        print(f"PK_DIST_CHECK case={name} world={world} bags_per_rank={bags_per_rank} ranks_identical={same_ranks} equals_single_process_one_launch={one_launch} equals_host_loop={host_loop} "
              f"iters={rp['num_iterations']} evals={rp['num_evaluations']} y={rp['y']:.15f}", flush=True)
        ok = ok and one_launch and host_loop
    V.set_solver_mode(0)
    dist.barrier()  # rank 0's verification solves must not overlap the next case's exchange
    return ok


px = PeerExchange(device, rank, world)
px.connect_with_torch()
ok = True
ok = run_case("one_bag_per_rank", "small", 1, px) and ok
ok = run_case("two_bags_per_rank", "small", 2, px) and ok
ok = run_case("c4_mixed_spinning_nonrepetitive", "c4", 1, px) and ok
dist.barrier()
px.close()
dist.destroy_process_group()
if not ok:
    sys.exit(1)
