#!/usr/bin/env python
"""bench.py -- NID cost-evaluations/s on BASELINE.json configs[1] (C2).

Workload (one "step"): ONE inner Nelder-Mead solve of the reference's calibration
(VisualCameraCalibration::estimate_pose_nelder_mead, visual_camera_calibration.cpp:70-139) on a synthetic
Ouster-OS1-64-like 1M-point cloud + 1920x1080 plumb_bob image, from a pose 0.5 deg / 2 cm off ground truth, with the
reference's default parameters (256 inner iterations, 16 bins, step 1e-3).  Metric: NID cost evaluations per second,
counting the evaluations the serial reference would have made (speculatively scored poses are reported separately).

  value : data resident in HBM (culled cloud + image uploaded, cost objects built) -- timed region = the solves
  e2e   : the same step through the host-buffer C ABI (vlcal_estimate_pose_nelder_mead): upload, GPU view culling,
          cost-object construction, solve, result -- host<->device copies inside the timed region
  --impl reference : the CPU oracle (line-by-line restatement of the reference, oracle/vlcal_oracle.c) on a bounded
          sample of the same step (the reference cannot be built as it ships: no Eigen/OpenCV/GTSAM in the image;
          oracle/_ref holds its NID sources compiled against stand-in headers -- used to pin the oracle and timed
          beside it in cpu_baseline.reference_build)

N > 1 (torchrun, one rank per GPU): weak scaling over bags -- rank r owns bag r and the joint objective sum_bags NID
(visual_camera_calibration.cpp:105-110) is formed INSIDE the histogram kernel: the finalizing block of every rank
stores its P candidate scores into every peer's cudaIpc-shared mailbox over NVLink, waits for the peers' and adds them
in rank order (--exchange p2p, default; torch.distributed/NCCL only bootstraps the handles and the timing barrier).
--exchange nccl does one NCCL all-reduce per Nelder-Mead batch from the host callback instead (A/B).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "nid_cost_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "C2: 1M-pt OS1-64-like cloud + 1920x1080 plumb_bob, one estimate_pose_nelder_mead inner solve (<=256 NM iterations, 16 bins) per step"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (0 default, 1 exact-fp64 only)")
    ap.add_argument("--no-tile-order", action="store_true", help="keep the cloud in input order (A/B for the tile-ordered layout)")
    ap.add_argument("--solver", default="auto", choices=["auto", "host", "device"], help="inner-solve loop: host-driven (auto = host; measured faster) or device-resident")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: fused in-kernel peer-memory exchange (default) or NCCL all-reduce per batch")
    ap.add_argument("--ref-iterations", type=int, default=12, help="NM iterations per reference-arm step (bounded sample)")
    return ap.parse_args()


def make_inputs(n_points, bag_index):
    from direct_visual_lidar_calibration_b200 import synthetic as S

    bag = S.make_bag("pinhole_1920x1080", "os1_64", n_points, config_index=1, bag_index=bag_index)
    # every rank must use the same start pose: bag 0's ground truth + the C2 perturbation
    bag["T_init"] = S.perturb(S.gt_T_camera_lidar(), (0.5, 0.5, 0.5), (0.02, 0.02, 0.02))
    return bag


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def oracle_objects(bag):
    from oracle import oracle as O

    cam = O.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
    return O, cam


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle on a bounded sample of the step (first `ref_iterations` NM iterations of the
    same inner solve, view culling included), serial over points exactly like the reference (1 thread for one bag)."""
    if rank != 0:
        return
    bag = make_inputs(args.points, 0)
    O, cam = oracle_objects(bag)
    p = O.default_calib_params()
    p.max_inner_iterations = args.ref_iterations
    bags = [(bag["image"], bag["points"], bag["intensities"])]
    times, evals = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = O.estimate_pose_nelder_mead(cam, bags, bag["T_init"], p)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            evals += r["num_evaluations"]
    total = sum(times)
    value = evals / total
    # the same bounded sample through the reference's OWN visual_camera_calibration.cpp + cost_calculator_nid.cpp (oracle/_ref,
    # compiled against stand-in third-party headers); the port above is the faster of the two and stays the headline
    ref_build = None
    try:
        from oracle import reference as R

        if R.available():
            rcam = R.Camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
            sys.stdout.flush()
            saved = os.dup(1)
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)  # the reference prints "cost:<best>" to stdout (visual_camera_calibration.cpp:115)
            try:
                t0 = time.perf_counter()
                R.calibrate_nelder_mead(rcam, bags, bag["T_init"], max_outer_iterations=1, max_inner_iterations=args.ref_iterations)
                dt = time.perf_counter() - t0
            finally:
                os.dup2(saved, 1)
                os.close(saved)
                os.close(devnull)
            per_step = evals / max(1, args.steps)  # same trajectory as the port (tests/test_reference_pin.py), hence the same count
            ref_build = {"value": per_step / dt, "unit": UNIT, "cores": 1, "ms_per_step": 1e3 * dt,
                         "note": "VisualCameraCalibration::calibrate (1 outer iteration) from the reference's own sources, stand-in Eigen/cv::Mat/GTSAM headers, -O2"}
    except Exception as e:
        ref_build = {"unavailable": repr(e)}
    sample = f"first {args.ref_iterations} Nelder-Mead iterations ({evals // max(1, args.steps)} evaluations) of the C2 inner solve per step, view culling included; serial over points like the reference"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points": args.points, "image": "1920x1080", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "reference_build": ref_build},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(bag, culled_points, culled_intens, max_fov):
    """The oracle timed on this box's host cores on a bounded sample: (i) reference-faithful -- serial over points,
    one thread (the reference only parallelises over bags); (ii) best-effort -- OpenMP over points, all cores."""
    O, cam = oracle_objects(bag)
    T = bag["T_init"]
    O.nid_calculate(cam, bag["image"], culled_points, culled_intens, 16, max_fov, T)  # warm
    n_faithful = 40
    t0 = time.perf_counter()
    for k in range(n_faithful):
        O.nid_calculate(cam, bag["image"], culled_points, culled_intens, 16, max_fov, T)
    t_f = (time.perf_counter() - t0) / n_faithful
    cores = os.cpu_count() or 1
    O.nid_calculate(cam, bag["image"], culled_points, culled_intens, 16, max_fov, T, omp=True)
    n_omp = 100
    t0 = time.perf_counter()
    for k in range(n_omp):
        O.nid_calculate(cam, bag["image"], culled_points, culled_intens, 16, max_fov, T, omp=True)
    t_o = (time.perf_counter() - t0) / n_omp
    n = culled_points.shape[0]
    ref_build = None
    try:  # the reference's own cost_calculator_nid.cpp (oracle/_ref, built where /root/reference exists), same sample
        from oracle import reference as R

        if R.available():
            rcam = R.Camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
            Ts = [T] * 20
            R.nid_calculate(rcam, bag["image"], culled_points, culled_intens, 16, Ts[:1])
            t0 = time.perf_counter()
            R.nid_calculate(rcam, bag["image"], culled_points, culled_intens, 16, Ts)
            t_r = (time.perf_counter() - t0) / len(Ts)
            ref_build = {"value": 1.0 / t_r, "unit": UNIT, "cores": 1, "ms_per_eval": 1e3 * t_r,
                         "note": "CostCalculatorNID::calculate from the reference's own source, compiled against the stand-in Eigen/cv::Mat headers of oracle/ref_standin (-O2, no -march); one thread per bag as in the reference"}
    except Exception as e:  # a missing prebuilt library only removes this cross-check
        ref_build = {"unavailable": repr(e)}
    return {
        "value": 1.0 / t_f, "unit": UNIT, "cores": 1, "kind": "port", "reference_build": ref_build,
        "sample": f"{n_faithful} evaluations of CostCalculatorNID::calculate on the culled C2 cloud ({n} points), serial over points as in the reference",
        "ms_per_eval": 1e3 * t_f, "mpoints_per_s": n / t_f * 1e-6,
        "best_effort": {"value": 1.0 / t_o, "unit": UNIT, "cores": cores, "note": "OpenMP over points with thread-private histograms -- NOT what the reference does", "ms_per_eval": 1e3 * t_o},
    }


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    import direct_visual_lidar_calibration_b200 as V
    from direct_visual_lidar_calibration_b200 import calibration as VC

    if not os.path.exists(V.library_path()):
        V.build_library()
    if not torch.cuda.is_available() or V.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = local_rank
    V.set_solver_mode({"auto": 0, "host": 1, "device": 2}[args.solver])
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    bag = make_inputs(args.points, rank)
    cam = V.create_camera(bag["camera_model"], bag["intrinsics"], bag["distortion"])
    W, H = bag["width"], bag["height"]
    data = V.VisualLiDARData(bag["image"], bag["points"], bag["intensities"])

    # all-reduce of the per-pose partial sums over ranks (NCCL, one small collective per Nelder-Mead batch)
    red_dev = torch.zeros(16, dtype=torch.float64, device="cuda")
    red_host = torch.zeros(16, dtype=torch.float64).pin_memory()
    n_collectives = [0]

    def allreduce(vals):
        k = vals.shape[0]
        red_host[:k] = torch.from_numpy(vals)
        red_dev[:k].copy_(red_host[:k], non_blocking=True)
        dist.all_reduce(red_dev[:k])
        red_host[:k].copy_(red_dev[:k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        vals[:] = red_host[:k].numpy()
        n_collectives[0] += 1

    px = None
    if world > 1 and args.exchange == "p2p":
        # fused path: the finalizing block of every evaluation exchanges the scores over NVLink peer memory
        from direct_visual_lidar_calibration_b200.distributed import PeerExchange

        ok = 1.0
        try:
            px = PeerExchange(device, rank, world)
        except Exception as e:  # no peer access / IPC on this box: every rank must take the same decision
            print(f"[bench] rank {rank}: peer exchange unavailable ({e})", file=sys.stderr)
            px, ok = None, 0.0
        flag = torch.tensor([ok], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 1.0:
            px.connect_with_torch()
        else:
            if px is not None:
                px.close()
            px = None
            args.exchange = "nccl"
    ar = allreduce if (world > 1 and px is None) else None

    # ---- resident setup (outside the timed region): cull at the start pose, build the cost object ------------
    cull = V.ViewCulling(cam, (W, H), V.ViewCullingParams(True), device=device)
    idx = cull.cull_indices(data.points, bag["T_init"])
    culled = V.VisualLiDARData(bag["image"], data.points[idx], data.intensities[idx])
    cost = V.CostCalculatorNID(cam, culled, V.NIDCostParams(16), device=device)
    cost.set_kernel_variant(args.variant)
    if not args.no_tile_order:
        cost.reorder_for_pose(bag["T_init"])  # same grouping the e2e path gets from its culling pass
    if px is not None:
        cost.attach_peer_exchange(px)
        px.set_default(True)  # cost objects built inside the e2e call attach it too
    n_culled = culled.size()
    params = V.VisualCameraCalibrationParams()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        return VC.estimate_pose_on_costs([cost], bag["T_init"], params, allreduce=ar)

    def e2e_step():
        calib = V.VisualCameraCalibration(cam, [data], params, device=device, allreduce=ar)
        T, r = calib.estimate_pose_nelder_mead(bag["T_init"])
        return T, r, calib.stats

    def timed(fn, steps, warmup, profile):
        for _ in range(warmup):
            fn()
        results, ms = [], []
        if profile:
            cost.set_profiling(True)
            cost.reset_profile()
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)  # L2 flush between steps (outside the per-step event pair)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            results.append(fn())
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        barrier()
        wall = time.perf_counter() - t_wall0
        prof = cost.profile() if profile else None
        if profile:
            cost.set_profiling(False)
        total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        return results, float(total_ms.item()), wall, prof

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res, total_ms, wall, prof = timed(resident_step, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if rank == 0 else None

    evals_ref = sum(r[1]["num_evaluations"] for r in res)            # what the serial reference would evaluate
    evals_cmp = sum(r[1]["num_evaluations_computed"] for r in res)   # poses actually scored
    batches = sum(r[1]["num_batches"] for r in res)
    secs = total_ms * 1e-3
    value = world * evals_ref / secs  # every rank scores its own bag for every evaluation
    mpoints = world * n_culled * evals_cmp / secs * 1e-6

    # ---- e2e: host buffers in, pose out, every step --------------------------------------------------------
    e2e_res, e2e_ms, _, _ = timed(e2e_step, max(3, min(args.steps, 5)), 3, profile=False)
    e2e_steps = len(e2e_res)
    e2e_evals = sum(r[1]["num_evaluations"] for r in e2e_res)
    e2e_cmp = sum(r[1]["num_evaluations_computed"] for r in e2e_res)
    e2e_value = world * e2e_evals / (e2e_ms * 1e-3)
    h2d = 16 * data.size() + W * H          # float4 cloud staging + image, per step
    d2h = 8 * (e2e_cmp // e2e_steps) + 8    # candidate scores per batch + kept-point count

    if rank != 0:
        if world > 1:
            dist.barrier()
            if px is not None:
                px.close()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the batched histogram kernel) -------------------------------------
    peak, peak_src = measured_hbm_peak()
    alg_bytes = 16 * n_culled + W * H  # per launch: one pass over the float4 cloud + the image plane
    k_ms = prof["kernel_ms_total"] / max(1, prof["kernel_launches"])
    achieved = alg_bytes / (k_ms * 1e-3) * 1e-9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath) and args.points == 1_000_000:
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"], tj["source"]
        except Exception:
            pass
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "kernel": "nid_hist_*_kernel", "avg_launch_us": 1e3 * k_ms, "algorithmic_bytes_per_launch": alg_bytes,
        "poses_per_launch": prof["poses_total"] / max(1, prof["kernel_launches"]),
        "kernel_share_of_step": prof["kernel_ms_total"] / total_ms,
        "note": "a 4-pose launch does 4x the ALU work of the byte count; cloud (16 MB) is L2-resident across NM iterations by design",
    }
    # the bound that is actually active: instruction issue.  ncu: 108 warp-instructions per (point, pose) in the hot loop
    # (profiles/README.md); 148 SMs x 4 schedulers x 1 warp-instruction/clk x 32 lanes at the sampled SM clock.
    sm_clock_hz = 1e6 * (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1965.0)
    pp_per_launch = n_culled * roofline["poses_per_launch"]
    alu_peak = 148 * 4 * 32 * sm_clock_hz / 108.0
    roofline["issue_bound"] = {
        "achieved_pointposes_per_s": pp_per_launch / (k_ms * 1e-3), "peak_pointposes_per_s": alu_peak,
        "frac": pp_per_launch / (k_ms * 1e-3) / alu_peak, "instr_per_pointpose": 108,
        "note": "kernel time includes its ~10 us serial tail (merge, finalize, publish) and launch latency; at 5 M points x 8 poses the same kernel reaches 168 G point-poses/s",
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(bag, culled.points, culled.intensities, cost.max_fov)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "dtype_note": "geometry decided in f64 semantics: fp32 filter with a rigorous error bound + exact f64 recheck (integer histograms identical to the all-f64 kernel); histogram int32; entropies f64",
        "config": {
            "workload": WORKLOAD, "points": data.size(), "culled_points": n_culled, "image": f"{W}x{H}", "bags": world, "parallelism": f"bags{world}" if world > 1 else "single", "exchange": (args.exchange if world > 1 else None),
            "l2": "flushed (256 MiB write) between steps; within a step the culled cloud is re-read every NM iteration by the algorithm itself",
            "kernel_variant": args.variant, "solver": args.solver,
        },
        "evals_per_step": evals_ref / args.steps, "evals_computed_per_step": evals_cmp / args.steps, "batches_per_step": batches / args.steps,
        "mpoints_per_s": mpoints, "wall_s_timed_region": wall,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                "host_breakdown_ms": {k: round(float(np.mean([r[2][k] for r in e2e_res])), 3) for k in ("upload_ms", "cull_ms", "solve_ms")}},
        "gpu_launches": int(prof["kernel_launches"]),
        "collectives": n_collectives[0],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "result_nid": res[-1][1]["y"], "nm_iterations": res[-1][1]["num_iterations"],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if px is not None:
            px.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
