"""Copies one tools/gpu_check.sh run (gpurun_out/<tag>_*) into profiles/r01_final_* and regenerates the figures bench.py
reads from there (ncu raw page of the top kernel -> DRAM traffic per launch).  usage: python tools/refresh_profiles.py r19"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def first_json_line(path):
    for line in open(path):
        if line.startswith("{"):
            return line
    return None


for src, dst in [("bench.json", "bench.json"), ("bench_hostloop.json", "bench_hostloop.json"), ("bench_exact.json", "bench_exact.json"), ("bench_kpt2.json", "bench_kpt2.json"),
                 ("bench_ref.json", "bench_ref.json")]:
    s = os.path.join(G, f"{tag}_{src}")
    if os.path.exists(s) and first_json_line(s):
        open(os.path.join(P, f"r01_final_{dst}"), "w").write(first_json_line(s))
for src, dst in [("launches.csv", "launches.csv"), ("pytest_gpu.log", "pytest_gpu.log"), ("smoke.log", "smoke.log"), ("kernel_scaling.jsonl", "kernel_scaling.jsonl"),
                 ("configs.jsonl", "configs.jsonl"), ("grad_bench.json", "grad_bench.json")]:
    s = os.path.join(G, f"{tag}_{src}")
    if os.path.exists(s) and os.path.getsize(s) > 0:
        shutil.copy(s, os.path.join(P, f"r01_final_{dst}" if dst != "grad_bench.json" else "r01_grad_bench.json"))

rep = os.path.join(G, f"{tag}_nid_hist.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    open(os.path.join(P, "r01_final_nid_hist_filter_raw.csv"), "w").write(raw)
    col = {h: i for i, h in enumerate(hdr)}
    r = body[-1]

    def val(name):
        return float(r[col[name]].replace(",", "")) if name in col else None

    units = rows[1]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(units[col["dram__bytes_read.sum"]], 1)
    wr *= scale.get(units[col["dram__bytes_write.sum"]], 1)
    out = {
        "kernel": r[col["Kernel Name"]][:80], "workload": "bench.py C2 (culled cloud, 1920x1080, ~4 poses/launch)",
        "dram_bytes_read_per_launch": int(rd), "dram_bytes_write_per_launch": int(wr),
        "duration_us_under_ncu": val("gpu__time_duration.sum") and val("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1),
        "source": f"profiles/r01_final_nid_hist_filter_raw.csv (ncu --set full --clock-control none, gpurun {tag})",
        "note": "cold-cache capture under ncu; in the timed region the cloud is L2-resident after the first launch of a solve",
    }
    json.dump(out, open(os.path.join(P, "r01_ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out))
