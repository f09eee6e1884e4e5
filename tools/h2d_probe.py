"""Probe: pinned H2D bandwidth of this box vs the e2e upload path (chunked double->float4 conversion + copy)."""
import json
import time

import torch

out = {}
for mb in (2, 16, 64):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    out[f"h2d_pinned_{mb}MB_GBps"] = round(10 * n / (s.elapsed_time(e) * 1e-3) * 1e-9, 2)
# pageable source, as a user's numpy array would be
n = 16 * 1024 * 1024
hp = torch.empty(n, dtype=torch.uint8)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d.copy_(hp)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(hp)
torch.cuda.synchronize()
out["h2d_pageable_16MB_GBps"] = round(10 * n / (time.perf_counter() - t0) * 1e-9, 2)
print(json.dumps(out))
