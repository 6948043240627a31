#!/usr/bin/env python
"""Measure the TF32 dense tensor-core peak of this box with cuBLAS (torch.matmul, allow_tf32) the way the driver
measured bf16 for MEASURED_PEAKS.json: 8192^3, best of 10 (burst) and back to back for 4 s (sustained).
The fp32-emulated (3xTF32) roofline denominator is this number / 3.  Measurement tool only."""
import json
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
flops = 2.0 * n ** 3
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 0
while time.time() - t0 < 4.0:
    for _ in range(20):
        a @ b
    iters += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sustained = flops * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
torch.backends.cuda.matmul.allow_tf32 = False
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    a @ b
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    a @ b
e1.record()
torch.cuda.synchronize()
fp32 = flops * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps({"tf32_tflops_burst": flops / (best * 1e-3) / 1e12, "tf32_tflops_sustained": sustained,
                  "emulated_fp32_peak_burst": flops / (best * 1e-3) / 1e12 / 3, "emulated_fp32_peak_sustained": sustained / 3,
                  "cublas_fp32_sgemm_tflops": fp32, "n": n}))
