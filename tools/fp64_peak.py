"""FP64 tensor-core peak of this B200 as a kept artefact (MEASURED_PEAKS.json has no FP64 entry):
cuBLAS DGEMM 8192^3 through torch.matmul, best of 10 (burst) and back to back for 4 s (sustained),
with SM clocks / power / throttle reasons sampled during the sustained run.
usage: python tools/fp64_peak.py > profiles/r02_fp64_peak.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import ClockSampler

n = 8192
dev = torch.device("cuda", 0)
a = torch.randn(n, n, dtype=torch.float64, device=dev)
b = torch.randn(n, n, dtype=torch.float64, device=dev)
for _ in range(3):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e30
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
sampler = ClockSampler(0)
sampler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0, reps = time.perf_counter(), 0
e0.record()
while time.perf_counter() - t0 < 4.0:
    for _ in range(10):
        torch.matmul(a, b)
    reps += 10
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
clocks = sampler.stop()
flops = 2.0 * n ** 3
print(json.dumps({"what": "cuBLAS DGEMM %d^3 via torch.matmul (float64)" % n,
                  "fp64_tflops_burst": flops / (best * 1e-3) / 1e12,
                  "fp64_tflops_sustained_4s": flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                  "theoretical_dmma_tflops_at_max_clock": 148 * 64 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12,
                  "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "clocks": clocks,
                  "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}))
