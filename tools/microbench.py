"""GPU micro-measurements used to steer the kernels (run under gpurun)."""
import json
import sys
import time
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpyreg_b200 import Engine
from gpyreg_b200.spec import ModelSpec
from bench import synth_data, benign_hyp, fp64_peak_tflops

out = {}
dev = torch.device("cuda", 0)
out["gpu"] = torch.cuda.get_device_name(0)
out["fp64_cublas_tflops_8192"] = fp64_peak_tflops(torch, dev, 8192, 5)
out["fp64_cublas_tflops_4096"] = fp64_peak_tflops(torch, dev, 4096, 5)
for bn in ("128", "64"):
    os.environ["GPB_GEMM_BN"] = bn
    e2 = Engine(0)
    for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 2048), (8192, 8192, 512), (8192, 8192, 256),
                      (8192, 8192, 128), (2048, 2048, 2048), (1024, 1024, 8192)]:
        ms = e2.debug_gemm_bench(M, N, K, 5)
        out[f"bn{bn}_gemm_{M}x{N}x{K}_tflops"] = round(2.0 * M * N * K / (ms * 1e-3) / 1e12, 2)
    e2.close()
print(json.dumps(out, indent=1), flush=True)
os.environ["GPB_GEMM_BN"] = os.environ.get("BN", "64")
eng = Engine(0)

for name, spec, N, Bs in [
    ("cfg2", ModelSpec(D=6, cov_kind=0, ard=True, mean_kind=1), 2000, (1, 64, 256)),
    ("cfg3", ModelSpec(D=10, cov_kind=1, degree=5, ard=True, mean_kind=2), 5000, (1, 8, 32)),
]:
    X, y = synth_data(N, spec.D, 0)
    eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    eng.set_data(X, y, None)
    for B in Bs:
        hyp = benign_hyp(spec, B, y, 1)
        for grad in (False, True):
            eng.nlz_batch(hyp, want_grad=grad)
            t0 = time.perf_counter()
            r = eng.nlz_batch(hyp, want_grad=grad)
            dt = time.perf_counter() - t0
            tm = eng.last_timings()
            flops = B * (N ** 3) * (1.0 if grad else 1 / 3)
            print(json.dumps({"wl": name, "B": B, "grad": grad, "wall_s": dt, "evals_per_s": B / dt,
                              "tflops_alg": flops / dt / 1e12, "phases_ms": tm,
                              "nlz0": float(r[0][0]), "mult_max": float(r[2].max())}), flush=True)
