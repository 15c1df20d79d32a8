#!/usr/bin/env python
"""Benchmark of the GP hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg3|cfg2|cfg5] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic input:
  cfg3 (default, the config BASELINE.json's metric is quoted on): nlZ + gradient for B
        hyperparameter rows, Matern-5 ARD + NegativeQuadratic + constant noise, N=5000, D=10.
  cfg2: same for SquaredExponential ARD + ConstantMean, N=2000, D=6.
  cfg5: predict, MaternIsotropic(3), D=10, N=2000, Ns samples, M test points per step.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------- workloads
def synth_data(N, D, seed=0):
    """X ~ U(-3,3), y = sin(sum x) - 0.125 |x|^2 + 0.1 eps."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    y = (np.sin(X.sum(1)) - 0.125 * (X ** 2).sum(1) + 0.1 * rng.standard_normal(N)).reshape(-1, 1)
    return X, y


def benign_hyp(spec, B, y, seed=1):
    """SURVEY.md 8(d) 'benign' hyperparameter rows (never trigger the jitter retry)."""
    rng = np.random.default_rng(seed)
    D = spec.D
    cols = [np.log(1.5) + 0.3 * rng.standard_normal((B, D if spec.ard else 1)),
            0.3 * rng.standard_normal((B, 1))]
    if spec.cov_kind == 2:
        cols.append(0.3 * rng.standard_normal((B, 1)))
    p = spec.noise_params
    if p[0] == 1:
        cols.append(np.log(0.1) + 0.3 * rng.standard_normal((B, 1)))
    if p[1] == 2:
        cols.append(0.2 * rng.standard_normal((B, 1)))
    if p[2] == 1:
        cols.append(np.median(y) + 0.3 * rng.standard_normal((B, 1)))
        cols.append(np.log(0.05) + 0.2 * rng.standard_normal((B, 1)))
    if spec.mean_kind >= 1:
        cols.append(float(np.mean(y)) + 0.1 * rng.standard_normal((B, 1)))
    if spec.mean_kind == 2:
        cols.append(0.3 * rng.standard_normal((B, D)))
        cols.append(np.log(3) + 0.2 * rng.standard_normal((B, D)))
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def workload(name):
    from gpyreg_b200.spec import ModelSpec
    if name == "cfg3":
        return dict(name="cfg3: nlZ+grad, Matern5-ARD + NegativeQuadratic + const noise, N=5000, D=10",
                    spec=ModelSpec(D=10, cov_kind=1, degree=5, ard=True, mean_kind=2), N=5000, B=64,
                    kind="nlz")
    if name == "cfg2":
        return dict(name="cfg2: nlZ+grad, SE-ARD + ConstantMean + const noise, N=2000, D=6",
                    spec=ModelSpec(D=6, cov_kind=0, ard=True, mean_kind=1), N=2000, B=1024, kind="nlz")
    if name == "cfg5":
        return dict(name="cfg5: predict, MaternIso(3) + ConstantMean + const noise, N=2000, D=10",
                    spec=ModelSpec(D=10, cov_kind=1, degree=3, ard=False, mean_kind=1), N=2000, B=32,
                    M=65536, kind="predict")
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Sample nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_eval_seconds(wl, hyp, X, y, reps):
    """Time the CPU oracle (NumPy/SciPy port of the reference path) on `reps` rows."""
    from oracle import gp_oracle as orc
    spec = wl["spec"]
    t0 = time.perf_counter()
    if wl["kind"] == "nlz":
        for b in range(reps):
            orc.core(spec, hyp[b % len(hyp)], X, y, None, True, True)
        units = reps
    else:
        posts = orc.posterior_batch(spec, hyp[:reps], X, y, None)
        Xs = np.random.default_rng(2).uniform(-3, 3, (2000, spec.D))
        orc.predict(spec, posts, X, y, Xs)
        units = 2000 * reps / wl["B"]     # test points for the full sample count, linear in Ns
    return time.perf_counter() - t0, units


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max((p.get("num_threads", 1) for p in threadpool_info()), default=os.cpu_count())
    except Exception:
        return os.cpu_count()


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is allowed every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    X, y = synth_data(wl["N"], wl["spec"].D, 0)
    hyp = benign_hyp(wl["spec"], max(8, args.steps + args.warmup), y, 1)
    unit = "evals/s" if wl["kind"] == "nlz" else "test points/s"
    # one step = ONE hyperparameter row (nlz) / 1 posterior sample x 2000 points (predict):
    # a bounded sample of the workload; evaluations are independent (f_min_fill.py:174-176)
    for i in range(args.warmup):
        cpu_eval_seconds(wl, hyp[i:i + 1], X, y, 1)
    t, units = 0.0, 0.0
    for i in range(args.steps):
        dt, u = cpu_eval_seconds(wl, hyp[args.warmup + i:args.warmup + i + 1], X, y, 1)
        t += dt
        units += u
    val = units / t
    sample = ("1 hyperparameter row per step (of B=%d), N=%d" % (wl["B"], wl["N"])) if wl["kind"] == "nlz" \
        else "1 of %d samples x 2000 test points per step, scaled linearly" % wl["B"]
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": val, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "N": wl["N"], "D": wl["spec"].D, "batch_per_gpu": wl["B"]},
        "cpu_baseline": {"value": val, "unit": unit, "cores": blas_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def metric_name(wl):
    if wl["kind"] == "predict":
        return "predict pts/s (N=%d, D=%d, FP64)" % (wl["N"], wl["spec"].D)
    return "nlZ+grad evals/s @N=%d,D=%d FP64" % (wl["N"], wl["spec"].D)


# ----------------------------------------------------------------------------- GPU arm
def fp64_peak_tflops(torch, dev, n=8192, reps=5):
    """FP64 tensor peak measured in-run: cuBLAS DGEMM n^3 through torch.matmul, best of reps.
    (MEASURED_PEAKS.json has no FP64 entry; this is the roofline denominator.)"""
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from gpyreg_b200 import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gpyreg_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL prints its version banner on stdout when NCCL_DEBUG is set in the environment:
    # divert it so stdout carries only the one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec, N, B = wl["spec"], wl["N"], (args.batch or wl["B"])
    X, y = synth_data(N, spec.D, 0)
    eng = Engine(local)
    eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    eng.set_data(X, y, None)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    P = spec.hyp_n
    # weak scaling: every rank owns its own B rows of one global (world*B, P) batch
    hyp_all = benign_hyp(spec, B * world, y, 1)
    hyp = np.ascontiguousarray(hyp_all[rank * B:(rank + 1) * B])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if wl["kind"] == "nlz":
        d_hyp = torch.from_numpy(hyp).to(dev)
        d_out = torch.empty((B, P + 1), dtype=torch.float64, device=dev)     # [nlZ | dnlZ]
        d_nlz = torch.empty(B, dtype=torch.float64, device=dev)
        d_dnlz = torch.empty((B, P), dtype=torch.float64, device=dev)
        gathered = torch.empty((world * B, P + 1), dtype=torch.float64, device=dev) if world > 1 else None

        def step_dev():
            eng.nlz_batch_dev(d_hyp.data_ptr(), B, True, d_nlz.data_ptr(), d_dnlz.data_ptr())
            if world > 1:       # the path's only exchange: all-gather of (nlZ, dnlZ) over NVLink
                d_out[:, 0] = d_nlz
                d_out[:, 1:] = d_dnlz
                dist.all_gather_into_tensor(gathered, d_out)

        def step_e2e():
            return eng.nlz_batch(hyp, want_grad=True)

        units_per_step = B
        h2d, d2h = hyp.nbytes, 8 * B * (P + 2) + 4 * B
        unit = "evals/s"
    else:
        M = wl["M"]
        post = eng.posterior_batch(hyp)
        Xs = np.random.default_rng(2 + rank).uniform(-3, 3, (M, spec.D))
        d_Xs = torch.from_numpy(Xs).to(dev)
        d_mu = torch.empty(M, dtype=torch.float64, device=dev)
        d_s2 = torch.empty(M, dtype=torch.float64, device=dev)
        gathered = torch.empty((world, 2, M), dtype=torch.float64, device=dev) if world > 1 else None
        d_pack = torch.empty((2, M), dtype=torch.float64, device=dev)

        def step_dev():
            eng.predict_dev(post, d_Xs.data_ptr(), M, False, False, d_mu.data_ptr(), d_s2.data_ptr())
            if world > 1:
                d_pack[0] = d_mu
                d_pack[1] = d_s2
                dist.all_gather_into_tensor(gathered.view(world * 2, M), d_pack)

        def step_e2e():
            return eng.predict(post, Xs)

        units_per_step = M
        h2d, d2h = Xs.nbytes, 16 * M
        unit = "test points/s"

    # ---- warm-up, then the timed region (device-resident inputs)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {}
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_dev()
        if wl["kind"] == "nlz":
            for k, v in eng.last_timings().items():
                phase[k] = phase.get(k, 0.0) + v
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = units_per_step * world * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C ABI (H2D of inputs + D2H of results per step)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_val = units_per_step * world * e2e_steps / float(dt.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (the FP64 tile GEMM, gemm_nt_kernel<*>)
    peak = fp64_peak_tflops(torch, dev)
    Np = -(-N // 128) * 128
    if wl["kind"] == "nlz":
        flops_step = float(B) * float(N) ** 3              # potrf N^3/3 + inverse 2N^3/3
        gemm_ms = (phase.get("factor", 0) + phase.get("inverse", 0)) / args.steps
    else:
        flops_step = float(wl["M"]) * B * float(N) ** 2    # triangular solve, N^2 per point and sample
        gemm_ms = ms / args.steps
    achieved = flops_step / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    roofline = {"bound": "tensor", "kernel": "gemm_nt_kernel (FP64 DMMA tile GEMM)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None,
                "traffic_note": "tensor-bound, not measured live; ncu capture of the largest trailing-update launch "
                                "(666 tiles x 64 matrices): 11.9 GB DRAM read+write vs 11.2 GB algorithmic C-tile "
                                "bytes, operands hit in L2 (profiles/r01_ncu_full_gemm_bench_default.txt)",
                "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul, best of 5, measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_step": flops_step,
                "phase_ms_per_step": {k: v / args.steps for k, v in phase.items()},
                "padded_N": Np}
    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample
    cpu = None
    if world == 1:
        use_all_host_threads()
        reps = 1 if N >= 4000 else 4
        cpu_eval_seconds(wl, hyp, X, y, 1) if N < 4000 else None      # warm the BLAS threads
        cpu_t, cpu_units = cpu_eval_seconds(wl, hyp, X, y, reps)
        cpu = {"value": cpu_units / cpu_t, "unit": unit, "cores": blas_threads(), "kind": "port",
               "sample": f"{reps} hyperparameter row(s) of the same workload, {cpu_t:.1f} s"}
    if wl["kind"] == "predict":
        line_extra = {"samples": B, "test_points_per_step_per_gpu": wl["M"]}
    else:
        line_extra = {}
    line = {
        "metric": metric_name(wl), "value": value, "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl["name"], "N": N, "D": spec.D, "P": P, "batch_per_gpu": B,
                   "global_batch": B * world, "parallelism": f"hyp-batch x{world}", **line_extra,
                   "l2": "working set (B x %.0f MB of matrices) far larger than the 126 MB L2"
                         % (2 * Np * Np * 8 / 1e6)},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--batch", type=int, default=0, help="hyperparameter rows per GPU per step")
    args = ap.parse_args()
    wl = workload(args.workload)
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
