#!/usr/bin/env python
"""Benchmark of the GP hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg3|cfg2|cfg4|cfg5] [--batch B] [--scaling weak|strong]
                    [--points-total M]

One "step" = one pass of the hot path over one batch of synthetic input (BASELINE.json configs):
  cfg3 (default, the config the metric is quoted on): nlZ + gradient for B hyperparameter rows,
        Matern-5 ARD + NegativeQuadratic + constant noise, N=5000, D=10.  Weak scaling: B=64 rows per
        GPU; --scaling strong: --batch rows in total, cut into contiguous blocks over the GPUs.
  cfg2: same for SquaredExponential ARD + ConstantMean, N=2000, D=6, B=1024.
  cfg4: ONE RationalQuadratic-ARD GP, N=32768, D=8: nlZ + gradient (a single factorisation does not
        shard: replicas only at N > 1); the line also carries the nlZ-only time and a cuSOLVER
        cross-check of nlZ.
  cfg5: predict, MaternIsotropic(3), D=10, N=2000, the SAME 256 posterior samples on every GPU, the
        test points sharded: 65536 points per GPU and step, generated on the device;
        --points-total 16777216 sets the number of steps so that all GPUs together sweep 2^24 points.
Multi-GPU steps go through the product's sharding layer (gpyreg_b200/sharding.py): the kernels write
into the send buffer of one NCCL all-gather.  Prints ONE JSON line (rank 0).
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

PREDICT_POINTS_PER_STEP = 65536


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
        return dict(key="cfg3", name="cfg3: nlZ+grad, Matern5-ARD + NegativeQuadratic + const noise, N=5000, D=10",
                    spec=ModelSpec(D=10, cov_kind=1, degree=5, ard=True, mean_kind=2), N=5000, B=64,
                    kind="nlz")
    if name == "cfg2":
        return dict(key="cfg2", name="cfg2: nlZ+grad, SE-ARD + ConstantMean + const noise, N=2000, D=6",
                    spec=ModelSpec(D=6, cov_kind=0, ard=True, mean_kind=1), N=2000, B=1024, kind="nlz")
    if name == "cfg4":
        return dict(key="cfg4", name="cfg4: one large GP, nlZ+grad, RationalQuadratic-ARD + ConstantMean + const noise, "
                                     "N=32768, D=8",
                    spec=ModelSpec(D=8, cov_kind=2, ard=True, mean_kind=1), N=32768, B=1, kind="nlz")
    if name == "cfg5":
        return dict(key="cfg5", name="cfg5: predict, MaternIso(3) + ConstantMean + const noise, N=2000, D=10, "
                                     "256 hyperparameter samples replicated, test points sharded",
                    spec=ModelSpec(D=10, cov_kind=1, degree=3, ard=False, mean_kind=1), N=2000, B=256,
                    M=PREDICT_POINTS_PER_STEP, kind="predict")
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
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            try:
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ----------------------------------------------------------------------------- CPU arm
_REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_reference():
    """The UNMODIFIED reference staged by __graft_entry__.build() in baseline/_ref (travels to the
    GPU box), behind SURVEY.md Appendix B's matplotlib stub; None when it is not there."""
    if not os.path.isdir(os.path.join(_REF_DIR, "gpyreg")):
        return None
    import types
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    try:
        import gpyreg
        return gpyreg
    except Exception:
        return None


def reference_gp(gpyreg, spec):
    """gpyreg.GP for a ModelSpec, built from the reference's own plugin classes."""
    cf, icf = gpyreg.covariance_functions, gpyreg.isotropic_covariance_functions
    if spec.cov_kind == 0:
        cov = cf.SquaredExponential() if spec.ard else icf.SquaredExponentialIsotropic()
    elif spec.cov_kind == 1:
        cov = cf.Matern(spec.degree) if spec.ard else icf.MaternIsotropic(spec.degree)
    else:
        cov = cf.RationalQuadraticARD()
    mean = (gpyreg.mean_functions.ZeroMean, gpyreg.mean_functions.ConstantMean,
            gpyreg.mean_functions.NegativeQuadratic)[spec.mean_kind]()
    p = spec.noise_params
    noise = gpyreg.noise_functions.GaussianNoise(p[0] == 1, p[1] >= 1, p[1] == 2, p[2] == 1)
    return gpyreg.GP(spec.D, cov, mean, noise)


class CpuArm:
    """The reference's CPU implementation of the path on the host cores: the reference package itself
    (kind "reference") when baseline/_ref is there, else the oracle port (kind "port").  Every call
    is a BOUNDED SAMPLE of the workload, scaled to the metric's unit."""

    def __init__(self, wl):
        self.wl, self.spec = wl, wl["spec"]
        self.N_cpu = min(wl["N"], 8192)       # cfg4: the CPU arm runs N=8192 and scales by (N/8192)^3
        self.X, self.y = synth_data(self.N_cpu, self.spec.D, 0)
        self.ref = load_reference()
        self.kind = "reference" if self.ref is not None else "port"
        if self.ref is not None:
            self.gp = reference_gp(self.ref, self.spec)
            self.gp.X, self.gp.y = self.X, self.y

    def seconds(self, hyp_row):
        """-> (seconds, units of the metric that work corresponds to)."""
        wl, spec = self.wl, self.spec
        t0 = time.perf_counter()
        if wl["kind"] == "nlz":
            if self.ref is not None:
                self.gp._GP__compute_nlZ(hyp_row, True, False)
            else:
                from oracle import gp_oracle as orc
                orc.core(spec, hyp_row, self.X, self.y, None, True, True)
            units = (self.N_cpu / wl["N"]) ** 3
        else:
            Xs = np.random.default_rng(2).uniform(-3, 3, (2000, spec.D))
            if self.ref is not None:
                self.gp.update(hyp=hyp_row[None, :])
                self.gp.predict(Xs)
            else:
                from oracle import gp_oracle as orc
                orc.predict(spec, orc.posterior_batch(spec, hyp_row[None, :], self.X, self.y, None), self.X, self.y, Xs)
            units = 2000.0 / wl["B"]          # test points for the full sample count, linear in Ns
        return time.perf_counter() - t0, units

    def sample_text(self, n):
        wl = self.wl
        if wl["kind"] == "predict":
            return "%d x (1 of %d samples, posterior + 2000 test points), scaled linearly in samples" % (n, wl["B"])
        s = "%d hyperparameter row(s) of the same workload" % n
        if self.N_cpu != wl["N"]:
            s += " at N=%d, scaled by (N/%d)^3" % (self.N_cpu, self.N_cpu)
        return s


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
    arm = CpuArm(wl)
    hyp = benign_hyp(wl["spec"], max(8, args.steps + args.warmup), arm.y, 1)
    unit = "evals/s" if wl["kind"] == "nlz" else "test points/s"
    for i in range(args.warmup):
        arm.seconds(hyp[i])
    t, units = 0.0, 0.0
    for i in range(args.steps):
        dt, u = arm.seconds(hyp[args.warmup + i])
        t += dt
        units += u
    val = units / t
    line = {
        "impl": "reference", "metric": metric_name(wl), "value": val, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "N": wl["N"], "D": wl["spec"].D, "batch_per_gpu": wl["B"]},
        "cpu_baseline": {"value": val, "unit": unit, "cores": blas_threads(), "kind": arm.kind,
                         "sample": arm.sample_text(1) + " per step"},
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
    (MEASURED_PEAKS.json has no FP64 entry; this is the roofline denominator.  A kept copy with the
    4 s sustained figure and clocks: profiles/r02_fp64_peak.json, tools/fp64_peak.py.)"""
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


def measured_traffic(wl, B):
    """DRAM bytes per step of the dominant kernel (all gemm_nt_kernel launches of one step), from
    the committed per-launch ncu table (tools/ncu_traffic.py -> profiles/r02_dram_bytes.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_dram_bytes.json")) as f:
            t = json.load(f).get(wl["key"])
    except (OSError, ValueError):
        return None, None
    if not t or t.get("batch") != B:
        return None, None
    return t["gemm_dram_bytes_per_step"], t


def cusolver_crosscheck(torch, eng, spec, X, y, h, dev):
    """Independent nlZ for one hyperparameter row: K from the plugin kernel, float64 Cholesky by
    cuSOLVER through torch (cfg4: the CPU oracle cannot run N=32768 inside a bench)."""
    N = X.shape[0]
    K = eng.cov(spec.cov_kind, spec.degree, spec.ard, h[:spec.cov_n], X)
    Kt = torch.from_numpy(K).to(dev)
    del K
    Kt.diagonal().add_(float(np.exp(2 * h[spec.cov_n])))
    m0 = h[spec.cov_n + spec.noise_n] if spec.mean_kind >= 1 else 0.0
    r = torch.from_numpy(y[:, 0] - m0).to(dev)
    L = torch.linalg.cholesky(Kt)
    z = torch.linalg.solve_triangular(L, r[:, None], upper=False)[:, 0]
    return float(0.5 * (z @ z) + torch.log(L.diagonal()).sum() + 0.5 * N * np.log(2 * np.pi))


def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from gpyreg_b200 import Engine
    from gpyreg_b200.sharding import (nlz_block_to_send_buffer, predict_block_to_send_buffer, shard_bounds,
                                      sharded_nlz_device)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gpyreg_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec, N = wl["spec"], wl["N"]
    strong = args.scaling == "strong"
    replicas = wl["key"] == "cfg4"                 # a single factorisation never spans GPUs
    if wl["kind"] == "predict":
        strong = False
    X, y = synth_data(N, spec.D, 0)
    eng = Engine(local)
    eng.set_model(spec.cov_kind, spec.degree, spec.ard, spec.mean_kind, spec.noise_params)
    eng.set_data(X, y, None)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    P = spec.hyp_n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    extra = {}
    if wl["kind"] == "nlz":
        # the global batch: weak scaling = B rows per GPU, strong scaling = B rows in all
        B_cfg = args.batch or wl["B"]
        B_global = B_cfg if strong else B_cfg * world
        hyp_all = benign_hyp(spec, B_global, y, 1)
        if replicas:
            lo, hi, per = 0, B_cfg, B_cfg
            hyp = hyp_all[:B_cfg]
            B_global = B_cfg * world
        else:
            per = -(-B_global // world)
            lo, hi = shard_bounds(B_global, world)[rank]
            hyp = np.ascontiguousarray(hyp_all[lo:hi])
        n_local = hi - lo
        d_hyp = torch.from_numpy(hyp).to(dev) if n_local else None
        width = 3 + P
        recv = torch.empty(world * per * width, dtype=torch.float64, device=dev) if world > 1 and not replicas else None

        def step_dev():
            # the product's multi-GPU step: kernels write into the all-gather's send buffer
            send = nlz_block_to_send_buffer(eng, d_hyp, n_local, P, True, per, dev)
            if recv is not None:
                dist.all_gather_into_tensor(recv, send)
            return send

        def step_e2e():
            if world > 1 and not replicas:
                return sharded_nlz_device(eng, hyp_all, True)     # upload block, gather, one D2H
            return eng.nlz_batch(hyp, want_grad=True)

        units_per_step = B_global
        h2d, d2h = hyp.nbytes, 8 * n_local * (P + 2) + 4 * n_local
        unit = "evals/s"
        B_report = B_cfg
    else:
        M = wl["M"]
        Ns = args.batch or wl["B"]
        hyp = benign_hyp(spec, Ns, y, 1)           # the SAME samples on every rank
        post = eng.posterior_batch(hyp)
        steps_total = args.steps
        gen = torch.Generator(device=dev)

        def points(step):
            """test-point block (step * world + rank) of the global sweep, generated on the device"""
            gen.manual_seed(1000003 * (step * world + rank) + 2)
            return torch.rand((M, spec.D), dtype=torch.float64, device=dev, generator=gen) * 6.0 - 3.0

        recv = torch.empty(world * 2 * M, dtype=torch.float64, device=dev) if world > 1 else None
        state = {"step": 0}
        d_pts = points(0)

        def step_dev():
            send = predict_block_to_send_buffer(eng, post, d_pts, M, 1, False, False, M, dev)
            if recv is not None:
                dist.all_gather_into_tensor(recv, send)
            state["step"] += 1
            return send

        Xs_host = np.random.default_rng(2 + rank).uniform(-3, 3, (M, spec.D))

        def step_e2e():
            return eng.predict(post, Xs_host)

        units_per_step = M * world
        h2d, d2h = Xs_host.nbytes, 16 * M
        unit = "test points/s"
        B_report = Ns
        n_local = Ns

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
    for i in range(args.steps):
        if wl["kind"] == "predict":
            d_pts = points(i)                      # a new block of the sweep every step (device RNG, in the timed region)
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
    value = units_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the public host-buffer API (H2D of inputs + D2H of results per step)
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
    e2e_val = units_per_step * e2e_steps / float(dt.item())

    # ---- the same step through the GP object (gpyreg.GP API: host overhead of the Python layer)
    e2e_gp = None
    if wl["kind"] == "nlz" and world == 1 and wl["key"] != "cfg4":
        from gpyreg_b200.gaussian_process import gp_from_spec
        gp = gp_from_spec(spec)
        gp.X, gp.y = X, y
        gp._nlz_batch(hyp, True, False)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            gp._nlz_batch(hyp, True, False)
        e2e_gp = units_per_step * e2e_steps / (time.perf_counter() - t0)

    if wl["key"] == "cfg4" and rank == 0:
        # nlZ-only latency and the cuSOLVER cross-check of nlZ
        warm = hyp.copy()
        warm[0, 0] += 1e-3
        eng.nlz_batch(warm, want_grad=False)
        t0 = time.perf_counter()
        nlz_only = eng.nlz_batch(hyp, want_grad=False)[0]
        extra["nlz_only_ms"] = 1e3 * (time.perf_counter() - t0)
        extra["nlz_only_tflops"] = N ** 3 / 3 / (extra["nlz_only_ms"] * 1e-3) / 1e12
        extra["nlZ"] = float(nlz_only[0])
        try:
            ref = cusolver_crosscheck(torch, eng, spec, X, y, hyp[0], dev)
            extra["nlZ_cusolver"] = ref
            extra["nlZ_rel_diff_vs_cusolver"] = abs(ref - extra["nlZ"]) / abs(ref)
        except Exception as e:                   # out of memory next to the workspace: report, do not fail
            extra["nlZ_cusolver_error"] = str(e)[:200]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (the FP64 tile GEMM, gemm_nt_kernel<*>)
    peak = fp64_peak_tflops(torch, dev)
    Np = -(-N // 128) * 128
    if wl["kind"] == "nlz":
        flops_step = float(n_local) * float(N) ** 3        # potrf N^3/3 + inverse 2N^3/3, this GPU
        gemm_ms = (phase.get("factor", 0) + phase.get("inverse", 0)) / args.steps
    else:
        flops_step = float(wl["M"]) * B_report * float(N) ** 2    # triangular solve, N^2 per point and sample
        gemm_ms = ms / args.steps
    achieved = flops_step / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    traffic, ttab = measured_traffic(wl, n_local)
    roofline = {"bound": "tensor", "kernel": "gemm_nt_kernel (FP64 DMMA tile GEMM), all launches of one step on one GPU",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "traffic_note": ("DRAM read+write bytes of the gemm_nt_kernel launches of one step, ncu "
                                 "dram__bytes_read.sum + dram__bytes_write.sum per launch, summed "
                                 "(profiles/r02_dram_bytes.json: %s)" % ttab.get("source", "")) if ttab else
                                "no committed ncu table for this workload / batch",
                "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul, best of 5, measured in this run "
                               "(MEASURED_PEAKS.json has no FP64 entry; kept copy: profiles/r02_fp64_peak.json)",
                "algorithmic_flops_per_step": flops_step,
                "phase_ms_per_step": {k: v / args.steps for k, v in phase.items()},
                "padded_N": Np}
    # ---- CPU baseline: the reference on this box's host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        arm = CpuArm(wl)
        chyp = benign_hyp(spec, 8, arm.y, 1)
        reps = 1 if arm.N_cpu >= 4000 else 4
        if arm.N_cpu < 4000:
            arm.seconds(chyp[7])                          # warm the BLAS threads
        cpu_t = cpu_units = 0.0
        for i in range(reps):
            dtc, u = arm.seconds(chyp[i])
            cpu_t += dtc
            cpu_units += u
        cpu = {"value": cpu_units / cpu_t, "unit": unit, "cores": blas_threads(), "kind": arm.kind,
               "sample": arm.sample_text(reps) + ", %.1f s" % cpu_t}
    if wl["kind"] == "predict":
        line_extra = {"samples": B_report, "test_points_per_step_per_gpu": wl["M"],
                      "test_points_total": wl["M"] * world * args.steps,
                      "parallelism": f"test points x{world}, all {B_report} posterior samples on every GPU"}
    elif replicas:
        line_extra = {"parallelism": f"replicas x{world} (one factorisation per GPU)"}
    else:
        line_extra = {"parallelism": f"hyp-batch x{world}" + (" (strong: global batch fixed)" if strong else "")}
    line = {
        "metric": metric_name(wl), "value": value, "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl["name"], "N": N, "D": spec.D, "P": P,
                   "batch_per_gpu": n_local if wl["kind"] == "nlz" else B_report,
                   "global_batch": units_per_step if wl["kind"] == "nlz" else B_report, **line_extra,
                   "l2": "working set (%d x %.0f MB of matrices) far larger than the 126 MB L2"
                         % (n_local, 2 * Np * Np * 8 / 1e6)},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if e2e_gp is not None:
        line["e2e_gp_api"] = {"value": e2e_gp, "unit": unit,
                              "what": "GP._nlz_batch (the gpyreg.GP-level batched entry point) on host arrays"}
    if extra:
        line["extra"] = extra
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
    ap.add_argument("--batch", type=int, default=0,
                    help="hyperparameter rows per GPU per step (weak) / in total (strong); cfg5: posterior samples")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--points-total", type=int, default=0,
                    help="cfg5: number of test points all GPUs sweep together (sets --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = workload(args.workload)
    if wl["kind"] == "predict" and args.points_total:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        args.steps = max(1, -(-args.points_total // (wl["M"] * world)))
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
