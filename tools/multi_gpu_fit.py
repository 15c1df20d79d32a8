"""SPMD use of the GP API under torchrun: every rank runs the same host logic (same seeds), batches
of hyperparameter rows and large test sets are sharded over the GPUs and all-gathered.  Checks that
all ranks agree and that the results equal a single-GPU evaluation bit for bit."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import gpyreg_b200 as g
from gpyreg_b200.covariance_functions import Matern
from bench import synth_data, benign_hyp
from gpyreg_b200.spec import ModelSpec

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
N, D = 1500, 6
X, y = synth_data(N, D, 0)
spec = ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=2)
gp = g.GP(D, Matern(5), g.mean_functions.NegativeQuadratic(), g.noise_functions.GaussianNoise(constant_add=True))
gp.X, gp.y = X, y
hyp = benign_hyp(spec, 37, y, 1)
nlz, dnlz = gp._nlz_batch(hyp, True, False)                        # sharded: 37 rows over the ranks
ref = gp.engine.nlz_batch(hyp, want_grad=True)                      # this rank alone
same_local = bool(np.array_equal(nlz, ref[0]) and np.array_equal(dnlz, ref[1]))
np.random.seed(0)
t0 = time.perf_counter()
hs, opt, _ = gp.fit(X=X, y=y, options={"n_samples": 8, "init_N": 256})
fit_s = time.perf_counter() - t0
Xs = np.random.default_rng(2).uniform(-3, 3, (4096 * world + 13, D))
mu, s2 = gp.predict(Xs)                                             # sharded over test points
mu1, s21 = gp.engine.predict(gp._post_batch, Xs)
same_pred = bool(np.array_equal(mu, mu1) and np.array_equal(s2, s21))
digest = torch.tensor([float(nlz.sum()), float(hs.sum()), float(mu.sum()), float(opt.fun)], dtype=torch.float64, device="cuda")
all_d = [torch.empty_like(digest) for _ in range(world)]
dist.all_gather(all_d, digest)
ranks_agree = all(bool(torch.equal(all_d[0], d)) for d in all_d)
if rank == 0:
    print(json.dumps({"world": world, "sharded_nlz_equals_single_gpu": same_local,
                      "sharded_predict_equals_single_gpu": same_pred, "ranks_agree": ranks_agree,
                      "fit_s": round(fit_s, 2), "opt_nlZ": float(opt.fun)}))
dist.destroy_process_group()
