"""SPMD use of the GP API under torchrun (one process per GPU): every rank runs the same host logic
with the same seeds; batches of hyperparameter rows (the f_min_fill design, lock-step L-BFGS starts,
slice-sampling chains) and large test sets are sharded over the GPUs through the device-resident
path of gpyreg_b200/sharding.py and all-gathered over NCCL.  Checks that all ranks agree and that
the sharded results equal a single-GPU evaluation bit for bit; reports the fit time.
usage: torchrun --nproc-per-node N tools/multi_gpu_fit.py [N_train] [n_chains]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import gpyreg_b200 as g
from bench import benign_hyp, synth_data
from gpyreg_b200.covariance_functions import Matern
from gpyreg_b200.spec import ModelSpec

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n_chains = int(sys.argv[2]) if len(sys.argv) > 2 else max(world, 2)
D = 6
X, y = synth_data(N, D, 0)
spec = ModelSpec(D=D, cov_kind=1, degree=5, ard=True, mean_kind=2)


def new_gp():
    return g.GP(D, Matern(5), g.mean_functions.NegativeQuadratic(), g.noise_functions.GaussianNoise(constant_add=True))


gp = new_gp()
gp.X, gp.y = X, y
hyp = benign_hyp(spec, 37, y, 1)
nlz, dnlz = gp._nlz_batch(hyp, True, False)                        # sharded: 37 rows over the ranks
ref = gp.engine.nlz_batch(hyp, want_grad=True)                      # this rank alone
same_nlz = bool(np.array_equal(nlz, ref[0]) and np.array_equal(dnlz, ref[1]))
out = {"world": world, "N": N, "D": D, "sharded_nlz_equals_single_gpu": same_nlz}
fits = {}
for tag, opts in (("sequential_sampler", {"n_samples": 8, "init_N": 512}),
                  ("multi_chain", {"n_samples": 2 * n_chains, "init_N": 512, "n_chains": n_chains})):
    gp = new_gp()
    np.random.seed(0)
    dist.barrier()
    t0 = time.perf_counter()
    hs, opt, _ = gp.fit(X=X, y=y, options=opts)
    torch.cuda.synchronize()
    fits[tag] = {"fit_s": round(time.perf_counter() - t0, 2), "opt_nlZ": float(opt.fun), "samples": int(hs.shape[0]),
                 "hyp_checksum": float(hs.sum())}
out["fit"] = fits
Xs = np.random.default_rng(2).uniform(-3, 3, (4096 * world + 13, D))
mu, s2 = gp.predict(Xs)                                             # sharded over test points (device-resident gather)
mu1, s21 = gp.engine.predict(gp._post_batch, Xs)
out["sharded_predict_equals_single_gpu"] = bool(np.array_equal(mu, mu1) and np.array_equal(s2, s21))
digest = torch.tensor([float(nlz.sum()), fits["sequential_sampler"]["hyp_checksum"], fits["multi_chain"]["hyp_checksum"],
                       float(mu.sum()), fits["multi_chain"]["opt_nlZ"]], dtype=torch.float64, device="cuda")
all_d = [torch.empty_like(digest) for _ in range(world)]
dist.all_gather(all_d, digest)
out["ranks_agree"] = all(bool(torch.equal(all_d[0], d)) for d in all_d)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
