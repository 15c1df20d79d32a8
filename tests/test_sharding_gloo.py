"""N > 1 host logic on CPU: world_size-2 gloo processes partition a hyperparameter batch /
a set of test points and all-gather the results (the evaluator is a NumPy stand-in: the
sharding layer never looks inside it)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from gpyreg_b200.sharding import (shard_bounds, sharded_nlz, sharded_nlz_device, sharded_predict,
                                  sharded_predict_device)


def fake_eval(rows, want_grad):
    nlz = rows.sum(1) ** 2
    return nlz, (2 * rows if want_grad else None), np.where(rows[:, 0] > 0, 10.0, 1.0), (rows[:, 1] > 2).astype(np.int32)


def fake_pred(pts):
    return pts.sum(1, keepdims=True), (pts ** 2).sum(1, keepdims=True)


def _view(ptr, n, dtype=np.float64):
    import ctypes
    ct = ctypes.c_double if dtype == np.float64 else ctypes.c_int32
    return np.ctypeslib.as_array((ct * n).from_address(ptr))


class PointerEngine:
    """Stand-in for Engine's device-pointer entry points on CPU tensors: reads and writes through
    the raw pointers exactly where the C ABI would (gpb_nlz_batch_dev / gpb_predict_dev)."""

    def __init__(self, P, D):
        self.P, self.D, self.calls = P, D, []

    def nlz_batch_dev(self, d_hyp, n, want_grad, d_nlz, d_dnlz, d_mult=0, d_status=0):
        self.calls.append(n)
        rows = _view(d_hyp, n * self.P).reshape(n, self.P)
        nlz, dnlz, mult, status = fake_eval(rows, want_grad)
        _view(d_nlz, n)[:] = nlz
        _view(d_mult, n)[:] = mult
        _view(d_status, n, np.int32)[:] = status
        if want_grad:
            _view(d_dnlz, n * self.P)[:] = dnlz.reshape(-1)

    def predict_dev(self, post, d_Xs, M, add_noise, separate, d_mu, d_s2):
        pts = _view(d_Xs, M * self.D).reshape(M, self.D)
        cols = post.count if separate else 1
        mu, s2 = fake_pred(pts)
        _view(d_mu, M * cols).reshape(M, cols)[:] = mu + np.arange(cols)
        _view(d_s2, M * cols).reshape(M, cols)[:] = s2 + (1.0 if add_noise else 0.0)


class FakePost:
    count = 3


def _worker(rank, world, port, B, M, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        hyp = rng.standard_normal((B, 5))
        Xs = rng.standard_normal((M, 3))
        calls = []

        def ev(rows, g):
            calls.append(rows.shape[0])
            return fake_eval(rows, g)
        nlz, dnlz, mult, status = sharded_nlz(ev, hyp, True)
        nlz0, none, _, _ = sharded_nlz(ev, hyp, False)
        mu, s2 = sharded_predict(fake_pred, Xs)
        ref = fake_eval(hyp, True)
        ok = (np.array_equal(nlz, ref[0]) and np.array_equal(dnlz, ref[1]) and np.array_equal(mult, ref[2])
              and np.array_equal(status, ref[3]) and none is None and np.array_equal(nlz0, ref[0])
              and np.array_equal(mu, fake_pred(Xs)[0]) and np.array_equal(s2, fake_pred(Xs)[1]))
        lo, hi = shard_bounds(B, world)[rank]
        ok = ok and all(c == hi - lo for c in calls)
        # device-resident path (the product path under NCCL), here on CPU tensors
        eng = PointerEngine(5, 3)
        for grad in (True, False):
            d = sharded_nlz_device(eng, hyp, grad)
            ok = ok and np.array_equal(d[0], ref[0]) and np.array_equal(d[2], ref[2]) and \
                np.array_equal(d[3], ref[3]) and d[3].dtype == np.int32 and \
                (np.array_equal(d[1], ref[1]) if grad else d[1] is None)
        ok = ok and all(c == hi - lo for c in eng.calls if hi > lo)
        for sep in (False, True):
            mu_d, s2_d = sharded_predict_device(eng, FakePost(), Xs, add_noise=True, separate=sep)
            cols = 3 if sep else 1
            ok = ok and np.array_equal(mu_d, fake_pred(Xs)[0] + np.arange(cols)) and \
                np.array_equal(s2_d, np.repeat(fake_pred(Xs)[1] + 1.0, cols, axis=1))
        q.put((rank, bool(ok), calls))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("B,M", [(7, 11), (2, 1), (64, 1000)])
def test_two_rank_gloo(B, M):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, M, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


def test_shard_bounds():
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    nlz, dnlz, mult, status = sharded_nlz(fake_eval, np.ones((3, 4)), True)     # no process group
    assert nlz.shape == (3,) and dnlz.shape == (3, 4)
