"""Multi-GPU use of the path: one process per GPU (torch.distributed), the hyperparameter
batch or the test points cut into contiguous blocks, every rank evaluates its block on its
own B200, and ONE all-gather returns the full result to every rank.  There
is no exchange inside the computation, so nothing else is communicated: X, y, s2 are
replicated (a few MB), a single factorisation never spans GPUs.

The evaluation callables are injected, so the partition / gather logic is testable on CPU
with the gloo backend (tests/test_sharding_gloo.py); on a GPU box the backend is NCCL and
the gathered tensors live on the device.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous blocks of ceil(n/world) rows: [(lo, hi)] per rank (empty blocks allowed)."""
    per = -(-n // world) if n else 0
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def _device(group=None):
    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_rows(local, n_total, group=None):
    """Gather row blocks produced under shard_bounds(n_total, world) into one (n_total, C)
    array on every rank.  Blocks are padded to the common block size so a single
    all_gather_into_tensor moves everything."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.ndim == 1:
        local = local[:, None]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = -(-n_total // world)
    cols = local.shape[1]
    dev = _device(group)
    send = torch.zeros((per, cols), dtype=torch.float64, device=dev)
    send[:local.shape[0]] = torch.from_numpy(local).to(dev)
    recv = torch.empty((world * per, cols), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    out = recv.cpu().numpy()
    keep = np.concatenate([np.arange(r * per, r * per + (hi - lo))
                           for r, (lo, hi) in enumerate(shard_bounds(n_total, world))])
    return out[keep.astype(np.int64)]


def sharded_nlz(evaluate, hyp, want_grad, group=None):
    """nlZ [, gradient] of every row of hyp (B, P); rank r evaluates block r through
    ``evaluate(rows, want_grad) -> (nlz, dnlz|None, sn2_mult, status)`` (Engine.nlz_batch).
    Returns the same 4-tuple for the whole batch on every rank."""
    hyp = np.atleast_2d(np.asarray(hyp, dtype=np.float64))
    B, P = hyp.shape
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(B, world)[rank]
    cols = 3 + (P if want_grad else 0)
    local = np.zeros((hi - lo, cols))
    if hi > lo:
        nlz, dnlz, mult, status = evaluate(hyp[lo:hi], want_grad)
        local[:, 0], local[:, 1], local[:, 2] = nlz, mult, status
        if want_grad:
            local[:, 3:] = dnlz
    full = all_gather_rows(local, B, group)
    return (full[:, 0], full[:, 3:] if want_grad else None, full[:, 1],
            full[:, 2].astype(np.int32))


def sharded_rows(evaluate, M, group=None):
    """Generic row-sharded evaluation: rank r calls ``evaluate(lo, hi)`` for its block of the M
    rows and gets back a tuple of (hi-lo, c_i) arrays; every rank receives the tuple of full
    (M, c_i) arrays.  Used for predictions over test points (all posterior samples replicated on
    every GPU, so the across-sample average stays local)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(M, world)[rank]
    widths, local = None, None
    if hi > lo:
        parts = [np.asarray(p, dtype=np.float64).reshape(hi - lo, -1) for p in evaluate(lo, hi)]
        widths = [p.shape[1] for p in parts]
        local = np.concatenate(parts, axis=1)
    if world > 1:                       # agree on the column layout (empty shards know nothing)
        t = torch.zeros(8, dtype=torch.int64, device=_device(group))
        if widths is not None:
            t[0] = len(widths)
            t[1:1 + len(widths)] = torch.tensor(widths, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        widths = [int(v) for v in t[1:1 + int(t[0])].tolist()]
    if widths is None:                  # M == 0 on a single rank: nothing was evaluated
        return ()
    if local is None:
        local = np.zeros((0, sum(widths)))
    full = all_gather_rows(local, M, group)
    out, c = [], 0
    for w in widths:
        out.append(full[:, c:c + w])
        c += w
    return tuple(out)


def sharded_predict(evaluate, Xs, group=None):
    """Predictive mean/variance at every row of Xs (M, D); rank r handles test-point block r
    through ``evaluate(points) -> (mu, s2)``.  Returns (mu, s2) for all M points."""
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    return sharded_rows(lambda lo, hi: evaluate(Xs[lo:hi]), Xs.shape[0], group)
