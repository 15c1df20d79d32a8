"""Multi-GPU use of the path: one process per GPU (torch.distributed), the hyperparameter
batch or the test points cut into contiguous blocks, every rank evaluates its block on its
own B200, and ONE all-gather returns the full result to every rank.  There
is no exchange inside the computation, so nothing else is communicated: X, y, s2 are
replicated (a few MB), a single factorisation never spans GPUs.

Two layers:
  * ``sharded_nlz_device`` / ``sharded_predict_device`` -- the product path under NCCL.  The
    rank's block goes to the device once, the C ABI's ``*_dev`` entry points write their results
    straight into the send buffer of ONE ``all_gather_into_tensor``, and the gathered result
    crosses to the host once (the GP API returns NumPy): no NumPy round trip between the
    kernels and the collective.  The engine is duck-typed (``nlz_batch_dev`` / ``predict_dev``
    taking raw pointers), so the same code runs under gloo on CPU tensors with a stand-in
    engine (tests/test_sharding_gloo.py).
  * ``sharded_nlz`` / ``sharded_rows`` -- generic versions with an injected NumPy evaluator, for
    calls the ``*_dev`` entry points do not cover (log predictive density, per-point noise).
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


# --------------------------------------------------------------------------------------
# device-resident path (NCCL): results go from the kernels into the collective's send buffer
# --------------------------------------------------------------------------------------
def _settle(dev):
    """The engine works on its own CUDA stream: buffers torch has just filled on ITS stream must be
    complete before the engine writes into them (the engine synchronises its stream before it
    returns, so the other direction is already ordered)."""
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()


def _world_rank(group):
    if dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def nlz_block_to_send_buffer(engine, d_rows, n, P, want_grad, per, dev):
    """Evaluate the ``n`` hyperparameter rows in ``d_rows`` (device tensor (n, P)) and return the
    send buffer of the all-gather, laid out as struct-of-arrays over ``per`` slots:
    [ nlZ (per) | sn2_mult (per) | status as int32 in the first half of (per) doubles | dnlZ (per*P) ].
    The C ABI writes each output directly into its region."""
    width = 3 + (P if want_grad else 0)
    send = torch.zeros(per * width, dtype=torch.float64, device=dev)
    if n > 0:
        _settle(dev)
        base, w = send.data_ptr(), 8 * per
        engine.nlz_batch_dev(d_rows.data_ptr(), n, want_grad, base, (base + 3 * w) if want_grad else 0,
                             base + w, base + 2 * w)
    return send


def unpack_nlz(recv, B, P, want_grad, world, per):
    """Gathered (world, per*width) buffer -> (nlz, dnlz|None, sn2_mult, status) for the B rows."""
    width = 3 + (P if want_grad else 0)
    host = recv.reshape(world, per * width).cpu()
    nlz, mult, status = np.empty(B), np.empty(B), np.empty(B, dtype=np.int32)
    dnlz = np.empty((B, P)) if want_grad else None
    for r, (lo, hi) in enumerate(shard_bounds(B, world)):
        n = hi - lo
        if n == 0:
            continue
        row = host[r]
        nlz[lo:hi] = row[:n].numpy()
        mult[lo:hi] = row[per:per + n].numpy()
        status[lo:hi] = row[2 * per:3 * per].view(torch.int32)[:n].numpy()
        if want_grad:
            dnlz[lo:hi] = row[3 * per:3 * per + n * P].reshape(n, P).numpy()
    return nlz, dnlz, mult, status


def sharded_nlz_device(engine, hyp, want_grad, group=None):
    """nlZ [, gradient] of every row of hyp (B, P) host array, identical on all ranks: rank r
    uploads block r, ``engine.nlz_batch_dev`` fills the send buffer, one all-gather, one D2H.
    Returns (nlz, dnlz|None, sn2_mult, status) for the whole batch on every rank."""
    hyp = np.ascontiguousarray(np.atleast_2d(hyp), dtype=np.float64)
    B, P = hyp.shape
    world, rank = _world_rank(group)
    per = -(-B // world)
    lo, hi = shard_bounds(B, world)[rank]
    dev = _device(group)
    d_rows = torch.from_numpy(hyp[lo:hi]).to(dev) if hi > lo else None
    send = nlz_block_to_send_buffer(engine, d_rows, hi - lo, P, want_grad, per, dev)
    if world > 1:
        recv = torch.empty(world * send.numel(), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        recv = send
    return unpack_nlz(recv, B, P, want_grad, world, per)


def predict_block_to_send_buffer(engine, post, d_pts, n, cols, add_noise, separate, per, dev):
    """[ mu (per*cols) | s2 (per*cols) ] for the rank's ``n`` test points (device tensor (n, D))."""
    send = torch.zeros(2 * per * cols, dtype=torch.float64, device=dev)
    if n > 0:
        _settle(dev)
        base = send.data_ptr()
        engine.predict_dev(post, d_pts.data_ptr(), n, add_noise, separate, base, base + 8 * per * cols)
    return send


def sharded_predict_device(engine, post, Xs, add_noise=False, separate=False, group=None):
    """Predictive mean / variance at the rows of Xs (M, D), test points sharded over the ranks and
    every posterior sample replicated (the across-sample average stays local): block upload,
    ``engine.predict_dev`` into the send buffer, one all-gather, one D2H.  -> (mu, s2), (M, cols)."""
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    M = Xs.shape[0]
    cols = post.count if separate else 1
    world, rank = _world_rank(group)
    per = -(-M // world)
    lo, hi = shard_bounds(M, world)[rank]
    dev = _device(group)
    d_pts = torch.from_numpy(Xs[lo:hi]).to(dev) if hi > lo else None
    send = predict_block_to_send_buffer(engine, post, d_pts, hi - lo, cols, add_noise, separate, per, dev)
    if world > 1:
        recv = torch.empty(world * send.numel(), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        recv = send
    host = recv.reshape(world, 2, per * cols).cpu().numpy()
    mu, s2 = np.empty((M, cols)), np.empty((M, cols))
    for r, (a, b) in enumerate(shard_bounds(M, world)):
        if b > a:
            mu[a:b] = host[r, 0, :(b - a) * cols].reshape(b - a, cols)
            s2[a:b] = host[r, 1, :(b - a) * cols].reshape(b - a, cols)
    return mu, s2
