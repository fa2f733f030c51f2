"""Multi-GPU plumbing: where the GP+ hot path shards, and nothing else.

The reference fans restarts out over joblib/loky processes (optim/mll_scipy.py:287-293) and scores
candidate tables per fidelity slice (bayesian_optimizations/BO_GP_plus.py:183-194).  Those are the
two places the work partitions:

* restarts   -- independent L-BFGS-B runs handed out by a cross-rank WORK QUEUE (an atomic counter in the process
                group's key-value store): a rank that finishes early claims the next restart instead of idling behind
                a static ``i % world`` partition.  One collective at the end gathers ``(nll, theta, counters)`` per
                restart: (p + 7) doubles each.
* candidates -- contiguous chunks per rank; one collective gathers ``(score, index)`` per rank and the
                arg-max (first index on ties) is taken on every rank.

One process per GPU under ``torchrun`` (``torch.distributed``, NCCL over NVLink on the GPU box, gloo
in the CPU tests).  Without an initialised process group everything degrades to a single process
that spreads restart workers over all visible GPUs.  A single Cholesky is never split across GPUs
(replicas only).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
from scipy.optimize import OptimizeResult

from . import _engine


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world() -> Tuple[int, int]:
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def local_devices() -> List[int]:
    """GPU indices this process may place engines on."""
    env = os.environ.get("GPPLUS_DEVICES")
    if env:
        return [int(s) for s in env.split(",") if s.strip() != ""]
    if _dist() is not None:
        return [int(os.environ.get("LOCAL_RANK", "0"))]
    n = _engine.device_count()
    return list(range(n)) if n > 0 else [0]


def _comm_device() -> torch.device:
    d = _dist()
    if d is not None and d.get_backend() == "nccl":
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return torch.device("cpu")


def shard_indices(count: int) -> List[int]:
    """Static round-robin partition (kept for callers that need a fixed ownership, e.g. warm-up)."""
    rank, size = world()
    return list(range(rank, count, size))


_queue_serial = 0


class RestartQueue:
    """Work queue of ``count`` restart indices shared by every worker thread of every rank.

    ``claim()`` returns the next unclaimed index (or None when the queue is drained).  Across ranks the queue is one
    atomic counter in the process group's store (``Store.add``, ~0.1 ms per claim -- restarts take 10 ms to minutes);
    without a process group it is a locked local counter.  Every rank must construct its queues in the same order."""

    def __init__(self, count: int):
        global _queue_serial
        import threading
        self.count = int(count)
        self._lock = threading.Lock()
        self._next = 0
        self._store = None
        d = _dist()
        if d is not None and d.get_world_size() > 1:
            from torch.distributed import distributed_c10d as c10d
            self._store = c10d._get_default_store()
            _queue_serial += 1
            self._key = "gpplus_b200/restart_queue/%d" % _queue_serial

    def claim(self):
        if self._store is None:
            with self._lock:
                i = self._next
                self._next += 1
        else:
            with self._lock:  # one client connection per process: keep its requests serial
                i = int(self._store.add(self._key, 1)) - 1
        return i if i < self.count else None


def shard_range(count: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``count`` items owned by this rank (candidate tables)."""
    rank, size = world()
    per = (count + size - 1) // size
    lo = min(rank * per, count)
    return lo, min(lo + per, count)


def broadcast_theta_list(theta0_list: Sequence[np.ndarray]) -> List[np.ndarray]:
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return [np.asarray(t, dtype=np.float64) for t in theta0_list]
    dev = _comm_device()
    shape = torch.tensor([len(theta0_list), len(theta0_list[0]) if len(theta0_list) else 0], dtype=torch.int64,
                         device=dev)
    d.broadcast(shape, src=0)
    k, p = int(shape[0]), int(shape[1])
    buf = torch.zeros(k, p, dtype=torch.float64, device=dev)
    if d.get_rank() == 0:
        buf.copy_(torch.as_tensor(np.stack([np.asarray(t, dtype=np.float64) for t in theta0_list])))
    d.broadcast(buf, src=0)
    arr = buf.cpu().numpy()
    return [arr[i].copy() for i in range(k)]


# per-restart record: [owned, status_code, fun, nit, nfev, njev, success, theta...]; status_code -1 = NotPSD, -2 = NaN
_HEAD = 7


def _encode(res, p: int) -> np.ndarray:
    rec = np.zeros(_HEAD + p)
    rec[0] = 1.0
    if isinstance(res, Exception):
        rec[1] = -2.0 if isinstance(res, _engine.NanError) else -1.0
        rec[2] = np.inf
        return rec
    rec[1] = float(getattr(res, "status", 0))
    rec[2] = float(res.fun)
    rec[3] = float(getattr(res, "nit", 0))
    rec[4] = float(getattr(res, "nfev", 0))
    rec[5] = float(getattr(res, "njev", 0))
    rec[6] = 1.0 if getattr(res, "success", False) else 0.0
    rec[_HEAD:] = np.asarray(res.x, dtype=np.float64)
    return rec


def _decode(rec: np.ndarray):
    if rec[1] == -1.0 and not np.isfinite(rec[2]):
        return _engine.NotPSDError("Matrix not positive definite after repeatedly adding jitter up to 1e-06.")
    if rec[1] == -2.0 and not np.isfinite(rec[2]):
        return _engine.NanError("NaN in the covariance matrix")
    return OptimizeResult(x=rec[_HEAD:].copy(), fun=float(rec[2]), nit=int(rec[3]), nfev=int(rec[4]),
                          njev=int(rec[5]), status=int(rec[1]), success=bool(rec[6]),
                          message="gathered from another rank")


def gather_restarts(local: Dict[int, object], count: int, p: int) -> List[object]:
    """All restart results in restart order on every rank, whichever rank ran which restart.  Local results keep
    their full ``OptimizeResult``; results of other ranks are rebuilt from the gathered record."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return [local[i] for i in range(count)]
    rank, size = world()
    dev = _comm_device()
    mine = np.zeros((count, _HEAD + p))
    for i, res in local.items():
        mine[i] = _encode(res, p)
    mine_t = torch.as_tensor(mine).to(dev)
    parts = [torch.zeros_like(mine_t) for _ in range(size)]
    d.all_gather(parts, mine_t)
    out: List[object] = [None] * count
    for r in range(size):
        arr = parts[r].cpu().numpy()
        for i in np.flatnonzero(arr[:, 0] == 1.0):
            if out[i] is None:
                out[i] = local[i] if r == rank else _decode(arr[i])
    missing = [i for i in range(count) if out[i] is None]
    if missing:
        raise RuntimeError("restarts %s were claimed by no rank" % missing)
    return out


def global_argmax(score: float, index: int) -> Tuple[float, int]:
    """Arg-max over ranks of per-rank (best score, global index); first index wins ties."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return score, index
    dev = _comm_device()
    mine = torch.tensor([score, float(index)], dtype=torch.float64, device=dev)
    parts = [torch.zeros_like(mine) for _ in range(d.get_world_size())]
    d.all_gather(parts, mine)
    best_s, best_i = -np.inf, -1
    for t in parts:
        s, i = float(t[0]), int(t[1])
        if i < 0:
            continue
        if best_i < 0 or s > best_s or (s == best_s and i < best_i):
            best_s, best_i = s, i
    return best_s, best_i
