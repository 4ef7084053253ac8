"""Multi-GPU realignment: one process per GPU (torchrun), reads sharded by cost, no data-path collective for
realignment, one 212 x int64 all-reduce per EM iteration.

Rank 0 runs the host pipeline (SAM in / SAM out, the Target functions of nanopore_b200.realign) with a
`ShardedRealigner`; the other ranks sit in `worker_loop()` and serve its calls:

    set_reference   broadcast of the packed reference            (NCCL broadcast over NVLink)
    set_hmm         broadcast of 25 + 80 model probabilities      (NCCL broadcast)
    realign         broadcast of the packed batch; every rank realigns its own shard; shards' CIGAR ops are
                    gathered on rank 0 and put back in input order (the reference re-joins by position,
                    reference nanopore/analyses/utils.py:597)
    expectations    every rank runs the E-step on its shard; the exact-integer statistics are summed with one
                    all-reduce -- integer addition is associative, so 1, 2, 4 and 8 GPUs train the same HMM bit
                    for bit (the reference sums per-job expectation files in double)

The reference has no collectives at all: its only parallelism is jobTree's one-job-per-read task farm
(utils.py:565-570).  Backend: NCCL when CUDA is present, gloo otherwise (CPU tests of this logic).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .batch import Batch
from .engine import FixedStats, Realigner
from .hmm import Hmm

CMD_EXIT, CMD_SET_REF, CMD_SET_HMM, CMD_REALIGN, CMD_EXPECT = range(5)


def init(backend=None):
    """Joins the process group described by the torchrun environment. Returns (rank, world)."""
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def _device():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def shard_reads(cost, world):
    """Longest-processing-time-first greedy: reads sorted by descending cost (stable) go to the least loaded
    rank (ties: lowest rank).  Returns one ascending index array per rank.  Deterministic."""
    cost = np.asarray(cost, dtype=np.int64)
    order = np.argsort(-cost, kind="stable")
    load = [0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        parts[r].append(int(i))
        load[r] += int(cost[i])
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def read_cost(batch):
    """Proxy for DP cells: anti-diagonals of the read's matrix (the band width is the same for all reads)."""
    return (batch.read_off[1:] - batch.read_off[:-1]) + (batch.ref_end - batch.ref_start)


def _bcast_array(a, src=0):
    """Broadcasts a numpy array (dtype and size known only on src)."""
    dev = _device()
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    codes = {np.dtype(np.uint8): 0, np.dtype(np.int64): 1, np.dtype(np.uint32): 2, np.dtype(np.float64): 3, np.dtype(np.int32): 4}
    back = {v: k for k, v in codes.items()}
    if dist.get_rank() == src:
        a = np.ascontiguousarray(a)
        meta[0], meta[1] = a.size, codes[a.dtype]
    dist.broadcast(meta, src)
    n, dt = int(meta[0]), back[int(meta[1])]
    if dist.get_rank() == src:
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
    else:
        t = torch.empty(n * dt.itemsize, dtype=torch.uint8, device=dev)
    if n:
        dist.broadcast(t, src)
    return t.cpu().numpy().view(dt).copy()


def _params_to_array(p):
    return np.array([p.band, p.anchor_trim, p.split_side, p.min_diags, p.tb_diags, p.threshold, p.gap_gamma, p.match_gamma],
                    dtype=np.float64)


def _params_from_array(a):
    return capi.default_params(band=int(a[0]), anchor_trim=int(a[1]), split_side=int(a[2]), min_diags=int(a[3]),
                               tb_diags=int(a[4]), threshold=float(a[5]), gap_gamma=float(a[6]), match_gamma=float(a[7]))


def _bcast_batch(batch):
    arrs = []
    for name in ("reads", "read_off", "ref_start", "ref_end", "in_ops", "in_off"):
        arrs.append(_bcast_array(getattr(batch, name) if batch is not None else None))
    return arrs


def _gather_to_root(a, dtype):
    """Variable-length gather of one array per rank on rank 0 (all_gather of sizes, then padded all_gather)."""
    dev = _device()
    world = dist.get_world_size()
    a = np.ascontiguousarray(a, dtype=dtype)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([a.size], dtype=torch.int64, device=dev))
    sizes = [int(s) for s in sizes]
    m = max(max(sizes), 1)
    buf = np.zeros(m, dtype=np.int64)
    buf[:a.size] = a.astype(np.int64)
    outs = [torch.empty(m, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(buf).to(dev))
    return [o.cpu().numpy()[:s].astype(dtype) for o, s in zip(outs, sizes)]


class _RankLocal:
    """What every rank does with its shard."""

    def __init__(self, local_factory):
        self.local = local_factory()
        self.ref = None

    def set_reference(self, codes):
        self.ref = codes
        self.local.set_reference(codes)

    def set_hmm(self, arr):
        if arr.size == 0:
            self.local.set_hmm(None)
        else:
            h = Hmm(int(arr[0]))
            h.transitions, h.emissions = arr[1:26].tolist(), arr[26:106].tolist()
            self.local.set_hmm(h)

    def shard(self, arrs):
        full = Batch(self.ref, *arrs)
        idx = shard_reads(read_cost(full), dist.get_world_size())[dist.get_rank()]
        return full, idx, full.subset(idx)

    def realign(self, arrs, params):
        full, idx, sub = self.shard(arrs)
        if sub.n:
            ops, off, _ = self.local.realign(sub, params)
        else:
            ops, off = np.zeros(0, np.uint32), np.zeros(1, np.int64)
        g_idx = _gather_to_root(idx, np.int64)
        g_ops = _gather_to_root(ops, np.uint32)
        g_off = _gather_to_root(off, np.int64)
        g_cells = _gather_to_root(np.array([getattr(self.local, "cells", 0)]), np.int64)
        return full.n, g_idx, g_ops, g_off, int(sum(int(c[0]) for c in g_cells))

    def expectations(self, arrs, params):
        _, _, sub = self.shard(arrs)
        st = self.local.expectations(sub, params) if sub.n else FixedStats()
        t = torch.from_numpy(st.as_tensor_array()).to(_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                    # 212 x int64: exact, order independent
        return FixedStats.from_tensor_array(t.cpu().numpy())


def _default_local_factory():
    return Realigner(device=torch.cuda.current_device())


class ShardedRealigner:
    """Rank 0's handle; same interface as engine.Realigner."""

    def __init__(self, hmm=None, local_factory=None):
        assert dist.get_rank() == 0, "ShardedRealigner lives on rank 0; other ranks run worker_loop()"
        self._rl = _RankLocal(local_factory or _default_local_factory)
        self.cells = 0
        self.set_hmm(hmm)

    def _cmd(self, c):
        t = torch.tensor([c], dtype=torch.int64, device=_device())
        dist.broadcast(t, 0)

    def set_reference(self, codes):
        self._cmd(CMD_SET_REF)
        self._rl.set_reference(_bcast_array(np.ascontiguousarray(codes, dtype=np.uint8)))

    def set_hmm(self, hmm):
        self._cmd(CMD_SET_HMM)
        if hmm is None:
            arr = np.zeros(0, dtype=np.float64)
        else:
            t, e = hmm.arrays()
            arr = np.concatenate(([float(hmm.type)], t, e))
        self._rl.set_hmm(_bcast_array(arr))

    def realign(self, batch, params, want_posteriors=False):
        if want_posteriors:
            raise NotImplementedError("posterior pairs are returned by the single-GPU Realigner only")
        self._cmd(CMD_REALIGN)
        p = _params_from_array(_bcast_array(_params_to_array(params)))
        n, g_idx, g_ops, g_off, self.cells = self._rl.realign(_bcast_batch(batch), p)
        per_read = [None] * n
        for idx, ops, off in zip(g_idx, g_ops, g_off):
            for k, i in enumerate(idx):
                per_read[int(i)] = ops[off[k]:off[k + 1]]
        off = np.concatenate(([0], np.cumsum([len(o) for o in per_read]))).astype(np.int64)
        ops = np.concatenate(per_read) if n else np.zeros(0, np.uint32)
        return ops.astype(np.uint32), off, None

    def expectations(self, batch, params):
        self._cmd(CMD_EXPECT)
        p = _params_from_array(_bcast_array(_params_to_array(params)))
        return self._rl.expectations(_bcast_batch(batch), p)

    def close(self):
        """Closes rank 0's local context; workers stay up for the next ShardedRealigner (shutdown() ends them)."""
        self._rl.local.close()


def shutdown():
    """Rank 0: releases the workers."""
    t = torch.tensor([CMD_EXIT], dtype=torch.int64, device=_device())
    dist.broadcast(t, 0)


def worker_loop(local_factory=None):
    """Ranks > 0: serve rank 0's calls until shutdown()."""
    rl = None
    factory = local_factory or _default_local_factory
    while True:
        t = torch.zeros(1, dtype=torch.int64, device=_device())
        dist.broadcast(t, 0)
        c = int(t[0])
        if c == CMD_EXIT:
            if rl is not None:
                rl.local.close()
            return
        if rl is None:
            rl = _RankLocal(factory)
        if c == CMD_SET_REF:
            rl.set_reference(_bcast_array(None))
        elif c == CMD_SET_HMM:
            rl.set_hmm(_bcast_array(None))
        elif c == CMD_REALIGN:
            p = _params_from_array(_bcast_array(None))
            rl.realign(_bcast_batch(None), p)
        elif c == CMD_EXPECT:
            p = _params_from_array(_bcast_array(None))
            rl.expectations(_bcast_batch(None), p)
        else:
            raise RuntimeError("unknown command %d" % c)


def run(main_fn, local_factory=None):
    """torchrun entry: rank 0 runs main_fn() with ShardedRealigner installed as the realigner of
    nanopore_b200.realign; other ranks serve.  With WORLD_SIZE unset it is a plain single-GPU run."""
    from . import realign as _realign
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return main_fn()
    rank, _ = init()
    if rank == 0:
        prev = _realign.setRealignerFactory(lambda hmm=None: ShardedRealigner(hmm, local_factory))
        try:
            return main_fn()
        finally:
            _realign.setRealignerFactory(prev)
            shutdown()
            dist.destroy_process_group()
    else:
        worker_loop(local_factory)
        dist.destroy_process_group()
        return None
