"""Multi-GPU realignment: one process per GPU (torchrun), reads sharded by cost, no data-path collective for
realignment, one 212 x int64 all-reduce per EM iteration.

Rank 0 runs the host pipeline (SAM in / SAM out, the Target functions of nanopore_b200.realign) with a
`ShardedRealigner`; the other ranks sit in `worker_loop()` and serve its calls:

    set_reference   broadcast of the packed reference            (NCCL broadcast over NVLink)
    set_hmm         broadcast of 25 + 80 model probabilities      (NCCL broadcast)
    set_batch       rank 0 balances the reads on estimated DP cells (longest first) and sends every rank ONLY its
                    shard (point to point); the shard stays resident on the rank for later calls
    realign         every rank realigns its shard; CIGAR ops (uint32, and posterior pairs on request) go back to
                    rank 0 point to point, unpadded, and are put back in input order (the reference re-joins by
                    position, reference nanopore/analyses/utils.py:597)
    expectations    every rank runs the E-step on its resident shard; the exact-integer statistics are summed with
                    one all-reduce -- integer addition is associative, so 1, 2, 4 and 8 GPUs train the same HMM
                    bit for bit (the reference sums per-job expectation files in double)
    base_expectations  every rank scatter-adds the posterior pairs of its shard into per-reference-position tables
                    on its GPU (one table per sample of reads); the int64 tables are summed with one all-reduce
                    (marginAlignSnpCaller.py:149-155 builds them by parsing one text file per read)

The reference has no collectives at all: its only parallelism is jobTree's one-job-per-read task farm
(utils.py:565-570).  Backend: NCCL when CUDA is present, gloo otherwise (CPU tests of this logic).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .batch import Batch, estimate_cells
from .engine import FixedStats, Realigner
from .hmm import Hmm

CMD_EXIT, CMD_SET_REF, CMD_SET_HMM, CMD_SET_BATCH, CMD_REALIGN, CMD_EXPECT, CMD_BASE_EXPECT = range(7)


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
    import heapq
    cost = np.asarray(cost, dtype=np.int64)
    order = np.argsort(-cost, kind="stable")
    heap = [(0, r) for r in range(world)]
    parts = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(int(i))
        heapq.heappush(heap, (load + int(cost[i]), r))
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def read_cost(batch, params=None):
    """Estimated DP cells of every read (batch.estimate_cells): what a rank's time is proportional to.  Anti-diagonals
    alone are the wrong proxy when band widths differ between reads (mixed lengths, sparse anchors)."""
    if params is None:
        return estimate_cells(batch)
    return estimate_cells(batch, params.band, params.anchor_trim, params.split_side)


# ---- wire format: a list of numpy arrays <-> one uint8 blob (sizes and dtypes in an int64 header) ----
_DT = [np.dtype(np.uint8), np.dtype(np.int64), np.dtype(np.uint32), np.dtype(np.float64), np.dtype(np.int32)]


def _pack(arrays):
    arrays = [np.ascontiguousarray(a) for a in arrays]
    head = np.array([len(arrays)] + [v for a in arrays for v in (_DT.index(a.dtype), a.size)], dtype=np.int64)
    parts = [np.array([head.size], dtype=np.int64).view(np.uint8), head.view(np.uint8)]
    for a in arrays:
        raw = a.view(np.uint8).reshape(-1)
        parts.append(raw)
        if raw.size % 8:
            parts.append(np.zeros(8 - raw.size % 8, dtype=np.uint8))      # keep every array 8-byte aligned
    return np.concatenate(parts)


def _unpack(blob):
    hs = int(blob[:8].view(np.int64)[0])
    head = blob[8:8 + 8 * hs].view(np.int64)
    out, o = [], 8 + 8 * hs
    for k in range(int(head[0])):
        dt, n = _DT[int(head[1 + 2 * k])], int(head[2 + 2 * k])
        nb = n * dt.itemsize
        out.append(blob[o:o + nb].view(dt).copy())
        o += nb + (-nb) % 8
    return out


_staging = {}                      # destination rank -> reusable host staging buffer (pinned under NCCL)


def _stage(dst, nbytes):
    """Host staging buffer for the shard of rank dst: pinned memory when the transport is NCCL (the H2D copy is a DMA
    straight out of it and does not block the host), grown on demand, reused from step to step."""
    t = _staging.get(dst)
    if t is None or t.numel() < nbytes:
        cap = int(nbytes * 1.25) + 4096
        t = torch.empty(cap, dtype=torch.uint8, pin_memory=(dist.get_backend() == "nccl"))
        _staging[dst] = t
    return t[:nbytes]


def _pack_shard(batch, idx, dst):
    """The wire blob of batch.subset(idx) gathered straight into rank dst's staging buffer: no intermediate sub-batch,
    one pass over the shard's bases and guide ops (native ragged gather).  -> uint8 tensor (a view of the staging buffer)."""
    from .batch import _gather_ranges
    idx = np.asarray(idx, dtype=np.int64)
    n = len(idx)
    rl = batch.read_off[idx + 1] - batch.read_off[idx] if n else np.zeros(0, np.int64)
    ol = batch.in_off[idx + 1] - batch.in_off[idx] if n else np.zeros(0, np.int64)
    read_off = np.concatenate(([0], np.cumsum(rl))).astype(np.int64)
    in_off = np.concatenate(([0], np.cumsum(ol))).astype(np.int64)
    sizes = [(0, int(read_off[-1])), (1, n + 1), (1, n), (1, n), (2, int(in_off[-1])), (1, n + 1)]      # _BATCH_FIELDS order
    head = np.array([len(sizes)] + [v for dt, sz in sizes for v in (dt, sz)], dtype=np.int64)
    pos, o = [], 8 + 8 * head.size
    for dt, sz in sizes:
        nb = sz * _DT[dt].itemsize
        pos.append((o, nb))
        o += nb + (-nb) % 8
    t = _stage(dst, o)
    buf = t.numpy()
    buf[:8] = np.array([head.size], dtype=np.int64).view(np.uint8)
    buf[8:8 + 8 * head.size] = head.view(np.uint8)
    view = lambda k, dt: buf[pos[k][0]:pos[k][0] + pos[k][1]].view(dt)
    _gather_ranges(batch.reads, batch.read_off, idx, out=view(0, np.uint8))
    view(1, np.int64)[:] = read_off
    view(2, np.int64)[:] = batch.ref_start[idx]
    view(3, np.int64)[:] = batch.ref_end[idx]
    _gather_ranges(batch.in_ops, batch.in_off, idx, out=view(4, np.uint32))
    view(5, np.int64)[:] = in_off
    return t


def _send_blob(blob, dst, wait=True):
    """blob: numpy uint8 array or uint8 tensor (e.g. a pinned staging view).  wait=False returns the pending work and the
    tensors it needs alive; the caller waits before the staging buffer is reused."""
    dev = _device()
    t = blob if isinstance(blob, torch.Tensor) else torch.from_numpy(blob)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
    dist.send(n, dst)
    if t.numel() == 0:
        return None
    td = t.to(dev, non_blocking=True)
    if wait:
        dist.send(td, dst)
        return None
    return dist.isend(td, dst), td, t


def _recv_blob(src):
    dev = _device()
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.recv(n, src)
    t = torch.empty(int(n[0]), dtype=torch.uint8, device=dev)
    if int(n[0]):
        dist.recv(t, src)
    return t.cpu().numpy()


def _bcast_arrays(arrays, src=0):
    """Broadcast of a list of numpy arrays (known only on src) as one blob."""
    dev = _device()
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    if dist.get_rank() == src:
        blob = _pack(arrays)
        n[0] = blob.size
    dist.broadcast(n, src)
    t = torch.from_numpy(blob).to(dev) if dist.get_rank() == src else torch.empty(int(n[0]), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return _unpack(t.cpu().numpy())


def _params_to_array(p):
    return np.array([p.band, p.anchor_trim, p.split_side, p.min_diags, p.tb_diags, p.threshold, p.gap_gamma, p.match_gamma],
                    dtype=np.float64)


def _params_from_array(a):
    return capi.default_params(band=int(a[0]), anchor_trim=int(a[1]), split_side=int(a[2]), min_diags=int(a[3]),
                               tb_diags=int(a[4]), threshold=float(a[5]), gap_gamma=float(a[6]), match_gamma=float(a[7]))


_BATCH_FIELDS = ("reads", "read_off", "ref_start", "ref_end", "in_ops", "in_off")


class _RankLocal:
    """What every rank does with its shard.  The shard stays on the rank until rank 0 sends another one, so an EM run
    (hundreds of E-steps over the same alignments, utils.py:509-523) moves the reads once."""

    def __init__(self, local_factory):
        self.local = local_factory()
        self.ref = None
        self.sub = None

    def set_reference(self, codes):
        self.ref = codes
        self.sub = None
        self.local.set_reference(codes)

    def set_hmm(self, arr):
        if arr.size == 0:
            self.local.set_hmm(None)
        else:
            h = Hmm(int(arr[0]))
            h.transitions, h.emissions = arr[1:26].tolist(), arr[26:106].tolist()
            self.local.set_hmm(h)

    def set_batch(self, arrs):
        self.sub = Batch(self.ref, *arrs)

    def realign(self, params, want_posteriors):
        """-> [ops, off, cells, (posterior off, ref_pos, read_pos, prob_1e7)] of this rank's shard."""
        sub = self.sub
        if sub.n:
            ops, off, post = self.local.realign(sub, params, want_posteriors=want_posteriors)
        else:
            ops, off = np.zeros(0, np.uint32), np.zeros(1, np.int64)
            post = {"off": np.zeros(1, np.int64), "ref_pos": np.zeros(0, np.int32), "read_pos": np.zeros(0, np.int32),
                    "prob_1e7": np.zeros(0, np.int32)} if want_posteriors else None
        out = [np.asarray(ops, dtype=np.uint32), np.asarray(off, dtype=np.int64),
               np.array([getattr(self.local, "cells", 0) if sub.n else 0], dtype=np.int64)]
        if want_posteriors:
            out += [np.asarray(post["off"], np.int64), np.asarray(post["ref_pos"], np.int32),
                    np.asarray(post["read_pos"], np.int32), np.asarray(post["prob_1e7"], np.int32)]
        return out

    def base_expectations(self, params, masks):
        """masks: uint8[n_samples, reads of this shard] -> int64[n_samples, reference length, 5], summed over ranks."""
        if self.sub.n:
            t = self.local.base_expectations(self.sub, params, masks=list(masks))
        else:
            t = np.zeros((len(masks), len(self.ref), 5), dtype=np.int64)
        t = torch.from_numpy(np.ascontiguousarray(t)).to(_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                    # int64: exact, independent of the sharding
        return t.cpu().numpy()

    def expectations(self, params):
        st = self.local.expectations(self.sub, params) if self.sub.n else FixedStats()
        t = torch.from_numpy(st.as_tensor_array()).to(_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                    # 212 x int64: exact, order independent
        return FixedStats.from_tensor_array(t.cpu().numpy())


def _default_local_factory():
    return Realigner(device=torch.cuda.current_device())


def _concat_ranges(per_rank_data, per_rank_off, g_idx, n, dtype):
    """Re-joins per-rank ragged results in input order: read g_idx[r][k] owns data_r[off_r[k]:off_r[k+1]].  The rows of all
    ranks are laid end to end and gathered through the inverse of the sharding permutation (one native ragged gather)."""
    from .batch import _gather_ranges
    order = np.concatenate([np.asarray(i, dtype=np.int64) for i in g_idx]) if g_idx else np.zeros(0, np.int64)
    assert len(order) == n
    cat = np.concatenate([np.asarray(d, dtype=dtype) for d in per_rank_data]) if per_rank_data else np.zeros(0, dtype)
    lens = np.concatenate([np.asarray(o[1:] - o[:-1], dtype=np.int64) for o in per_rank_off]) if per_rank_off else np.zeros(0, np.int64)
    cat_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    inv = np.empty(n, dtype=np.int64)
    inv[order] = np.arange(n, dtype=np.int64)
    return _gather_ranges(cat, cat_off, inv)


class ShardedRealigner:
    """Rank 0's handle; same interface as engine.Realigner."""

    def __init__(self, hmm=None, local_factory=None):
        assert dist.get_rank() == 0, "ShardedRealigner lives on rank 0; other ranks run worker_loop()"
        self._rl = _RankLocal(local_factory or _default_local_factory)
        self.cells = 0
        self.rank_cells = []
        self._sent = None              # (batch, cost parameters, shards) of the batch the ranks hold
        self.set_hmm(hmm)

    def _cmd(self, c):
        t = torch.tensor([c], dtype=torch.int64, device=_device())
        dist.broadcast(t, 0)

    def set_reference(self, codes):
        self._cmd(CMD_SET_REF)
        self._sent = None
        self._rl.set_reference(_bcast_arrays([np.ascontiguousarray(codes, dtype=np.uint8)])[0])

    def set_hmm(self, hmm):
        self._cmd(CMD_SET_HMM)
        if hmm is None:
            arr = np.zeros(0, dtype=np.float64)
        else:
            t, e = hmm.arrays()
            arr = np.concatenate(([float(hmm.type)], t, e))
        self._rl.set_hmm(_bcast_arrays([arr])[0])

    def _ensure_batch(self, batch, params):
        """Scatters the batch: every rank receives ONLY its shard (cost-balanced on estimated DP cells), once."""
        key = (params.band, params.anchor_trim, params.split_side)
        if self._sent is not None and self._sent[0] is batch and self._sent[1] == key:
            return self._sent[2]
        self._cmd(CMD_SET_BATCH)
        shards = shard_reads(read_cost(batch, params), dist.get_world_size())
        pending = []
        for r in range(1, dist.get_world_size()):
            # packed straight into rank r's (pinned) staging buffer and sent without blocking: the transfer of shard r
            # overlaps the packing of shard r + 1
            pending.append(_send_blob(_pack_shard(batch, shards[r], r), r, wait=False))
        sub0 = batch.subset(shards[0])
        self._rl.set_batch([getattr(sub0, f) for f in _BATCH_FIELDS])
        for w in pending:
            if w is not None:
                w[0].wait()
        if pending and dist.get_backend() == "nccl":
            torch.cuda.current_stream().synchronize()      # the staging buffers may be refilled from here on
        self._sent = (batch, key, shards)
        return shards

    def realign(self, batch, params, want_posteriors=False):
        shards = self._ensure_batch(batch, params)
        self._cmd(CMD_REALIGN)
        _bcast_arrays([_params_to_array(params), np.array([1 if want_posteriors else 0], dtype=np.int64)])
        res = [self._rl.realign(params, want_posteriors)]
        for r in range(1, dist.get_world_size()):
            res.append(_unpack(_recv_blob(r)))                      # only rank 0 receives; nothing is padded or widened
        self.rank_cells = [int(x[2][0]) for x in res]            # DP cells per rank: the balance of the shards
        self.cells = int(sum(self.rank_cells))
        ops, off = _concat_ranges([x[0] for x in res], [x[1] for x in res], shards, batch.n, np.uint32)
        post = None
        if want_posteriors:
            post = {}
            for k, name in ((4, "ref_pos"), (5, "read_pos"), (6, "prob_1e7")):
                post[name], post["off"] = _concat_ranges([x[k] for x in res], [x[3] for x in res], shards, batch.n, np.int32)
        return ops, off, post

    def base_expectations(self, batch, params, masks=None):
        shards = self._ensure_batch(batch, params)
        ms = np.ones((1, batch.n), dtype=np.uint8) if masks is None else np.stack([np.asarray(m, dtype=np.uint8) for m in masks])
        self._cmd(CMD_BASE_EXPECT)
        _bcast_arrays([_params_to_array(params)])
        for r in range(1, dist.get_world_size()):
            _send_blob(_pack([np.ascontiguousarray(ms[:, shards[r]]).reshape(-1), np.array([ms.shape[0]], dtype=np.int64)]), r)
        out = self._rl.base_expectations(params, np.ascontiguousarray(ms[:, shards[0]]))
        self.cells = 0
        return out

    def expectations(self, batch, params):
        self._ensure_batch(batch, params)
        self._cmd(CMD_EXPECT)
        _bcast_arrays([_params_to_array(params)])
        return self._rl.expectations(params)

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
            rl.set_reference(_bcast_arrays(None)[0])
        elif c == CMD_SET_HMM:
            rl.set_hmm(_bcast_arrays(None)[0])
        elif c == CMD_SET_BATCH:
            rl.set_batch(_unpack(_recv_blob(0)))
        elif c == CMD_REALIGN:
            pa, wp = _bcast_arrays(None)
            _send_blob(_pack(rl.realign(_params_from_array(pa), bool(wp[0]))), 0)
        elif c == CMD_EXPECT:
            rl.expectations(_params_from_array(_bcast_arrays(None)[0]))
        elif c == CMD_BASE_EXPECT:
            pa = _params_from_array(_bcast_arrays(None)[0])
            m, ns = _unpack(_recv_blob(0))
            rl.base_expectations(pa, m.reshape(int(ns[0]), -1))
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
