"""BASELINE.json configs[2], [3] and [4] through the product's sharded path on the N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/configs_multi_gpu.py [--scale 1.0] [--configs 3,4,5] [--out gpurun_out/configs_multi_gpu.json]

  config 3  100k reads x 5 kb vs a 4.6 Mb reference, band 100, sharded by read           -> realigned CIGARs on rank 0
  config 4  Baum-Welch EM, 5 iterations over 50k reads x 8 kb, statistics all-reduced     -> trained HMM on rank 0
  config 5  1M reads, Pareto lengths 500 bp - 50 kb, band 50, balanced on estimated cells -> realigned CIGARs on rank 0

Rank 0 owns the batch (as it owns the SAM file in the pipeline) and drives nanopore_b200.parallel.ShardedRealigner:
reference / HMM broadcast over NCCL, every rank receives only its cost-balanced shard, results return point to point,
EM statistics are one 212 x int64 all-reduce per iteration.  Times are wall clock on rank 0 around whole calls (they
include packing, the scatter, the kernels on every rank and the gather); the per-rank DP cells show the balance.  A
sample of each config is repeated on rank 0's own single-GPU Realigner and must give the same bits.  --scale shrinks the
read counts (CPU dry runs in tests/, smaller boxes).  Synthetic reads are generated before CUDA is touched, on a pool of
host processes."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanopore_b200 import capi, em, parallel, synth           # noqa: E402
from nanopore_b200.batch import Batch                          # noqa: E402


def _gen(job):
    n, read_len, ref, seed, lengths = job
    b = synth.make_batch(n, read_len, len(ref), seed=seed, lengths=lengths, ref=ref)
    return b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off


def generate(n, read_len, ref_len, seed, lengths=None, procs=1, chunk=2000):
    """n synthetic reads against one random reference, generated in chunks on a pool of processes."""
    ref = synth.random_reference(ref_len, np.random.default_rng(seed))
    jobs = [(min(chunk, n - a), read_len, ref, seed * 1000003 + a, None if lengths is None else lengths[a:a + chunk]) for a in range(0, n, chunk)]
    if procs > 1 and len(jobs) > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_gen, jobs)
    else:
        parts = [_gen(j) for j in jobs]
    cat_off = lambda k: np.concatenate([[0]] + [p[k][1:] + o for p, o in zip(parts, np.concatenate(([0], np.cumsum([p[k][-1] for p in parts])[:-1])))])
    return Batch(ref, np.concatenate([p[0] for p in parts]), cat_off(1), np.concatenate([p[2] for p in parts]),
                 np.concatenate([p[3] for p in parts]), np.concatenate([p[4] for p in parts]), cat_off(5))


def same_as_single(sr_result, batch, idx, params, single, what):
    """The sharded result of reads idx against rank 0's own single-GPU run of just those reads."""
    sub = batch.subset(idx)
    if what == "realign":
        ops, off, _ = single.realign(sub, params)
        sops, soff = sr_result
        return all(np.array_equal(ops[off[k]:off[k + 1]], sops[soff[i]:soff[i + 1]]) for k, i in enumerate(idx))
    raise ValueError(what)


def main(argv=None, local_factory=None, single_factory=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--len-scale", type=float, default=1.0, help="shrinks read lengths too (CPU dry runs only)")
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--out", default="gpurun_out/configs_multi_gpu.json")
    ap.add_argument("--procs", type=int, default=0, help="generator processes on rank 0 (0 = host threads - ranks)")
    ap.add_argument("--verify", type=int, default=64, help="reads of each config repeated on one GPU for the bit check")
    a = ap.parse_args(argv)
    want = [int(c) for c in a.configs.split(",") if c]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank_env = int(os.environ.get("RANK", "0"))
    batches, gen_s = {}, {}
    if rank_env == 0:                                          # before any CUDA / NCCL state exists in this process
        procs = a.procs or max(1, (os.cpu_count() or 1) - world)
        ls = a.len_scale
        shapes = {3: (int(100000 * a.scale), int(5000 * ls), int(max(4600000 * min(1.0, a.scale * 50), 20000)), None),
                  4: (int(50000 * a.scale), int(8000 * ls), 50000, None),
                  5: (int(1000000 * a.scale), 0, 50000, "pareto")}
        for c in want:
            n, rl, ref_len, kind = shapes[c]
            n = max(n, 8)
            t0 = time.perf_counter()
            lengths = synth.pareto_lengths(n, seed=5, lo=int(500 * ls), hi=int(50000 * ls)) if kind == "pareto" else None
            batches[c] = generate(n, rl, ref_len, seed=40 + c, lengths=lengths, procs=procs)
            gen_s[c] = time.perf_counter() - t0
    rank, world = parallel.init()
    if rank != 0:
        parallel.worker_loop(local_factory)
        return 0
    from nanopore_b200.engine import Realigner
    import torch
    single_factory = single_factory or (lambda: Realigner(torch.cuda.current_device()))
    lines = []
    for c in want:
        b = batches[c]
        bases = int(b.read_off[-1])
        if c in (3, 5):
            p = capi.default_params(band=100 if c == 3 else 50)
            sr = parallel.ShardedRealigner(None, local_factory)
            t0 = time.perf_counter(); sr.set_reference(b.ref); t_ref = time.perf_counter() - t0
            t0 = time.perf_counter(); ops, off, _ = sr.realign(b, p); t_first = time.perf_counter() - t0      # scatter + realign + gather
            rank_cells = list(sr.rank_cells)
            t0 = time.perf_counter(); ops2, off2, _ = sr.realign(b, p); t_again = time.perf_counter() - t0    # shards resident
            sr.close()
            idx = np.unique(np.linspace(0, b.n - 1, min(a.verify, b.n)).astype(np.int64))
            one = single_factory(); one.set_reference(b.ref)
            ok = same_as_single((ops, off), b, idx, p, one, "realign") and np.array_equal(ops, ops2) and np.array_equal(off, off2)
            one.close()
            lens = ops >> 2
            spans = bool(np.array_equal(np.add.reduceat(np.where((ops & 3) != 2, lens, 0).astype(np.int64), off[:-1]), b.read_off[1:] - b.read_off[:-1]))
            lines.append({"config": c, "workload": ("100k reads x 5 kb vs 4.6 Mb, band 100" if c == 3 else "1M Pareto reads 500 bp - 50 kb vs 50 kb, band 50") + (" x scale %g" % a.scale if a.scale != 1 else ""),
                          "gpus": world, "reads": int(b.n), "read_bases": bases, "ref_len": int(len(b.ref)), "cells": int(sum(rank_cells)),
                          "reads_per_s_first_call": b.n / t_first, "reads_per_s_resident": b.n / t_again, "s_first_call": t_first, "s_resident_call": t_again,
                          "s_broadcast_reference": t_ref, "rank_cells": rank_cells, "balance_max_over_mean": max(rank_cells) / (sum(rank_cells) / len(rank_cells)) if sum(rank_cells) else 1.0,
                          "cigars_span_reads": spans, "sample_equals_single_gpu": bool(ok), "sample_reads": int(len(idx)), "generate_s": gen_s[c]})
        else:
            p = em.parseRealignOptions("--diagonalExpansion=10 --splitMatrixBiggerThanThis=300")       # utils.py:511
            hmm = em._stock_start("fiveStateAsymmetric")
            sr = parallel.ShardedRealigner(hmm, local_factory)
            sr.set_reference(b.ref)
            t0 = time.perf_counter(); st0 = sr.expectations(b, p); t_first = time.perf_counter() - t0         # scatter + prepare + first E-step
            t0 = time.perf_counter()
            rl = em.expectationMaximisation(sr, b, hmm, p, 5, trainEmissions=True)
            t_em = time.perf_counter() - t0
            # exactness of the reduction: the same E-step on a sample, sharded vs one GPU
            idx = np.unique(np.linspace(0, b.n - 1, min(4 * a.verify, b.n)).astype(np.int64))
            sub = b.subset(idx)
            sr.set_hmm(em._stock_start("fiveStateAsymmetric"))
            st_sh = sr.expectations(sub, p)
            sr.close()
            one = single_factory(); one.set_reference(b.ref); one.set_hmm(em._stock_start("fiveStateAsymmetric"))
            st_1 = one.expectations(sub, p)
            one.close()
            lines.append({"config": c, "workload": "EM: 5 iterations over 50k reads x 8 kb vs 50 kb, band 10, split 300" + (" x scale %g" % a.scale if a.scale != 1 else ""),
                          "gpus": world, "reads": int(b.n), "read_bases": bases, "iterations": 5, "s_first_estep_incl_scatter": t_first, "s_em_5_iterations": t_em,
                          "read_iterations_per_s": 5 * b.n / t_em, "running_likelihoods": rl,
                          "monotone": all(y >= x - 1e-9 * abs(x) for x, y in zip(rl[1:], rl[2:])),
                          "first_loglik": float(st0.values()[105]), "sample_statistics_equal_single_gpu": bool(st_sh == st_1), "sample_reads": int(len(idx)),
                          "generate_s": gen_s[c]})
        print(json.dumps(lines[-1]), flush=True)
    parallel.shutdown()
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    bad = [ln["config"] for ln in lines if not (ln.get("sample_equals_single_gpu", True) and ln.get("sample_statistics_equal_single_gpu", True) and ln.get("cigars_span_reads", True))]
    return 1 if bad else 0


if __name__ == "__main__":
    rc = main()
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
    sys.exit(rc)
