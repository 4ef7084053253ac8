"""Multi-GPU check of nanopore_b200.parallel on real devices (run under torchrun on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
Rank 0 realigns a batch and runs an E-step through ShardedRealigner (reads sharded over the N GPUs, NCCL
broadcast of reference / HMM / batch, integer all-reduce of the statistics) and compares with its own single-GPU
Realigner: CIGARs identical, statistics identical as integers."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanopore_b200 import capi, parallel, synth          # noqa: E402
from nanopore_b200.engine import Realigner               # noqa: E402
from nanopore_b200.hmm import Hmm                        # noqa: E402


def main():
    rank, world = parallel.init()
    if rank != 0:
        parallel.worker_loop()
        return
    import torch
    n = int(os.environ.get("CHECK_READS", "96"))
    lengths = synth.pareto_lengths(n, seed=9, lo=500, hi=8000)
    b = synth.make_batch(n, 0, 20000, seed=77, lengths=lengths)
    p = capi.default_params(band=50)
    hmm = Hmm.loadHmm(os.path.join(os.path.dirname(os.path.abspath(parallel.__file__)), "mappers", "blasr_hmm_0.txt"))
    one = Realigner(torch.cuda.current_device(), hmm=hmm)
    one.set_reference(b.ref)
    t0 = time.perf_counter(); ops1, off1, _ = one.realign(b, p); t1 = time.perf_counter() - t0
    pe = capi.default_params(band=10, split_side=300)
    st1 = one.expectations(b, pe)                      # warm-up (allocations)
    t0 = time.perf_counter(); st1 = one.expectations(b, pe); e1 = time.perf_counter() - t0
    one.close()
    sr = parallel.ShardedRealigner(hmm)
    sr.set_reference(b.ref)
    t0 = time.perf_counter(); opsN, offN, _ = sr.realign(b, p); tN = time.perf_counter() - t0
    stN = sr.expectations(b, pe)                       # warm-up
    t0 = time.perf_counter(); stN = sr.expectations(b, pe); eN = time.perf_counter() - t0
    t0 = time.perf_counter(); opsN, offN, _ = sr.realign(b, p); tN = time.perf_counter() - t0
    sr.close()
    parallel.shutdown()
    ok = bool(np.array_equal(ops1, opsN) and np.array_equal(off1, offN) and st1 == stN)
    shards = parallel.shard_reads(parallel.read_cost(b), world)
    print(json.dumps({"multi_gpu_check": "OK" if ok else "MISMATCH", "world": world, "reads": n, "cigar_ops": int(len(ops1)),
                      "cells": int(sr.cells), "shard_sizes": [int(len(s)) for s in shards],
                      "shard_cost": [int(parallel.read_cost(b)[s].sum()) for s in shards],
                      "loglik": float(stN.values()[105]), "t_single_s": round(t1, 3), "t_sharded_s": round(tN, 3),
                      "estep_single_s": round(e1, 3), "estep_sharded_s": round(eN, 3)}), flush=True)
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
