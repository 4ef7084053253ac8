"""One rank of the world_size-2 gloo test of nanopore_b200.parallel (launched by test_parallel_gloo.py).
The CPU checker stands in for the GPU library on every rank: this exercises sharding, the command protocol,
the gather back into input order and the integer all-reduce -- not the kernels."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from nanopore_b200 import capi, parallel, synth            # noqa: E402
from oracle_ctx import oracle_realigner_factory               # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, world = parallel.init("gloo")
    local_factory = lambda: oracle_realigner_factory()(None)
    if rank != 0:
        parallel.worker_loop(local_factory)
        return
    lengths = [150, 420, 90, 300, 260, 510, 33]
    b = synth.make_batch(len(lengths), 0, 1500, seed=31, lengths=lengths)
    p = capi.default_params(band=10, split_side=300)
    sr = parallel.ShardedRealigner(None, local_factory)
    sr.set_reference(b.ref)
    ops, off, _ = sr.realign(b, p)
    _, _, post = sr.realign(b, p, want_posteriors=True)       # the shards are resident: nothing is sent again
    st = sr.expectations(b, p)
    st_again = sr.expectations(b, p)                          # EM calls this hundreds of times on the same batch
    assert st_again == st
    masks = [np.ones(b.n, np.uint8), np.array([1, 0, 1, 0, 0, 1, 0], np.uint8)]
    tables = sr.base_expectations(b, p, masks=masks)          # per-rank scatter-add, int64 all-reduce
    from nanopore_b200.hmm import Hmm
    h = Hmm.loadHmm(os.path.join(HERE, "golden", "blasr_hmm_0.txt"))
    sr.set_hmm(h)
    ops2, off2, _ = sr.realign(b, p)
    sr.close()
    parallel.shutdown()
    shards = [s.tolist() for s in parallel.shard_reads(parallel.read_cost(b), world)]
    json.dump({"ops": ops.tolist(), "off": off.tolist(), "hi": st.hi.tolist(), "lo": st.lo.tolist(), "cells": sr.cells,
               "ops_trained": ops2.tolist(), "off_trained": off2.tolist(), "shards": shards,
               "post": {k: v.tolist() for k, v in post.items()}, "tables": tables.tolist()}, open(out_path, "w"))


if __name__ == "__main__":
    main()
