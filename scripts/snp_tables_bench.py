"""Per-reference-position base-expectation tables of the SNP caller at a realistic size: the
device path (k_base_expect: posterior pairs scatter-added where they lie, one table per coverage sample, only the
tables come back) against returning every pair and summing on the host, which is what a caller of
--outputAllPosteriorProbs does (reference nanopore/analyses/marginAlignSnpCaller.py:136-155).  Same integers either way.
usage: python scripts/snp_tables_bench.py [reads=2000] [read_len=5000] [ref_len=48000]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanopore_b200 import posteriors, synth                 # noqa: E402
from nanopore_b200.engine import Realigner                  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
R = int(sys.argv[3]) if len(sys.argv) > 3 else 48000
b = synth.make_batch(n, L, R, seed=12)
p = posteriors.posteriorParams()                            # band 10, split 100 (marginAlignSnpCaller.py:136)
rng = np.random.default_rng(3)
cov = [1.0] + [f for f in (0.6, 0.3, 0.15, 0.05) for _ in range(3)]       # 13 samples: all reads + 3 replicates of 4 coverages
masks = [(rng.random(n) < f).astype(np.uint8) for f in cov]
r = Realigner(0)
r.set_reference(b.ref)
r.base_expectations(b, p, masks=masks[:1])                  # warm-up
t0 = time.perf_counter(); tables = r.base_expectations(b, p, masks=masks); t_dev = time.perf_counter() - t0
t0 = time.perf_counter(); _, _, post = r.realign(b, p, want_posteriors=True); t_pairs = time.perf_counter() - t0
t0 = time.perf_counter()
host = np.zeros((len(masks), len(b.ref), 5), dtype=np.int64)
for k, m in enumerate(masks):
    for i in np.nonzero(m)[0]:
        s = slice(post["off"][i], post["off"][i + 1])
        np.add.at(host[k], (b.ref_start[i] + post["ref_pos"][s].astype(np.int64), np.minimum(b.read(i)[post["read_pos"][s]], 4)),
                  post["prob_1e7"][s].astype(np.int64))
t_host = time.perf_counter() - t0
r.close()
print(json.dumps({"workload": "%d reads x %d bp vs %d bp, band 10, split 100, 13 coverage samples" % (n, L, R), "pairs": int(post["off"][-1]),
                  "identical_tables": bool(np.array_equal(tables, host)), "s_device_tables_13_samples": t_dev, "reads_per_s_device": n / t_dev,
                  "s_return_pairs": t_pairs, "s_host_sum_13_samples": t_host, "reads_per_s_pairs_plus_host_sum": n / (t_pairs + t_host),
                  "reference_invocations_replaced": int(sum(int(m.sum()) for m in masks))}), flush=True)
