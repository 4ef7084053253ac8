"""CPU baseline of BASELINE.json configs[2..4]: the CPU port (oracle/, the stand-in for the
reference's cactus_realign, which cannot be built here) timed on a bounded sample of each config's reads, one read
per task like the reference's job farm (reference nanopore/analyses/utils.py:565-570), on all host cores and on the 4
workers of the reference's default (Makefile:1).  Prints one JSON line per config.  Test / measurement infrastructure
(lives under tests/: nothing outside tests/, smoke() and bench.py's CPU legs touches oracle/).
usage: python tests/tools/cpu_baseline_configs.py [reads_per_core=2]"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle                                               # noqa: E402
from nanopore_b200 import synth                             # noqa: E402

cores = os.cpu_count() or 1
per_core = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = oracle.Model()


def timed(fn, idx, threads):
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        cells = sum(ex.map(fn, idx))
    return time.perf_counter() - t0, cells


for name, n, read_len, ref_len, band, split, lengths, estep in (
        ("C3: 5 kb reads vs 4.6 Mb, band 100", per_core * cores, 5000, 4600000, 100, 3000, None, False),
        ("C4: E-step, 8 kb reads vs 50 kb, band 10, split 300", per_core * cores, 8000, 50000, 10, 300, None, True),
        ("C5: Pareto 500 bp - 50 kb reads vs 50 kb, band 50", 4 * per_core * cores, 0, 50000, 50, 3000, "pareto", False)):
    ln = synth.pareto_lengths(n, seed=5) if lengths else None
    b = synth.make_batch(n, read_len, ref_len, seed=43, lengths=ln)
    p = oracle.make_params(expansion=band, split_side=split)
    if estep:
        fn = lambda i: oracle.expectations_fixed(m, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), p)[2]
    else:
        fn = lambda i: oracle.realign(m, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), p)["cells"]
    idx = list(range(n))
    t_all, cells = timed(fn, idx, cores)
    sub = idx[: max(4, n // 8)]
    t_4, cells4 = timed(fn, sub, 4)
    print(json.dumps({"config": name, "kind": "port", "sample_reads": n, "cores": cores, "reads_per_s_all_cores": n / t_all,
                      "cells_per_s_all_cores": cells / t_all, "s_all_cores": t_all, "reads_per_s_4_workers": len(sub) / t_4,
                      "sample_reads_4_workers": len(sub), "read_bases": int(b.read_off[-1])}), flush=True)
