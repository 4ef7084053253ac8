"""Multi-rank path on CPU: world_size 2 over gloo.  Reads are sharded by cost, realigned per rank and
re-joined in input order; EM statistics are all-reduced as exact integers, so two ranks give the bits of one."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

from nanopore_b200 import capi, parallel, synth
from nanopore_b200.engine import FixedStats
from nanopore_b200.hmm import Hmm

from oracle_ctx import oracle_realigner_factory

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_reads_is_balanced_and_deterministic():
    cost = np.array([100, 5, 80, 80, 7, 60, 1, 300])
    parts = parallel.shard_reads(cost, 3)
    assert sorted(np.concatenate(parts).tolist()) == list(range(8))
    loads = [int(cost[p].sum()) for p in parts]
    b = synth.make_batch(5, 0, 3000, seed=3, lengths=[200, 2000, 200, 900, 50])
    est = parallel.read_cost(b, capi.default_params(band=10))
    assert est[1] > est[3] > est[0] > est[4] > 0                       # the cost follows the DP cells, not the window
    assert max(loads) == 300 and sorted(loads)[0] >= 160                 # LPT: the long read alone, the rest split
    assert all(np.array_equal(a, b) for a, b in zip(parts, parallel.shard_reads(cost, 3)))
    assert all((np.diff(p) > 0).all() for p in parts if len(p) > 1)      # input order inside a shard
    assert [len(p) for p in parallel.shard_reads(np.array([5]), 4)] == [1, 0, 0, 0]


def test_world_size_2_matches_single_rank(tmp_path):
    out = str(tmp_path / "rank0.json")
    port = free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0, o
    got = json.load(open(out))
    assert all(len(s) > 0 for s in got["shards"])
    # single-process run of the same work
    lengths = [150, 420, 90, 300, 260, 510, 33]
    b = synth.make_batch(len(lengths), 0, 1500, seed=31, lengths=lengths)
    p = capi.default_params(band=10, split_side=300)
    r = oracle_realigner_factory()(None)
    r.set_reference(b.ref)
    ops, off, _ = r.realign(b, p)
    assert got["ops"] == ops.tolist() and got["off"] == off.tolist() and got["cells"] == r.cells
    _, _, post = r.realign(b, p, want_posteriors=True)              # posterior pairs come back through the shards too
    assert all(got["post"][k] == post[k].tolist() for k in ("off", "ref_pos", "read_pos", "prob_1e7"))
    assert len(got["post"]["ref_pos"]) > 0
    st = r.expectations(b, p)
    assert FixedStats(got["hi"], got["lo"]) == st                        # exact: independent of the sharding
    masks = [np.ones(b.n, np.uint8), np.array([1, 0, 1, 0, 0, 1, 0], np.uint8)]
    t = r.base_expectations(b, p, masks=masks)                           # per-position base expectations: int64 all-reduce
    assert np.array_equal(np.array(got["tables"], dtype=np.int64), t) and t[1].sum() > 0 and t[0].sum() > t[1].sum()
    r.set_hmm(Hmm.loadHmm(os.path.join(HERE, "golden", "blasr_hmm_0.txt")))
    ops2, off2, _ = r.realign(b, p)
    assert got["ops_trained"] == ops2.tolist() and got["off_trained"] == off2.tolist()
    assert got["ops_trained"] != got["ops"]                              # the broadcast model was really used


def test_command_line_under_two_ranks_writes_the_single_process_sam(tmp_path):
    """scripts/realign_sam.py under world_size 2 (parallel.run: rank 0 runs the pipeline, rank 1 serves): the SAM it
    writes is byte-identical to a single-process run, including a trained model broadcast to the other rank."""
    import filecmp
    import importlib.util
    from nanopore_b200 import realign
    from helpers_sam import make_experiment
    ref_fa, fq, sam_path, _ = make_experiment(str(tmp_path), n_reads=7, read_len=350, seed=23)
    out2 = str(tmp_path / "out_2ranks.sam")
    port = free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_cli_worker.py"), sam_path, fq, ref_fa, out2,
                                       "--trained", "blasr_hmm_40.txt"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0, o
    spec = importlib.util.spec_from_file_location("realign_sam", os.path.join(os.path.dirname(HERE), "scripts", "realign_sam.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out1 = str(tmp_path / "out_1rank.sam")
    prev = realign.setRealignerFactory(oracle_realigner_factory())
    try:
        assert mod.main([sam_path, fq, ref_fa, out1, "--trained", "blasr_hmm_40.txt"]) == 0
    finally:
        realign.setRealignerFactory(prev)
    assert filecmp.cmp(out1, out2, shallow=False)


def test_baseline_configs_script_dry_run_on_two_ranks(tmp_path):
    """scripts/configs_multi_gpu.py (BASELINE.json configs 3, 4, 5 through ShardedRealigner) at toy size over gloo: every
    config's sample equals the single-rank run, CIGARs span the reads, EM likelihoods rise."""
    out = str(tmp_path / "configs.json")
    port = free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_configs_worker.py"), "--scale", "0.00002", "--len-scale", "0.03",
                                       "--procs", "2", "--verify", "4", "--out", out], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0, o
    lines = [json.loads(ln) for ln in open(out)]
    assert [ln["config"] for ln in lines] == [3, 4, 5]
    assert lines[0]["sample_equals_single_gpu"] and lines[0]["cigars_span_reads"] and len(lines[0]["rank_cells"]) == 2
    assert lines[1]["sample_statistics_equal_single_gpu"] and lines[1]["monotone"] and len(lines[1]["running_likelihoods"]) == 5
    assert lines[2]["sample_equals_single_gpu"] and lines[2]["reads"] == 20 and min(lines[2]["rank_cells"]) > 0


def test_pack_shard_equals_pack_of_subset():
    """The scatter packs a shard straight into its staging buffer (native ragged gather): same wire blob content as
    packing batch.subset(idx), for ordinary, single-read and empty shards."""
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), RANK="0", WORLD_SIZE="1")
        dist.init_process_group("gloo")
        created = True
    try:
        b = synth.make_batch(300, 400, 3000, seed=2)
        rng = np.random.default_rng(1)
        for idx in (np.sort(rng.choice(300, 120, replace=False)), np.array([7]), np.zeros(0, np.int64), np.arange(300)):
            got = parallel._unpack(parallel._pack_shard(b, idx, 1).numpy().copy())
            sub = b.subset(idx)
            for a, f in zip(got, parallel._BATCH_FIELDS):
                assert np.array_equal(a, getattr(sub, f)) and a.dtype == getattr(sub, f).dtype, f
    finally:
        if created:
            dist.destroy_process_group()
