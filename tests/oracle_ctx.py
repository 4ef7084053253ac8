"""CPU checker behind the PhmmContext interface, for `-m "not gpu"` tests of the HOST logic only (SAM plumbing,
chunking, sharding, reduction order, EM bookkeeping).  Test infrastructure: the product never imports this."""
import numpy as np

import oracle


class OracleContext:
    def __init__(self, trans=None, emis=None):
        self.set_model(trans, emis)
        self.ref = None
        self._cells = 0
        self.calls = []

    def set_model(self, trans=None, emis=None, model_type=1):
        self.model = oracle.Model(trans, emis) if trans is not None else oracle.Model()

    def set_reference(self, codes):
        self.ref = np.ascontiguousarray(codes, dtype=np.uint8)

    @staticmethod
    def _oparams(p):
        return oracle.make_params(expansion=p.band, trim=p.anchor_trim, split_side=p.split_side, gap_gamma=p.gap_gamma,
                                  match_gamma=p.match_gamma, min_diags=p.min_diags, tb_diags=p.tb_diags, threshold=p.threshold)

    def realign_batch(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params, want_posteriors=False):
        n = len(read_off) - 1
        self.calls.append(("realign", n))
        op = self._oparams(params)
        ops, off = [], [0]
        post = {"off": [0], "ref_pos": [], "read_pos": [], "prob_1e7": []}
        self._cells = 0
        for i in range(n):
            r = oracle.realign(self.model, self.ref[ref_start[i]:ref_end[i]], reads[read_off[i]:read_off[i + 1]],
                               in_ops[in_off[i]:in_off[i + 1]], op)
            self._cells += r["cells"]
            ops.append(r["ops"])
            off.append(off[-1] + len(r["ops"]))
            o = np.lexsort((r["py"], r["px"]))
            post["ref_pos"].append(r["px"][o]); post["read_pos"].append(r["py"][o]); post["prob_1e7"].append(r["pw"][o])
            post["off"].append(post["off"][-1] + len(o))
        cat = lambda l, dt: np.concatenate(l).astype(dt) if l else np.zeros(0, dt)
        pd = None
        if want_posteriors:
            pd = {"off": np.array(post["off"], np.int64), "ref_pos": cat(post["ref_pos"], np.int32),
                  "read_pos": cat(post["read_pos"], np.int32), "prob_1e7": cat(post["prob_1e7"], np.int32)}
        return cat(ops, np.uint32), np.array(off, np.int64), pd

    def expectations_batch_fixed(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params):
        n = len(read_off) - 1
        self.calls.append(("expect", n))
        op = self._oparams(params)
        hi, lo = np.zeros(106, np.int64), np.zeros(106, np.int64)
        self._cells = 0
        for i in range(n):
            hi, lo, c = oracle.expectations_fixed(self.model, self.ref[ref_start[i]:ref_end[i]],
                                                  reads[read_off[i]:read_off[i + 1]], in_ops[in_off[i]:in_off[i + 1]], op, hi, lo)
            self._cells += c
        return hi, lo

    # split form + device tables of base expectations, restated on the host from the checker's posterior pairs
    def prepare(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params):
        self._prepared = (np.asarray(reads), np.asarray(read_off), np.asarray(ref_start), np.asarray(ref_end), np.asarray(in_ops),
                          np.asarray(in_off), params)
        self._post = None

    def run(self):
        if self._post is None:
            self._post = self.realign_batch(*self._prepared, want_posteriors=True)[2]

    def base_expectations_reset(self, n_tables=1):
        self._tables = np.zeros((n_tables, len(self.ref), 5), dtype=np.int64)

    def add_base_expectations(self, read_mask=None, table=0):
        self.run()
        reads, read_off, ref_start = self._prepared[0], self._prepared[1], self._prepared[2]
        post = self._post
        for i in range(len(read_off) - 1):
            if read_mask is not None and not read_mask[i]:
                continue
            s = slice(post["off"][i], post["off"][i + 1])
            b = reads[read_off[i] + post["read_pos"][s].astype(np.int64)]
            np.add.at(self._tables[table], (ref_start[i] + post["ref_pos"][s].astype(np.int64), np.minimum(b, 4)),
                      post["prob_1e7"][s].astype(np.int64))

    def base_expectations_fetch(self, ref_len, table=0):
        assert ref_len == len(self.ref)
        return self._tables[table].copy()

    def stats(self):
        return {"cells": self._cells}

    def close(self):
        pass


def oracle_realigner_factory(**kw):
    """factory(hmm=None) for nanopore_b200.realign.setRealignerFactory."""
    from nanopore_b200.engine import Realigner

    def factory(hmm=None):
        ctx = OracleContext()
        r = Realigner(ctx=ctx, **kw)
        r.set_hmm(hmm)
        return r
    return factory
