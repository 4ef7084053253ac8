"""Builds a small on-disk experiment (reference FASTA, read FASTQ, mapper-like SAM of local hits) for the host tests."""
import os

import numpy as np

from nanopore_b200 import synth
from nanopore_b200.bioio import fastaWrite, fastqWrite, reverseComplement


def make_experiment(dirname, n_reads=6, read_len=400, contig_lens=(1500, 1100), seed=0, hits=(1, 2, 3), unmapped=1):
    rng = np.random.default_rng(seed)
    os.makedirs(dirname, exist_ok=True)
    contigs = [("ref%d some description" % k, synth.random_reference(L, rng)) for k, L in enumerate(contig_lens)]
    ref_fa = os.path.join(dirname, "reference.fa")
    with open(ref_fa, "w") as fh:
        for name, codes in contigs:
            fastaWrite(fh, name, synth.decode(codes))
    fq = os.path.join(dirname, "reads.fq")
    sam = os.path.join(dirname, "mapping.sam")
    lines = ["@HD\tVN:1.0\tSO:unsorted"] + ["@SQ\tSN:%s\tLN:%d" % (n.split()[0], len(c)) for n, c in contigs]
    truth = {}
    with open(fq, "w") as fh:
        for i in range(n_reads + unmapped):
            name = "read_%d" % i
            k = int(rng.integers(0, len(contigs)))
            cname, ref = contigs[k][0].split()[0], contigs[k][1]
            start = int(rng.integers(0, len(ref) - read_len + 1))
            codes, runs = synth.simulate_read(ref, start, read_len, rng)
            aligned_seq = synth.decode(codes)                      # reference-strand orientation
            reverse = bool(rng.random() < 0.5)
            fastqWrite(fh, name + " extra words", reverseComplement(aligned_seq) if reverse else aligned_seq, [50] * len(aligned_seq))
            if i >= n_reads:
                lines.append("\t".join([name, "4", "*", "0", "0", "*", "*", "0", "0", aligned_seq, "*"]))
                continue
            ops = synth.unpack_ops(runs)
            # cut the true script into local hits at M runs
            m_idx = [j for j, (c, _) in enumerate(ops) if c == 0]
            nh = min(int(rng.choice(hits)), len(m_idx))
            cuts = sorted(rng.choice(len(m_idx), size=nh, replace=False).tolist()) if nh > 1 else [0]
            bounds = []
            for h in range(nh):
                a = m_idx[cuts[h]] if nh > 1 else m_idx[0]
                b = (m_idx[cuts[h + 1] - 1] if h + 1 < nh else m_idx[-1]) if nh > 1 else m_idx[-1]
                if b < a:
                    b = a
                bounds.append((a, b))
            truth[name] = (cname, start, reverse, len(aligned_seq))
            for h, (a, b) in enumerate(bounds):
                q0 = sum(l for c, l in ops[:a] if c in (0, 1))
                r0 = sum(l for c, l in ops[:a] if c in (0, 2))
                q1 = q0 + sum(l for c, l in ops[a:b + 1] if c in (0, 1))
                hit = "".join("%d%s" % (l, "MID"[c]) for c, l in ops[a:b + 1])
                L = len(aligned_seq)
                if h % 2 == 0:                                      # soft clips: whole sequence in SEQ
                    cig = ("%dS" % q0 if q0 else "") + hit + ("%dS" % (L - q1) if L - q1 else "")
                    seq = aligned_seq
                else:                                               # hard clips: only the aligned part in SEQ
                    cig = ("%dH" % q0 if q0 else "") + hit + ("%dH" % (L - q1) if L - q1 else "")
                    seq = aligned_seq[q0:q1]
                flag = (16 if reverse else 0) | (2048 if h else 0)
                lines.append("\t".join([name, str(flag), cname, str(start + r0 + 1), "30", cig, "*", "0", "0", seq, "*", "NM:i:0"]))
    with open(sam, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return ref_fa, fq, sam, truth
