#!/usr/bin/env python
"""SAM in -> realigned SAM out on the GPU(s): the command-line face of the realignment path.

    python scripts/realign_sam.py mapping.sam reads.fq reference.fa realigned.sam [--hmm F | --em | --trained blasr_hmm_0.txt]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/realign_sam.py ...

Does what a `*Realign*` mapper of the reference does after mapping (reference nanopore/mappers/abstractMapper.py:25-39):
chain, optionally train the HMM by EM, realign, write the SAM with the new CIGARs.  Under torchrun rank 0 runs this
script's work and the other ranks serve its realignment calls (nanopore_b200.parallel)."""
import argparse
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanopore_b200 import parallel                                  # noqa: E402
from nanopore_b200.mappers.abstractMapper import AbstractMapper     # noqa: E402
from nanopore_b200.target import Stack                              # noqa: E402


def main(argv=None, local_factory=None):
    """local_factory: what each rank uses to realign its shard (default: the GPU Realigner of that rank)."""
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("sam"); ap.add_argument("reads_fastq"); ap.add_argument("reference_fasta"); ap.add_argument("out_sam")
    ap.add_argument("--gapGamma", type=float, default=0.5)
    ap.add_argument("--matchGamma", type=float, default=0.0)
    ap.add_argument("--em", action="store_true", help="train the HMM by EM first (writes <out_sam>.hmm.txt and .xml)")
    ap.add_argument("--trained", default=None, help="use a trained model file, e.g. blasr_hmm_0.txt / blasr_hmm_20.txt")
    args = ap.parse_args(argv)

    class Realign(AbstractMapper):
        def run(self):
            self.realignSamFile(gapGamma=args.gapGamma, matchGamma=args.matchGamma, doEm=args.em,
                                useTrainedModel=args.trained is not None, trainedModelFile=args.trained or "blasr_hmm_0.txt")

    def work():
        shutil.copyfile(args.sam, args.out_sam)                     # the mapper overwrites its outputSamFile in place
        m = Realign(args.reads_fastq, "reads", args.reference_fasta, args.out_sam, emptyHmmFile=args.out_sam + ".hmm.txt")
        failed = Stack(m).startJobTree(None)
        if failed:
            raise RuntimeError("%d job(s) failed" % failed)         # nanopore/pipeline.py:209-210
        return 0

    return parallel.run(work, local_factory)


if __name__ == "__main__":
    sys.exit(main() or 0)
