#!/usr/bin/env python
"""modifyHmm.py IN.hmm GC_CONTENT SUBSTITUTION_RATE OUT.hmm

Command-line form of the HMM post-processing of the realignment path (reference
scripts/modifyHmm.py:7-30): load a trained pair-HMM, rescale the emissions to the background
frequencies implied by the reference GC content (utils.py:614-619), fold the expected variation
rate into the match emissions (utils.py:621-624), report the per-state marginals and write the
result.  `blasr_hmm_0.txt 0.5 0.2` reproduces the reference's blasr_hmm_20.txt, `0.4` its
blasr_hmm_40.txt (tests/test_hmm_kat.py).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nanopore_b200.hmm import (SYMBOL_NUMBER, Hmm, modifyHmmEmissionsByExpectedVariationRate,  # noqa: E402
                               normaliseHmmByReferenceGCContent, toMatrix)


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) != 5:
        sys.stderr.write(__doc__)
        return 2
    print("ARGS", argv)
    hmm = Hmm.loadHmm(argv[1])
    gcContent = float(argv[2])
    print("Got GC content", gcContent)
    normaliseHmmByReferenceGCContent(hmm, gcContent)
    substitutionRate = float(argv[3])
    print("Got substitution rate", substitutionRate)
    modifyHmmEmissionsByExpectedVariationRate(hmm, substitutionRate)
    sq = SYMBOL_NUMBER ** 2
    for state in range(hmm.stateNumber):
        n = toMatrix(hmm.emissions[sq * state:sq * (state + 1)])
        print("For state, ref frequencies", [sum(r) for r in n])
        print("For state, read frequencies", [sum(c) for c in zip(*n)])
    hmm.write(argv[4])
    return 0


if __name__ == "__main__":
    sys.exit(main())
