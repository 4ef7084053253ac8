"""AbstractMapper: the plugin base class every `*Realign*` mapper of the reference goes through
(reference nanopore/mappers/abstractMapper.py:7-39), with the same constructor, `chainSamFile()` and
`realignSamFile(gapGamma, matchGamma, doEm, useTrainedModel, trainedModelFile)`.  The realignment it schedules is
the GPU path of nanopore_b200.realign instead of one cactus_realign process per read.
"""
import os
import re
import shutil

from ..hmm import Hmm, modifyHmmEmissionsByExpectedVariationRate, normaliseHmmByReferenceGCContent
from ..realign import chainSamFile, realignSamFileTargetFn
from ..target import Target


def pathToTrainedModels():
    """Directory of the trained HMM file shipped with the mappers (reference nanopore/mappers/blasr_hmm_0.txt)."""
    return os.path.dirname(os.path.abspath(__file__))


def trainedModelPath(trainedModelFile, workDir):
    """Resolves a trainedModelFile name.  blasr_hmm_0.txt ships with the package; blasr_hmm_<R>.txt (the reference
    ships R = 20 and 40) is blasr_hmm_0.txt pushed through scripts/modifyHmm.py with GC 0.5 and substitution rate
    R/100 -- that is how the reference's own files were made (they are reproduced to 5e-13, tests/test_hmm_kat.py)
    -- so it is derived on demand into workDir instead of being stored."""
    path = os.path.join(pathToTrainedModels(), trainedModelFile)
    if os.path.exists(path):
        return path
    m = re.match(r"^blasr_hmm_([0-9]+)\.txt$", trainedModelFile)
    if m is None:
        raise RuntimeError("Trained model file %s not found in %s" % (trainedModelFile, pathToTrainedModels()))
    hmm = Hmm.loadHmm(os.path.join(pathToTrainedModels(), "blasr_hmm_0.txt"))
    normaliseHmmByReferenceGCContent(hmm, 0.5)
    modifyHmmEmissionsByExpectedVariationRate(hmm, int(m.group(1)) / 100.0)
    out = os.path.join(workDir, trainedModelFile)
    hmm.write(out)
    return out


class AbstractMapper(Target):
    """A mapper plugin produces `outputSamFile` in run(); the two methods below post-process that file in place."""

    def __init__(self, readFastqFile, readType, referenceFastaFile, outputSamFile, emptyHmmFile=None):
        super().__init__()
        self.readFastqFile, self.readType = readFastqFile, readType
        self.referenceFastaFile = referenceFastaFile
        self.outputSamFile = outputSamFile
        self.emptyHmmFile = emptyHmmFile            # where doEm writes the trained model (pipeline.py:121)

    def _scratch_copy(self, directory):
        scratch = os.path.join(directory, "temp.sam")
        shutil.copyfile(self.outputSamFile, scratch)
        return scratch

    def chainSamFile(self):
        """Rewrites outputSamFile with at most one (global) alignment per read and reference."""
        chainSamFile(self._scratch_copy(self.getLocalTempDir()), self.outputSamFile, self.readFastqFile,
                     self.referenceFastaFile)

    def realignSamFile(self, gapGamma=0.5, matchGamma=0.0, doEm=False, useTrainedModel=False,
                       trainedModelFile="blasr_hmm_0.txt"):
        """Schedules chain -> [EM] -> realign of outputSamFile (same defaults and model choice as the reference:
        doEm trains into emptyHmmFile, useTrainedModel loads a shipped / derived model, neither = stock model)."""
        if doEm and useTrainedModel:
            raise RuntimeError("Attempting to train stock model")
        hmmFile = None
        if doEm:
            hmmFile = self.emptyHmmFile
        elif useTrainedModel:
            hmmFile = trainedModelPath(trainedModelFile, self.getGlobalTempDir())
        self.addChildTargetFn(realignSamFileTargetFn,
                              args=(self._scratch_copy(self.getGlobalTempDir()), self.outputSamFile, self.readFastqFile,
                                    self.referenceFastaFile, gapGamma, matchGamma, hmmFile, doEm))
