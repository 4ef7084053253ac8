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
    """Base class for mappers. Inherit this class to create a mapper."""

    def __init__(self, readFastqFile, readType, referenceFastaFile, outputSamFile, emptyHmmFile=None):
        Target.__init__(self)
        self.readFastqFile = readFastqFile
        self.referenceFastaFile = referenceFastaFile
        self.outputSamFile = outputSamFile
        self.readType = readType
        self.emptyHmmFile = emptyHmmFile

    def chainSamFile(self):
        """Converts the sam file so that there is at most one global alignment of each read."""
        tempSamFile = os.path.join(self.getLocalTempDir(), "temp.sam")
        shutil.copyfile(self.outputSamFile, tempSamFile)
        chainSamFile(tempSamFile, self.outputSamFile, self.readFastqFile, self.referenceFastaFile)

    def realignSamFile(self, gapGamma=0.5, matchGamma=0.0, doEm=False, useTrainedModel=False,
                       trainedModelFile="blasr_hmm_0.txt"):
        """Chains and then realigns the resulting global alignments."""
        tempSamFile = os.path.join(self.getGlobalTempDir(), "temp.sam")
        if useTrainedModel and doEm:
            raise RuntimeError("Attempting to train stock model")
        shutil.copyfile(self.outputSamFile, tempSamFile)
        if doEm:
            hmmFile = self.emptyHmmFile
        elif useTrainedModel:
            hmmFile = trainedModelPath(trainedModelFile, self.getGlobalTempDir())
        else:
            hmmFile = None
        self.addChildTargetFn(realignSamFileTargetFn, args=(tempSamFile, self.outputSamFile, self.readFastqFile,
                                                            self.referenceFastaFile, gapGamma, matchGamma, hmmFile, doEm))
