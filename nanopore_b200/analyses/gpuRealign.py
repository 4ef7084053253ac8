"""GpuRealign: the realignment path as an analysis plugin under the reference's driver
(nanopore/pipeline.py:131-142 constructs `analysis(readFastqFile, readType, referenceFastaFile, samFile, outputDir)`).

Reads <samFile>, chains and realigns it on the GPU exactly like `AbstractMapper.realignSamFile` does for the
`*Realign*` mappers, and writes <outputDir>/realigned.sam; then marks the analysis DONE.  Class attributes select
the variant the way the reference's mapper subclasses do (e.g. nanopore/mappers/last_params.py:15-38).
"""
import os

from ..mappers.abstractMapper import trainedModelPath
from ..realign import realignSamFileTargetFn
from .abstractAnalysis import AbstractAnalysis


class GpuRealign(AbstractAnalysis):
    gapGamma = 0.5
    matchGamma = 0.0
    doEm = False
    useTrainedModel = False
    trainedModelFile = "blasr_hmm_0.txt"

    def run(self):
        AbstractAnalysis.run(self)
        if self.useTrainedModel and self.doEm:
            raise RuntimeError("Attempting to train stock model")
        outputSamFile = os.path.join(self.outputDir, "realigned.sam")
        if self.doEm:
            hmmFile = os.path.join(self.outputDir, "hmm.txt")
        elif self.useTrainedModel:
            hmmFile = trainedModelPath(self.trainedModelFile, self.getGlobalTempDir())
        else:
            hmmFile = None
        self.addChildTargetFn(realignSamFileTargetFn, args=(self.samFile, outputSamFile, self.readFastqFile,
                                                            self.referenceFastaFile, self.gapGamma, self.matchGamma,
                                                            hmmFile, self.doEm))
        self.setFollowOnFn(self.finish)


class GpuRealignEm(GpuRealign):
    doEm = True


class GpuRealignTrainedModel(GpuRealign):
    useTrainedModel = True
