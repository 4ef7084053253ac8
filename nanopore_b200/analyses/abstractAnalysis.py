"""AbstractAnalysis: base class of analysis plugins (reference nanopore/analyses/abstractAnalysis.py:5-41), same
constructor and DONE-file protocol."""
import os

from ..bioio import logger
from ..target import Target


class AbstractAnalysis(Target):
    """Base class to for analysis targets. Inherit this class to create an analysis."""

    def __init__(self, readFastqFile, readType, referenceFastaFile, samFile, outputDir):
        Target.__init__(self)
        self.readFastqFile = readFastqFile
        self.referenceFastaFile = referenceFastaFile
        self.samFile = samFile
        self.outputDir = outputDir
        self.readType = readType

    def run(self):
        """Base method that does some logging."""
        logger.info("This analysis target has read fastq file: %s, reference fasta file: %s, sam file: %s and will "
                    "output to the directory: %s" % (self.readFastqFile, self.referenceFastaFile, self.samFile, self.outputDir))

    def finish(self):
        """Called when an analysis has finished successfully to indicate that it should not be repeated."""
        open(os.path.join(self.outputDir, "DONE"), "w").close()

    @staticmethod
    def reset(outputDir):
        if AbstractAnalysis.isFinished(outputDir):
            os.remove(os.path.join(outputDir, "DONE"))

    @staticmethod
    def isFinished(outputDir):
        return os.path.exists(os.path.join(outputDir, "DONE"))

    @staticmethod
    def formatRatio(numerator, denominator):
        if denominator == 0:
            return float("nan")
        return float(numerator) / denominator
