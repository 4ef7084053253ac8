"""Analysis plugin base class of the realignment path.

Keeps the surface the reference's driver relies on (reference nanopore/analyses/abstractAnalysis.py:5-41 and its
call sites nanopore/pipeline.py:137-142): five positional constructor arguments stored under the same attribute
names, `run()` / `finish()`, and the static `reset` / `isFinished` / `formatRatio` helpers.  "Finished" is a
zero-length marker file called DONE inside the analysis' output directory.
"""
import os

from ..bioio import logger
from ..target import Target

_MARKER = "DONE"


def _marker(outputDir):
    return os.path.join(outputDir, _MARKER)


class AbstractAnalysis(Target):
    def __init__(self, readFastqFile, readType, referenceFastaFile, samFile, outputDir):
        super().__init__()
        self.readFastqFile, self.readType = readFastqFile, readType
        self.referenceFastaFile, self.samFile = referenceFastaFile, samFile
        self.outputDir = outputDir

    def run(self):
        """Subclasses call this first; it only logs what the analysis was given."""
        logger.info("analysis %s: reads %s, reference %s, sam %s -> %s", type(self).__name__, self.readFastqFile,
                    self.referenceFastaFile, self.samFile, self.outputDir)

    def finish(self):
        """Marks the analysis complete so that the driver does not schedule it again."""
        with open(_marker(self.outputDir), "w"):
            pass

    @staticmethod
    def isFinished(outputDir):
        return os.path.exists(_marker(outputDir))

    @staticmethod
    def reset(outputDir):
        """Forgets a completed run (no-op when there is none)."""
        try:
            os.remove(_marker(outputDir))
        except FileNotFoundError:
            pass

    @staticmethod
    def formatRatio(numerator, denominator):
        """numerator / denominator as float; NaN instead of a ZeroDivisionError."""
        return float(numerator) / denominator if denominator != 0 else float("nan")
