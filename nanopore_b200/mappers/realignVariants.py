"""The `*Chain` / `*Realign*` variants the reference declares by hand for every base mapper
(reference nanopore/mappers/last_params.py:10-38, and the same six classes in last.py, bwa.py, bwa_params.py, blasr.py,
blasr_params.py, lastz.py, lastzParams.py): run the base mapper, then post-process its SAM file through
AbstractMapper.chainSamFile / realignSamFile with the variant's arguments.

The base mappers themselves (LAST, BWA, BLASR, LASTZ wrappers) are outside this repository's scope (SURVEY.md 8:
external binaries).  `realignVariants(Base)` derives the six variants from ANY mapper class whose run() writes
self.outputSamFile -- the reference's own Last / LastParams / Bwa ... once nanopore_b200's AbstractMapper is the base
class they import, or `SamFileMapper` below, which "maps" by copying an existing SAM file.
"""
import shutil

from .abstractMapper import AbstractMapper

# suffix -> (method, keyword arguments), exactly the calls of last_params.py:13,18,23,28,33,38
VARIANTS = {
    "Chain": ("chainSamFile", {}),
    "Realign": ("realignSamFile", {}),
    "RealignEm": ("realignSamFile", dict(doEm=True, gapGamma=0.5, matchGamma=0.0)),
    "RealignTrainedModel": ("realignSamFile", dict(useTrainedModel=True)),
    "RealignTrainedModel20": ("realignSamFile", dict(useTrainedModel=True, trainedModelFile="blasr_hmm_20.txt")),
    "RealignTrainedModel40": ("realignSamFile", dict(useTrainedModel=True, trainedModelFile="blasr_hmm_40.txt")),
}


def realignVariants(Base):
    """{class name: class} of the six variants of mapper class Base, named Base.__name__ + suffix like the reference's."""
    out = {}
    for suffix, (method, kwargs) in VARIANTS.items():
        def run(self, _method=method, _kwargs=kwargs):
            Base.run(self)
            getattr(self, _method)(**_kwargs)
        name = Base.__name__ + suffix
        out[name] = type(name, (Base,), {"run": run, "__doc__": "%s, then %s(%s) (reference last_params.py:10-38)" % (
            Base.__name__, method, ", ".join("%s=%r" % kv for kv in kwargs.items()))})
    return out


class SamFileMapper(AbstractMapper):
    """A "mapper" whose mapping step copies an existing SAM file (class attribute or constructor keyword
    `mappedSamFile`) to outputSamFile: lets the variants run where no aligner binary is installed."""
    mappedSamFile = None

    def __init__(self, readFastqFile, readType, referenceFastaFile, outputSamFile, emptyHmmFile=None, mappedSamFile=None):
        super().__init__(readFastqFile, readType, referenceFastaFile, outputSamFile, emptyHmmFile)
        if mappedSamFile is not None:
            self.mappedSamFile = mappedSamFile

    def run(self):
        if self.mappedSamFile is None:
            raise RuntimeError("SamFileMapper needs mappedSamFile")
        shutil.copyfile(self.mappedSamFile, self.outputSamFile)


globals().update(realignVariants(SamFileMapper))        # SamFileMapperChain, SamFileMapperRealign, ...
