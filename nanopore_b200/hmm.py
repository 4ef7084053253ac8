"""HMM model file and post-processing (hot-path rows a9-a11 of SURVEY.md 8(a)).

`Hmm` mirrors the object the reference gets from
`cactus.bar.cactus_expectationMaximisation.Hmm` (absent upstream module; used
at reference nanopore/analyses/utils.py:534,538 and scripts/modifyHmm.py:10,30):
attributes `type`, `stateNumber`, `transitions`, `emissions`, `likelihood`,
classmethod `loadHmm(file)`, method `write(file)`.

File layout (nanopore/mappers/blasr_hmm_0.txt:1-2):
  line 1: <type> t[0..n*n-1] <likelihood>      transitions row-major from*n+to
  line 2: e[state*16 + x*4 + y]                x = reference base, y = read base
Types: 0 fiveState, 1 fiveStateAsymmetric, 2 threeState,
3 threeStateAsymmetric.
"""
import numpy as np

SYMBOL_NUMBER = 4

FIVE_STATE = 0
FIVE_STATE_ASYMMETRIC = 1
THREE_STATE = 2
THREE_STATE_ASYMMETRIC = 3

_STATE_NUMBER = {FIVE_STATE: 5, FIVE_STATE_ASYMMETRIC: 5, THREE_STATE: 3, THREE_STATE_ASYMMETRIC: 3}
_TYPE_NAMES = {"fiveState": FIVE_STATE, "fiveStateAsymmetric": FIVE_STATE_ASYMMETRIC,
               "threeState": THREE_STATE, "threeStateAsymmetric": THREE_STATE_ASYMMETRIC}


class Hmm:
    def __init__(self, modelType=FIVE_STATE_ASYMMETRIC):
        if isinstance(modelType, str):
            modelType = _TYPE_NAMES[modelType]
        self.type = int(modelType)
        self.stateNumber = _STATE_NUMBER[self.type]
        self.transitions = [0.0] * self.stateNumber ** 2
        self.emissions = [0.0] * (SYMBOL_NUMBER ** 2 * self.stateNumber)
        self.likelihood = 0.0

    # -- file I/O ---------------------------------------------------------
    @staticmethod
    def loadHmm(file):
        with open(file, "r") as fh:
            l1 = fh.readline().split()
            l2 = fh.readline().split()
        hmm = Hmm(int(l1[0]))
        n = hmm.stateNumber
        if len(l1) != n * n + 2:
            raise RuntimeError("Got the wrong number of transitions in %s: %d" % (file, len(l1) - 2))
        hmm.transitions = [float(v) for v in l1[1:-1]]
        hmm.likelihood = float(l1[-1])
        if len(l2) != n * SYMBOL_NUMBER ** 2:
            raise RuntimeError("Got the wrong number of emissions in %s: %d" % (file, len(l2)))
        hmm.emissions = [float(v) for v in l2]
        return hmm

    def write(self, file):
        with open(file, "w") as fh:
            fh.write("%s " % self.type + " ".join(repr(float(v)) for v in self.transitions) +
                     " %s\n" % repr(float(self.likelihood)))
            fh.write(" ".join(repr(float(v)) for v in self.emissions) + "\n")

    # -- expectations -> probabilities (M-step) ------------------------------
    def normalise(self):
        """Row-normalise transitions and per-state normalise emissions."""
        n = self.stateNumber
        t = np.array(self.transitions, dtype=np.float64).reshape(n, n)
        s = t.sum(axis=1, keepdims=True)
        s[s == 0.0] = 1.0
        self.transitions = (t / s).reshape(-1).tolist()
        e = np.array(self.emissions, dtype=np.float64).reshape(n, SYMBOL_NUMBER ** 2)
        s = e.sum(axis=1, keepdims=True)
        s[s == 0.0] = 1.0
        self.emissions = (e / s).reshape(-1).tolist()

    def arrays(self):
        """(trans[25], emis[80]) float64 arrays for the C ABI."""
        if self.stateNumber != 5:
            raise RuntimeError("only 5-state models are supported on this path (model type %d)" % self.type)
        return np.array(self.transitions, dtype=np.float64), np.array(self.emissions, dtype=np.float64)


def toMatrix(e):
    """utils.py:611"""
    return [list(e[SYMBOL_NUMBER * i:SYMBOL_NUMBER * (i + 1)]) for i in range(SYMBOL_NUMBER)]


def fromMatrix(m):
    """utils.py:612"""
    out = []
    for row in m:
        out.extend(list(row))
    return out


def normaliseHmmByReferenceGCContent(hmm, gcContent):
    """utils.py:614-619 -- rescales every reference-base row of the emission
    matrices to the background frequency implied by gcContent; insert states
    (2, 4) have no reference base and are left alone."""
    sq = SYMBOL_NUMBER ** 2
    for state in range(hmm.stateNumber):
        if state not in (2, 4):
            n = toMatrix(hmm.emissions[sq * state:sq * (state + 1)])
            rows = []
            for i in range(SYMBOL_NUMBER):
                tot = sum(n[i])
                bg = gcContent / 2.0 if i in (1, 2) else (1.0 - gcContent) / 2.0
                rows.append([(n[i][j] / tot) * bg for j in range(SYMBOL_NUMBER)])
            hmm.emissions[sq * state:sq * (state + 1)] = fromMatrix(rows)


def modifyHmmEmissionsByExpectedVariationRate(hmm, substitutionRate):
    """utils.py:621-624 -- match emissions := M . N with N[i][j] = 1-r on the
    diagonal and r/3 elsewhere.  The reference's `i / SYMBOL_NUMBER` is Python-2
    integer division; `//` here."""
    sq = SYMBOL_NUMBER ** 2
    n = np.array(toMatrix([(1.0 - substitutionRate) if i % SYMBOL_NUMBER == i // SYMBOL_NUMBER
                           else substitutionRate / (SYMBOL_NUMBER - 1) for i in range(sq)]))
    hmm.emissions[:sq] = fromMatrix(np.dot(np.array(toMatrix(hmm.emissions[:sq])), n).tolist())


def setHmmIndelEmissionsToBeFlat(hmm):
    """utils.py:626-629"""
    sq = SYMBOL_NUMBER ** 2
    for state in range(1, hmm.stateNumber):
        hmm.emissions[sq * state:sq * (state + 1)] = [1.0 / sq] * sq
