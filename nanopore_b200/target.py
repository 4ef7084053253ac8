"""In-process stand-in for the slice of jobTree's `Target` / `Stack` the realignment plugins use.

jobTree (an empty submodule in the reference) is a cluster task farm and out of scope as a scheduler; the
plugins only need something to subclass and to hang children / follow-ons on.  Surface mirrored (sites:
reference nanopore/pipeline.py:110-154,207; nanopore/analyses/utils.py:478,528,531,546-572;
nanopore/mappers/abstractMapper.py:21,28,38): `Target.__init__`, `run`, `addChildTarget`, `addChildTargetFn`,
`setFollowOnTarget`, `setFollowOnTargetFn`, `setFollowOnFn`, `getLocalTempDir`, `getGlobalTempDir`, `logToMaster`,
`Target.makeTargetFn`; `Stack(target).startJobTree(options)` -> number of failed jobs,
`Stack.addJobTreeOptions(parser)`.

Semantics kept: a target's children all finish before its follow-on starts; a failing target counts as one
failed job and its follow-on is not run.  The per-read fan-out the reference builds out of child targets
(utils.py:565-570) does not exist here -- the realignment of all reads is one batched library call.
"""
import os
import shutil
import tempfile
import traceback

from .bioio import logger


class Target:
    def __init__(self, time=None, memory=None, cpu=None):
        self._children = []
        self._follow_on = None
        self._stack = None
        self._local_tmp = None
        self._global_tmp = None

    # ---- to override ----
    def run(self):
        pass

    # ---- scheduling ----
    def addChildTarget(self, childTarget):
        self._children.append(childTarget)

    def addChildTargetFn(self, fn, args=(), time=None, memory=None, cpu=None):
        self.addChildTarget(_FnTarget(fn, args, pass_target=True))

    def addChildFn(self, fn, args=(), time=None, memory=None, cpu=None):
        self.addChildTarget(_FnTarget(fn, args, pass_target=False))

    def setFollowOnTarget(self, followOn):
        assert self._follow_on is None, "a target has at most one follow-on"
        self._follow_on = followOn

    def setFollowOnTargetFn(self, fn, args=(), time=None, memory=None, cpu=None):
        self.setFollowOnTarget(_FnTarget(fn, args, pass_target=True))

    def setFollowOnFn(self, fn, args=(), time=None, memory=None, cpu=None):
        self.setFollowOnTarget(_FnTarget(fn, args, pass_target=False))

    @staticmethod
    def makeTargetFn(fn, args=(), time=None, memory=None, cpu=None):
        return _FnTarget(fn, args, pass_target=True)

    # ---- services ----
    def getLocalTempDir(self):
        if self._local_tmp is None:
            self._local_tmp = tempfile.mkdtemp(prefix="local_", dir=self._stack.temp_root if self._stack else None)
        return self._local_tmp

    def getGlobalTempDir(self):
        """Lives until the whole job tree has finished (follow-ons read files their predecessor left here)."""
        if self._global_tmp is None:
            self._global_tmp = tempfile.mkdtemp(prefix="global_", dir=self._stack.temp_root if self._stack else None)
        return self._global_tmp

    def logToMaster(self, string):
        logger.info(string)
        if self._stack is not None:
            self._stack.messages.append(string)


class _FnTarget(Target):
    def __init__(self, fn, args, pass_target):
        Target.__init__(self)
        self.fn, self.args, self.pass_target = fn, tuple(args), pass_target

    def run(self):
        if self.pass_target:
            self.fn(self, *self.args)
        else:
            self.fn(*self.args)


class Stack:
    def __init__(self, target):
        self.target = target
        self.temp_root = None
        self.messages = []
        self.failures = []       # (target, formatted traceback)

    @staticmethod
    def addJobTreeOptions(parser):
        """Accepts the jobTree flags the reference launcher passes (pipeline.sh:9) so command lines carry over."""
        add = parser.add_option if hasattr(parser, "add_option") else parser.add_argument
        add("--jobTree", dest="jobTree", default=None)
        add("--logInfo", dest="logInfo", action="store_true", default=False)
        add("--maxThreads", dest="maxThreads", default="4")
        add("--batchSystem", dest="batchSystem", default="singleMachine")
        add("--defaultMemory", dest="defaultMemory", default=None)
        add("--logFile", dest="logFile", default=None)
        add("--stats", dest="stats", action="store_true", default=False)

    def _execute(self, t):
        """Runs t, then its children (depth first, in the order added), then its follow-on."""
        while t is not None:
            t._stack = self
            try:
                t.run()
            except Exception:
                self.failures.append((t, traceback.format_exc()))
                logger.critical("target %s failed:\n%s", type(t).__name__, self.failures[-1][1])
                return False
            ok = True
            for c in t._children:
                ok = self._execute(c) and ok
            if t._local_tmp:
                shutil.rmtree(t._local_tmp, ignore_errors=True)
            if not ok:
                return False
            t = t._follow_on
        return True

    def startJobTree(self, options=None):
        self.temp_root = tempfile.mkdtemp(prefix="nanopore_b200_jobs_")
        try:
            self._execute(self.target)
        finally:
            shutil.rmtree(self.temp_root, ignore_errors=True)
        return len(self.failures)
