"""One rank of the CPU dry run of scripts/configs_multi_gpu.py (launched by test_parallel_gloo.py): the script's own
main() with the CPU checker standing in for the GPU library on every rank, tiny read counts and lengths."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from nanopore_b200 import parallel                              # noqa: E402
from oracle_ctx import oracle_realigner_factory                  # noqa: E402

spec = importlib.util.spec_from_file_location("configs_multi_gpu", os.path.join(ROOT, "scripts", "configs_multi_gpu.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
parallel.init("gloo")
factory = lambda: oracle_realigner_factory()(None)
sys.exit(mod.main(sys.argv[1:], local_factory=factory, single_factory=factory))
