"""One rank of the world_size-2 gloo run of scripts/realign_sam.py (launched by test_parallel_gloo.py): rank 0 runs the
command line, rank 1 serves its calls; the CPU checker stands in for the GPU library on both."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch.distributed as dist                              # noqa: E402
from oracle_ctx import oracle_realigner_factory               # noqa: E402

if __name__ == "__main__":
    dist.init_process_group("gloo")
    spec = importlib.util.spec_from_file_location("realign_sam", os.path.join(ROOT, "scripts", "realign_sam.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rc = mod.main(sys.argv[1:], local_factory=lambda: oracle_realigner_factory()(None))
    sys.exit(rc or 0)
