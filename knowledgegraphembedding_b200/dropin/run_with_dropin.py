#!/usr/bin/env python
"""Run the reference's UNMODIFIED `codes/run.py` on the B200 drop-in:

    python run_with_dropin.py /path/to/KnowledgeGraphEmbedding/codes/run.py --cuda --do_train ... (run.py's own flags)

`run.py` does `from model import KGEModel` (run.py:18); Python resolves a script's imports from the script's directory
first, so this launcher puts this directory (whose `model.py` re-exports the B200 `KGEModel`) in front of `codes/` and
then executes run.py as `__main__`.  `dataloader.py` still comes from `codes/`.

Two environment switches exist for A/B runs and tests:
  KGE_RUN_REFERENCE=1   do not shadow `model`: run.py runs entirely on the reference's own model.py
  KGE_RUN_SEED=<int>    seed python / numpy / torch before run.py starts (run.py itself never seeds), which makes the
                        parameter init, the DataLoader shuffling and the workers' negative sampling reproducible
"""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    codes = os.path.dirname(script)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (here, codes)]
    sys.path.insert(0, codes)
    if not os.environ.get("KGE_RUN_REFERENCE"):
        sys.path.insert(0, here)
    seed = os.environ.get("KGE_RUN_SEED")
    if seed is not None:
        import random

        import numpy as np
        import torch
        random.seed(int(seed))
        np.random.seed(int(seed))
        torch.manual_seed(int(seed))
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
