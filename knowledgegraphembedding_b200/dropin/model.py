"""Drop-in replacement for the reference's `codes/model.py`.

`codes/run.py` does `from model import KGEModel` (run.py:18).  Put this directory in front of `codes/` on
sys.path -- or replace codes/model.py by this one file -- and run.py trains and evaluates through the
B200 kernels unchanged.  See INTEGRATION.md.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from knowledgegraphembedding_b200.model import KGEModel  # noqa: E402,F401
