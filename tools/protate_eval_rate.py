#!/usr/bin/env python
"""pRotatE filtered-evaluation rate at FB15k shapes: exact kernel (KGE_EVAL_SIMT=1) vs the two-stage path."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from knowledgegraphembedding_b200 import KGEModel  # noqa: E402

nentity, nrel, d, gamma, nq = 14951, 1345, 1000, 24.0, 4096
torch.manual_seed(0)
m = KGEModel("pRotatE", nentity, nrel, d, gamma).cuda()
all_true, rng = bench.synthetic_triples(nentity, nrel, 483142, seed=2)
test = [all_true[i] for i in rng.choice(len(all_true), nq, replace=False)]
out = {}
for tag in ("exact", "two_stage"):
    if tag == "exact":
        os.environ["KGE_EVAL_SIMT"] = "1"
    else:
        os.environ.pop("KGE_EVAL_SIMT", None)
    ranks = None
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ranks = [m.filtered_ranks(test, all_true, mode) for mode in ("head-batch", "tail-batch")]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out[tag] = (2 * nq / dt, np.concatenate(ranks))
assert np.array_equal(out["exact"][1], out["two_stage"][1])
print({"exact_queries_per_s": out["exact"][0], "two_stage_queries_per_s": out["two_stage"][0],
       "ambiguous_pairs": m._ws.get("two_stage_last_ambiguous"), "queries": 2 * nq})
