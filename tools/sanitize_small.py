#!/usr/bin/env python
"""Small end-to-end run of every kernel family, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knowledgegraphembedding_b200 import KGEModel  # noqa: E402
from knowledgegraphembedding_b200.sampler import BidirectionalGpuIterator  # noqa: E402

FLAGS = {"TransE": (False, False), "DistMult": (False, False), "ComplEx": (True, True), "RotatE": (True, False),
         "pRotatE": (False, False)}
rng = np.random.RandomState(0)
nentity, nrel = 700, 6
tri = sorted({(int(rng.randint(nentity)), int(rng.randint(nrel)), int(rng.randint(nentity))) for _ in range(3000)})
for model, (de, dr) in FLAGS.items():
    for d in (64, 10):
        m = KGEModel(model, nentity, nrel, d, 9.0, double_entity_embedding=de, double_relation_embedding=dr).cuda()
        opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=1e-3)
        it = BidirectionalGpuIterator(tri, nentity, nrel, 24, 40, "cuda", seed=1)
        args = types.SimpleNamespace(cuda=True, negative_adversarial_sampling=True, adversarial_temperature=1.0,
                                     uni_weight=False, regularization=1e-4 if model in ("ComplEx", "DistMult") else 0.0,
                                     countries=False, test_batch_size=8, test_log_steps=1000, nentity=nentity,
                                     nrelation=nrel)
        for _ in range(3):
            log = KGEModel.train_step(m, opt, it, args)
        s = m((torch.tensor(tri[:8]), torch.randint(nentity, (8, 5))), 'head-batch')
        s.sum().backward()
        metrics = KGEModel.test_step(m, tri[:50], tri, args)
        print(model, d, round(log["loss"], 4), round(metrics["MRR"], 4), flush=True)
torch.cuda.synchronize()
print("done")
