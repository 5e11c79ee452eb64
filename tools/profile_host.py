"""Host-side profile of KGEModel.train_step (cProfile) on one GPU: where the Python time of a step goes."""
import cProfile
import os
import pstats
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                             # noqa: E402
from knowledgegraphembedding_b200 import KGEModel        # noqa: E402
from oracle import kge_oracle as O                       # noqa: E402


def main():
    model, nentity, nrel, d, gamma, B, N, lr, de, dr = bench.WORKLOADS["rotate_fb15k"]
    dev = torch.device("cuda", 0)
    st = O.init_tables(model, nentity, nrel, d, gamma, de, dr, seed=0)
    m = KGEModel(model, nentity, nrel, d, gamma, double_entity_embedding=de, double_relation_embedding=dr)
    with torch.no_grad():
        m.entity_embedding.copy_(torch.from_numpy(st["entity_embedding"]))
        m.relation_embedding.copy_(torch.from_numpy(st["relation_embedding"]))
    m = m.to(dev)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=lr)
    args = types.SimpleNamespace(cuda=True, negative_adversarial_sampling=True, adversarial_temperature=1.0,
                                 uni_weight=False, regularization=0.0)
    pool = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(n).pin_memory(), torch.from_numpy(w).pin_memory(), md)
            for p, n, w, md in bench.make_batches(nentity, nrel, B, N, 8, seed=1)]

    class It:
        i = 0

        def __next__(self):
            It.i += 1
            return pool[It.i % len(pool)]

    it = It()
    for _ in range(20):
        KGEModel.train_step(m, opt, it, args)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        KGEModel.train_step(m, opt, it, args)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
