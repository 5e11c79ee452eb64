"""Pin the C oracle (oracle/kge_oracle.c) against the numpy oracle (itself pinned to the reference's golden
vectors) and against the golden vectors directly.  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relinf
from oracle import c_oracle as C
from oracle import kge_oracle as O
from knowledgegraphembedding_b200.filter_index import FilterIndex

MODELS = ["TransE", "DistMult", "ComplEx", "RotatE", "pRotatE"]
FLAGS = {"TransE": (False, False), "DistMult": (False, False), "ComplEx": (True, True),
         "RotatE": (True, False), "pRotatE": (False, False)}


def test_sincos_accuracy():
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-200, 200, 200000), rng.uniform(-4, 4, 200000), [0.0, 1e6, -3e7]]).astype(np.float32)
    s, c = C.sincos(x)
    assert np.max(np.abs(s - np.sin(x.astype(np.float64)))) < 2.0 * 2 ** -24
    assert np.max(np.abs(c - np.cos(x.astype(np.float64)))) < 2.0 * 2 ** -24


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("d", [12, 10])
def test_forward_matches_reference_golden(model, d):
    g = np.load(os.path.join(GOLDEN, f"small_{model}_d{d}.npz"))
    st = {"entity_embedding": g["init_entity_embedding"], "relation_embedding": g["init_relation_embedding"]}
    if model == "pRotatE":
        st["modulus"] = g["init_modulus"]
    rho = O.embedding_range(float(g["gamma"]), d)
    for mode in O.MODES:
        sample = g["positive"] if mode == "single" else (g["positive"], g["negative"])
        s = C.forward(model, st, sample, mode, float(g["gamma"]), rho)
        assert relinf(s, g["score_" + mode]) < 1e-5, (model, mode)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("d", [12, 10])
def test_eval_scores_and_ranks_match_reference_golden(model, d):
    g = np.load(os.path.join(GOLDEN, f"small_{model}_d{d}.npz"))
    st = {"entity_embedding": g["eval_E"], "relation_embedding": g["eval_R"]}
    if model == "pRotatE":
        st["modulus"] = g["init_modulus"]
    nentity, gamma = int(g["nentity"]), float(g["gamma"])
    rho = O.embedding_range(gamma, d)
    test, all_true = g["eval_test"], g["eval_all_true"]
    index = FilterIndex(all_true, nentity, int(g["nrelation"]))
    ranks, rows = [], []
    for mode in ("head-batch", "tail-batch"):
        off, ent = index.csr(test, mode)
        sc = C.eval_scores(model, st, test, mode, gamma, rho, off, ent)
        rows.append(sc)
        r = C.ranks_from_scores(sc, test, mode)
        # the count formula equals the reference procedure (stable descending argsort) on the same matrix
        pos_col = test[:, 0] if mode == "head-batch" else test[:, 2]
        np.testing.assert_array_equal(r, [O.rank_from_scores(row, p) for row, p in zip(sc, pos_col)])
        ranks.append(r)
    assert relinf(np.concatenate(rows), g["eval_scores"]) < 1e-5
    np.testing.assert_array_equal(np.concatenate(ranks), g["eval_ranks"])


def test_filter_index_matches_reference_encoding():
    """FilterIndex + bias encoding == dataloader.py:134-154 (checked through the numpy oracle's restatement)."""
    rng = np.random.RandomState(5)
    nentity, nrelation = 50, 4
    all_true = sorted({(int(rng.randint(nentity)), int(rng.randint(nrelation)), int(rng.randint(nentity)))
                       for _ in range(400)})
    test = [all_true[i] for i in rng.choice(len(all_true), 30, replace=False)] + [(1, 2, 3), (49, 0, 49)]
    th, tt = O.build_true_sets(all_true)
    index = FilterIndex(all_true, nentity, nrelation)
    for mode in ("head-batch", "tail-batch"):
        off, ent = index.csr(test, mode)
        assert off[0] == 0 and off[-1] == ent.size and len(off) == len(test) + 1
        for i, tr in enumerate(test):
            _, bias, pos = O.candidates_and_bias(tr, mode, nentity, th, tt)
            got = set(int(e) for e in ent[off[i]:off[i + 1]])
            assert got - {pos} == set(np.nonzero(bias)[0].tolist())
    empty = FilterIndex([], nentity, nrelation)
    off, ent = empty.csr(test, "head-batch")
    assert off[-1] == 0 and ent.size == 0


def test_full_width_rows_agree_with_numpy_oracle():
    """FB15k-width rows (d=1000): index-order fp32 accumulation vs numpy's pairwise sum stays inside 1e-5."""
    rng = np.random.RandomState(3)
    for model in MODELS:
        de, dr = FLAGS[model]
        gamma, d = 24.0, 1000
        st = O.init_tables(model, 300, 7, d, gamma, de, dr, seed=1)
        pos = np.stack([rng.randint(300, size=8), rng.randint(7, size=8), rng.randint(300, size=8)], 1)
        neg = rng.randint(300, size=(8, 64))
        rho = O.embedding_range(gamma, d)
        for mode in ("head-batch", "tail-batch"):
            a = C.forward(model, st, (pos, neg), mode, gamma, rho)
            b = O.forward(model, st, (pos, neg), mode, gamma, d)
            assert relinf(a, b) < 1e-5, (model, mode)


CFGS = {
    "adv_sub": dict(adversarial=True, alpha=0.7, uni_weight=False),
    "adv_uni": dict(adversarial=True, alpha=1.0, uni_weight=True),
    "mean_sub": dict(adversarial=False, uni_weight=False),
    "adv_sub_reg": dict(adversarial=True, alpha=1.0, uni_weight=False, regularization=1e-3),
}


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("d", [12, 10])
@pytest.mark.parametrize("cfg", list(CFGS))
def test_train_steps_match_reference_golden(model, d, cfg):
    """ko_train_step over the 4 golden steps (losses, first-step grads, final tables) -- the function bench.py times
    as the CPU baseline."""
    g = np.load(os.path.join(GOLDEN, f"small_{model}_d{d}.npz"))
    st = {"entity_embedding": g["init_entity_embedding"], "relation_embedding": g["init_relation_embedding"]}
    if model == "pRotatE":
        st["modulus"] = g["init_modulus"]
    ts = C.TrainState(model, st, float(g["gamma"]), d)
    lr = 1e-3
    for step in range(4):
        if step == 2:
            lr /= 10
            ts.reset_optimizer()
        batch = (g[f"train_{cfg}_pos{step}"], g[f"train_{cfg}_neg{step}"], g[f"train_{cfg}_w{step}"],
                 "tail-batch" if step % 2 == 0 else "head-batch")
        log, grads = C.train_step(ts, batch, lr=lr, return_grads=True, **CFGS[cfg])
        ref = g[f"train_{cfg}_logs"][step]
        got = [log.get("regularization", 0.0), log["positive_sample_loss"], log["negative_sample_loss"], log["loss"]]
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-7)
        if step == 0:
            assert relinf(grads["entity_embedding"], g[f"train_{cfg}_gE0"]) < 1e-5
            assert relinf(grads["relation_embedding"], g[f"train_{cfg}_gR0"]) < 1e-5
            if model == "pRotatE":
                assert relinf(grads["modulus"], g[f"train_{cfg}_gM0"]) < 1e-5
    assert relinf(ts.state["entity_embedding"], g[f"train_{cfg}_E"]) < 1e-5
    assert relinf(ts.state["relation_embedding"], g[f"train_{cfg}_R"]) < 1e-5
    if model == "pRotatE":
        assert relinf(ts.state["modulus"], g[f"train_{cfg}_M"]) < 1e-5
