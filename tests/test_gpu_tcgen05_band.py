"""Pins the error band of the tcgen05 evaluation path (csrc/kge_eval_gemm.cu).  The rank counts of that path are exact
only if every tensor-core approximation lies within  band(K) * |q| * |E_j|  of the canonical fp32 score, and the band
rests on an assumption about how tcgen05.mma kind::tf32 accumulates.  Here the approximations themselves are dumped
(approx_scores_out) and compared with the exact kernel's scores -- which are bit-identical to the C oracle's -- for
K = 1000, 2000, 4000 and for adversarial operand magnitudes (large dynamic range inside a row, sign-alternating
cancellation, tiny and huge rows): the measured error must stay below HALF the band."""
import numpy as np
import pytest
import torch

from oracle import kge_oracle as O
from test_gpu_parity import FLAGS, make_model

pytestmark = pytest.mark.gpu


def adversarial_tables(model, nentity, nrel, d, gamma, kind, seed):
    de, dr = FLAGS[model]
    st = O.init_tables(model, nentity, nrel, d, gamma, de, dr, seed=seed)
    rng = np.random.RandomState(seed + 100)
    E = st["entity_embedding"].astype(np.float64)
    if kind == "dynamic_range":            # magnitudes spread over 2^-12 .. 2^12 inside every row
        E = E * np.exp2(rng.randint(-12, 13, size=E.shape))
    elif kind == "cancellation":           # near-equal magnitudes with alternating signs: large |q||E|, small sum
        E = np.abs(E).mean() * (1.0 + 1e-3 * rng.standard_normal(E.shape)) * np.where(np.arange(E.shape[1]) % 2, -1.0, 1.0)
    elif kind == "row_scales":             # whole rows tiny or huge
        E = E * np.exp2(rng.randint(-20, 21, size=(E.shape[0], 1)))
    st["entity_embedding"] = E.astype(np.float32)
    return st


@pytest.mark.parametrize("model,d", [("DistMult", 1000), ("ComplEx", 1000), ("DistMult", 4000), ("ComplEx", 500)])
@pytest.mark.parametrize("kind", ["uniform", "dynamic_range", "cancellation", "row_scales"])
def test_tensor_core_error_is_inside_half_the_band(model, d, kind):
    from knowledgegraphembedding_b200 import _lib
    nentity, nrel, gamma, nq = 2048 + 37, 7, 20.0, 160
    st = adversarial_tables(model, nentity, nrel, d, gamma, kind, seed=d % 97)
    m = make_model(model, nentity, nrel, d, gamma, st)
    K = m.entity_dim                       # contraction length: 1000, 2000, 4000, 1000
    rng = np.random.RandomState(3)
    test = [(int(rng.randint(nentity)), int(rng.randint(nrel)), int(rng.randint(nentity))) for _ in range(nq)]
    band = float(_lib.load().kge_eval_gemm_band(K))
    E = st["entity_embedding"].astype(np.float64)
    enorm = np.sqrt((E * E).sum(1))
    for mode in ("head-batch", "tail-batch"):
        _, exact = m.filtered_ranks(test, [], mode, return_scores=True)            # no filter: plain score matrix
        ranks_fast, approx = m.filtered_ranks(test, [], mode, return_approx=True)
        assert approx is not None, "tcgen05 path was not taken"
        exact, approx = exact.cpu().numpy().astype(np.float64), approx.cpu().numpy().astype(np.float64)
        # |q| from the identity  score(q, j) = <q, E_j>: recover q's norm through the kernel's own query vectors
        import ctypes
        from knowledgegraphembedding_b200.model import _ptr, _stream
        dev = m.entity_embedding.device
        qd = torch.tensor(test, dtype=torch.int64, device=dev)
        qvec = torch.empty(nq * K, device=dev)
        desc = m._descriptor()
        _lib.call("kge_eval_query_vectors", ctypes.byref(desc), _lib.MODE_IDS[mode], _ptr(qd), nq, _ptr(qvec), None, _stream(dev))
        qn = qvec.view(nq, K).double().norm(dim=1).cpu().numpy()
        scale = qn[:, None] * enorm[None, :]
        ok = scale > 0
        rel = np.abs(approx - exact)[ok] / scale[ok]
        assert np.isfinite(rel).all()
        assert rel.max() <= 0.5 * band, (model, K, kind, mode, rel.max(), band)
        ranks_exact = m.filtered_ranks(test, [], mode, exact=True)
        np.testing.assert_array_equal(ranks_fast, ranks_exact)
        print(f"{model} K={K} {kind} {mode}: max rel err {rel.max():.3e} = {rel.max() / band:.3f} x band({band:.3e})")
