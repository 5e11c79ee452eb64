"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/kge_b200.h declares,
argument errors map to the reference's exceptions, shard arithmetic, and the gloo world_size-2 data path."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from knowledgegraphembedding_b200 import _lib
    header = open(os.path.join(ROOT, "include", "kge_b200.h")).read()
    declared = set(re.findall(r"\b(kge_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.kge_abi_version() == 2


def test_abi_argument_errors_without_a_gpu():
    from knowledgegraphembedding_b200 import _lib
    lib = _lib.load()
    desc = _lib.KgeModelStruct(model=7, device=0, nentity=4, nrelation=2, hidden_dim=4, entity_dim=4, relation_dim=4,
                               gamma=1.0, embedding_range=0.1, entity=8, relation=8, modulus=None)
    rc = lib.kge_score_forward(ctypes.byref(desc), 0, None, None, 1, 1, None, None, None)
    assert rc == _lib.ERR_INVALID and b"model 7 not supported" in lib.kge_last_error()
    desc.model = _lib.ROTATE
    rc = lib.kge_score_forward(ctypes.byref(desc), 0, None, None, 1, 1, None, None, None)
    assert rc == _lib.ERR_INVALID and b"RotatE should use --double_entity_embedding" in lib.kge_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_model_shell_matches_reference_surface():
    from knowledgegraphembedding_b200 import KGEModel
    m = KGEModel(model_name="RotatE", nentity=7, nrelation=3, hidden_dim=4, gamma=12.0, double_entity_embedding=True)
    assert [n for n, _ in m.named_parameters()] == ["gamma", "embedding_range", "entity_embedding", "relation_embedding"]
    assert m.entity_embedding.shape == (7, 8) and m.relation_embedding.shape == (7 - 4, 4)
    assert not m.gamma.requires_grad and not m.embedding_range.requires_grad
    rho = (12.0 + 2.0) / 4
    assert abs(m.embedding_range.item() - rho) < 1e-6 and m.entity_embedding.abs().max().item() <= rho
    assert callable(m.train_step) and callable(m.test_step)
    for name in ("TransE", "DistMult", "ComplEx", "RotatE", "pRotatE"):
        assert callable(getattr(m, name))
    p = KGEModel("pRotatE", 7, 3, 4, 12.0)
    assert p.modulus.shape == (1, 1) and abs(p.modulus.item() - 0.5 * rho) < 1e-6
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, dtype=torch.long))


def test_dropin_module_name():
    code = ("import sys; sys.path.insert(0, %r); import model; "
            "print(model.KGEModel.__module__)") % os.path.join(ROOT, "knowledgegraphembedding_b200", "dropin")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert out.strip() == "knowledgegraphembedding_b200.model"


def test_shard_bounds_cover_exactly():
    from knowledgegraphembedding_b200 import shard_bounds
    for total in (0, 1, 7, 1024, 14951, 123182):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from knowledgegraphembedding_b200 import shard_bounds
from knowledgegraphembedding_b200.model import _dist
from oracle import kge_oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = _dist()
assert (rank, world) == (int(sys.argv[1]), 2)
# batch-sharded training: per-rank oracle gradients of the row slice, all-reduced, equal the full-batch gradients
nentity, nrel, d, gamma, B, N = 60, 4, 8, 6.0, 10, 6
st = O.init_tables("RotatE", nentity, nrel, d, gamma, True, False, seed=0)
rng = np.random.RandomState(0)
pos = np.stack([rng.randint(nentity, size=B), rng.randint(nrel, size=B), rng.randint(nentity, size=B)], 1)
neg = rng.randint(nentity, size=(B, N)); w = (rng.rand(B) + 0.1).astype(np.float32)
neg_s = O.forward("RotatE", st, (pos, neg), "tail-batch", gamma, d); pos_s = O.forward("RotatE", st, pos, "single", gamma, d)
_, _, _, dneg, dpos = O.loss_and_dscore(neg_s, pos_s, w, True, 1.0, False)
b, e = shard_bounds(B, rank, world)
g = O.score_backward("RotatE", st, (pos[b:e], neg[b:e]), "tail-batch", dneg[b:e], gamma, d)
g2 = O.score_backward("RotatE", st, pos[b:e], "single", dpos[b:e], gamma, d)
flat = torch.from_numpy(np.concatenate([(g[k] + g2[k]).ravel() for k in ("entity_embedding", "relation_embedding")]))
dist.all_reduce(flat)
full = O.score_backward("RotatE", st, (pos, neg), "tail-batch", dneg, gamma, d)
full2 = O.score_backward("RotatE", st, pos, "single", dpos, gamma, d)
want = np.concatenate([(full[k] + full2[k]).ravel() for k in ("entity_embedding", "relation_embedding")])
assert np.allclose(flat.numpy(), want, rtol=1e-9, atol=1e-12)
# entity-sharded evaluation: integer counts over the two entity slices add up to rank - 1
all_true = sorted({(int(rng.randint(nentity)), int(rng.randint(nrel)), int(rng.randint(nentity))) for _ in range(80)})
test = all_true[:6]
ranks, scores = O.filtered_ranks("RotatE", st, test, all_true, nentity, gamma, d, return_scores=True)
eb, ee = shard_bounds(nentity, rank, world)
cols = [t[0] for t in test] + [t[2] for t in test]
cnt = torch.tensor([int(sum((row[j] > row[p]) or (row[j] == row[p] and j < p) for j in range(eb, ee)))
                    for row, p in zip(scores, cols)], dtype=torch.int32)
dist.all_reduce(cnt)
assert np.array_equal(cnt.numpy() + 1, ranks)
# host-side row sharding of a train batch: the two ranks' shards tile the batch, weights stay whole
from knowledgegraphembedding_b200.model import _shard_rows
batch = (torch.from_numpy(pos), torch.from_numpy(neg), torch.from_numpy(w), "tail-batch")
sh = _shard_rows(batch)
assert sh.total == B and sh.row_begin == b and sh.positive.shape[0] == e - b and sh.weight.shape[0] == B
assert _shard_rows(sh) is sh
got = [None, None]
dist.all_gather_object(got, (sh.row_begin, sh.positive.numpy(), sh.negative.numpy()))
assert np.array_equal(np.concatenate([g[1] for g in got]), pos) and np.array_equal(np.concatenate([g[2] for g in got]), neg)
# peer-exchange bookkeeping: moments that are current only on the owning rank become whole on every rank
from knowledgegraphembedding_b200.peer import exchange_regions, gather_sliced_moments, region_slices
nE, nR = 52 * 12, 5 * 8                                    # [dE | dR(pad to 4) | dM(4)] like KGEModel._grad_workspace
param_floats = nE + nR + 4
regions, _ = exchange_regions(param_floats, 52, 12, 3)
offsets = [0, nE, nE + nR]
truth = [torch.arange(n, dtype=torch.float32) + 1000 * k for k, n in enumerate((nE, nR, 1))]
pairs = []
for t, off in zip(truth, offsets):
    m, v = torch.full_like(t, -1.0), torch.full_like(t, -2.0)
    for region in regions:
        lo4, hi4 = region_slices(region, world)[rank]
        a, b_ = max(4 * lo4, off) - off, min(4 * hi4, off + t.numel()) - off
        if a < b_:
            m[a:b_] = t[a:b_]; v[a:b_] = 2 * t[a:b_]
    pairs.append((m, v))
gather_sliced_moments(pairs, offsets, regions)
for (m, v), t in zip(pairs, truth):
    assert torch.equal(m, t) and torch.equal(v, 2 * t)
# the peer-exchange scheme itself (csrc/kge_peer.cu), restated over gloo: every rank owns one slice per region of the
# parameter part, sums the ranks' partial gradients of that slice in rank order, runs Adam there with its own moments,
# and the new parameters are broadcast; after two steps every replica equals the oracle's full-batch train steps.
def peer_step(params, moments, partial, step, lr, regions):
    flat_g = torch.from_numpy(np.concatenate([partial[k].ravel() for k in ("entity_embedding", "relation_embedding")]))
    gathered = [torch.zeros_like(flat_g) for _ in range(world)]
    dist.all_gather(gathered, flat_g)                       # stand-in for the NVLink loads of the peers' workspaces
    flat_p = np.concatenate([params[k].ravel() for k in ("entity_embedding", "relation_embedding")])
    new_p = torch.from_numpy(flat_p.copy())
    for region in regions:
        lo4, hi4 = region_slices(region, world)[rank]
        a, b_ = 4 * lo4, min(4 * hi4, flat_p.size)
        if a < b_:
            g = gathered[0].numpy()[a:b_].copy()
            for r in range(1, world):
                g = g + gathered[r].numpy()[a:b_]
            O.adam_update(flat_p[a:b_], g, moments[0][a:b_], moments[1][a:b_], step, lr)
            new_p[a:b_] = torch.from_numpy(flat_p[a:b_])
        for r in range(world):                              # the owner's slice reaches every replica
            l4, h4 = region_slices(region, world)[r]
            x, y = 4 * l4, min(4 * h4, flat_p.size)
            if x < y:
                dist.broadcast(new_p[x:y], src=r)
    nE_ = params["entity_embedding"].size
    params["entity_embedding"][...] = new_p.numpy()[:nE_].reshape(params["entity_embedding"].shape)
    params["relation_embedding"][...] = new_p.numpy()[nE_:].reshape(params["relation_embedding"].shape)

nentity, nrel, d, gamma, B, N, lr = 60, 4, 8, 6.0, 10, 6, 1e-2
st = O.init_tables("RotatE", nentity, nrel, d, gamma, True, False, seed=0)
ref = O.TrainState("RotatE", st, gamma, d)
mine = {k: v.copy() for k, v in st.items()}
total = mine["entity_embedding"].size + mine["relation_embedding"].size          # both multiples of 4 here
regions, _ = exchange_regions(total, nentity, 2 * d, 3)
moments = (np.zeros(total, np.float32), np.zeros(total, np.float32))
for step in (1, 2):
    rng2 = np.random.RandomState(100 + step)
    pos = np.stack([rng2.randint(nentity, size=B), rng2.randint(nrel, size=B), rng2.randint(nentity, size=B)], 1)
    neg = rng2.randint(nentity, size=(B, N)); w = (rng2.rand(B) + 0.1).astype(np.float32)
    O.train_step(ref, (pos, neg, w, "tail-batch"), lr=lr, adversarial=True, alpha=1.0)
    neg_s = O.forward("RotatE", mine, (pos, neg), "tail-batch", gamma, d); pos_s = O.forward("RotatE", mine, pos, "single", gamma, d)
    _, _, _, dneg, dpos = O.loss_and_dscore(neg_s, pos_s, w, True, 1.0, False)
    b, e = shard_bounds(B, rank, world)
    g1 = O.score_backward("RotatE", mine, (pos[b:e], neg[b:e]), "tail-batch", dneg[b:e], gamma, d)
    g2 = O.score_backward("RotatE", mine, pos[b:e], "single", dpos[b:e], gamma, d)
    partial = {k: (g1[k] + g2[k]).astype(np.float32) for k in g1}
    peer_step(mine, moments, partial, step, lr, regions)
for k in ("entity_embedding", "relation_embedding"):
    assert np.max(np.abs(mine[k] - ref.state[k])) <= 1e-5 * np.max(np.abs(ref.state[k])) + 2 * lr * 1e-3, k
    mineT = torch.from_numpy(mine[k].copy()); other = mineT.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mineT, other), "replicas differ"                      # bit-identical on every rank
# the entity-sharded optimizer (kge_train_rows_sharded / kge_train_entity_sharded), restated over gloo at the level of
# its ownership rules: rank r alone sums the gradient rows of ITS entity range over all ranks' row shards, runs Adam there
# with moments nobody else keeps, and its rows reach every replica; the relation table goes through the region scheme.
from knowledgegraphembedding_b200.peer import entity_ranges, gather_moment_ranges, moment_ranges
mine = {k: v.copy() for k, v in st.items()}
ref = O.TrainState("RotatE", st, gamma, d)
De = 2 * d
own = entity_ranges(nentity, De, world)
nE_, nR_ = mine["entity_embedding"].size, mine["relation_embedding"].size
small = moment_ranges([0], [nR_], [(0, nR_ // 4)], world)
mE, vE = np.zeros(nE_, np.float32), np.zeros(nE_, np.float32)
mR, vR = np.zeros(nR_, np.float32), np.zeros(nR_, np.float32)
for step in (1, 2):
    rng2 = np.random.RandomState(200 + step)
    pos = np.stack([rng2.randint(nentity, size=B), rng2.randint(nrel, size=B), rng2.randint(nentity, size=B)], 1)
    neg = rng2.randint(nentity, size=(B, N)); w = (rng2.rand(B) + 0.1).astype(np.float32)
    O.train_step(ref, (pos, neg, w, "head-batch"), lr=lr, adversarial=True, alpha=1.0)
    neg_s = O.forward("RotatE", mine, (pos, neg), "head-batch", gamma, d); pos_s = O.forward("RotatE", mine, pos, "single", gamma, d)
    _, _, _, dneg, dpos = O.loss_and_dscore(neg_s, pos_s, w, True, 1.0, False)
    b, e = shard_bounds(B, rank, world)
    g1 = O.score_backward("RotatE", mine, (pos[b:e], neg[b:e]), "head-batch", dneg[b:e], gamma, d)
    g2 = O.score_backward("RotatE", mine, pos[b:e], "single", dpos[b:e], gamma, d)
    for name, ranges, mm, vv in (("entity_embedding", own, mE, vE), ("relation_embedding", small[0], mR, vR)):
        part = torch.from_numpy((g1[name] + g2[name]).astype(np.float32).ravel())
        gathered = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)                     # stand-in for the rows the owner reads from its gather area
        flat = mine[name].reshape(-1)
        new = torch.from_numpy(flat.copy())
        for r, a, b_ in ranges:
            if r == rank:
                g = gathered[0].numpy()[a:b_].copy()
                for q in range(1, world):
                    g = g + gathered[q].numpy()[a:b_]
                O.adam_update(flat[a:b_], g, mm[a:b_], vv[a:b_], step, lr)
                new[a:b_] = torch.from_numpy(flat[a:b_])
        for r, a, b_ in ranges:
            dist.broadcast(new[a:b_], src=r)                # the owner's NVLink stores into every replica
        flat[...] = new.numpy()
for k in ("entity_embedding", "relation_embedding"):
    assert np.max(np.abs(mine[k] - ref.state[k])) <= 1e-5 * np.max(np.abs(ref.state[k])) + 2 * lr * 1e-3, k
pairs = [(torch.from_numpy(mE), torch.from_numpy(vE)), (torch.from_numpy(mR), torch.from_numpy(vR))]
gather_moment_ranges(pairs, (own,) + small)
assert np.max(np.abs(mE - ref.adam["entity_embedding"]["m"].ravel())) <= 1e-6
assert np.max(np.abs(vR - ref.adam["relation_embedding"]["v"].ravel())) <= 1e-6
dist.destroy_process_group()
print("ok")
"""


def test_world_size_2_gloo_data_path(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0 and out.strip().endswith("ok"), err[-2000:]


def test_peer_exchange_region_arithmetic():
    from knowledgegraphembedding_b200.peer import exchange_regions, region_slices
    for nentity, De, nrel_floats, nslices in ((14951, 2000, 1345 * 1000, 1), (14951, 2000, 1345 * 1000, 3),
                                              (301, 32, 5 * 16, 4), (7, 8, 12, 3), (10, 6, 8, 2)):
        nE4 = (nentity * De + 3) // 4 * 4
        param_floats = nE4 + (nrel_floats + 3) // 4 * 4 + 4
        regions, entities = exchange_regions(param_floats, nentity, De, nslices)
        assert regions[0][0] == 0 and regions[-1][1] == param_floats // 4
        assert all(a[1] == b[0] for a, b in zip(regions, regions[1:]))          # regions tile the parameter part
        assert entities[0][0] == 0 and entities[-1][1] == nentity
        assert all(a[1] == b[0] for a, b in zip(entities, entities[1:]))
        if De % 4:
            assert len(regions) == 1                                           # rows are not float4 multiples: no slicing
        for (lo4, hi4), (eb, ee) in list(zip(regions, entities))[:-1]:
            assert hi4 * 4 == ee * De                                          # a region ends with its entity range
        for world in (2, 3, 8):
            for region in regions:
                sl = region_slices(region, world)
                assert sl[0][0] == region[0] and sl[-1][1] == region[1]
                assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
                sizes = [b - a for a, b in sl]
                assert max(sizes) - min(sizes) <= 1


def test_entity_ownership_matches_the_kernels_rule():
    """peer.entity_ranges (host: who keeps which Adam moments) against the device rule of Mirror / owner_of
    (csrc/kge_train_args.cuh): base = nentity / G, the first nentity % G ranks own one row more."""
    from knowledgegraphembedding_b200.peer import entity_ranges
    from knowledgegraphembedding_b200.model import shard_bounds

    def owner_of(i, base, rem):                            # restatement of the device function
        cut = rem * (base + 1)
        return i // (base + 1) if i < cut else rem + (i - cut) // base

    for nentity, world, De in ((14951, 8, 2000), (14951, 2, 2000), (123182, 8, 1000), (40943, 4, 1000), (9, 8, 4), (8, 8, 4)):
        base, rem = divmod(nentity, world)
        rs = entity_ranges(nentity, De, world)
        assert rs[0][1] == 0 and rs[-1][2] == nentity * De and all(a[2] == b[1] for a, b in zip(rs, rs[1:]))
        for r, a, b in rs:
            assert (a // De, b // De) == shard_bounds(nentity, r, world)
            for i in {a // De, b // De - 1, (a // De + b // De) // 2}:
                assert owner_of(i, base, rem) == r


def test_shard_abi_argument_errors_without_a_gpu():
    import ctypes
    from knowledgegraphembedding_b200 import _lib
    lib = _lib.load()
    m = _lib.KgeModelStruct(model=_lib.ROTATE, device=0, nentity=100, nrelation=5, hidden_dim=8, entity_dim=16,
                            relation_dim=8, gamma=6.0, embedding_range=1.0, entity=4096, relation=8192, modulus=None)
    assert lib.kge_train_gather_bytes(ctypes.byref(m), 2, 8, 16) >= 2 * 8 * (16 * 4 * 2 + 16 * 4 * 4 + 12)
    assert lib.kge_train_shard_workspace_bytes(ctypes.byref(m), 2, 8, 16) >= 2 * 8 * 19 * 8
    sh = _lib.KgeShard(world=1, rank=0)
    rc = lib.kge_train_rows_sharded(ctypes.byref(m), _lib.TAIL_BATCH, 0, 1.0, None, None, None, None, 8, 4, 16, None, None,
                                    None, None, ctypes.byref(sh), None, None, None)
    assert rc == _lib.ERR_INVALID and b"bad shard description" in lib.kge_last_error()
    sh = _lib.KgeShard(world=2, rank=0, block_bytes=1 << 20, gather_offset=256, rows_max=4)
    sh.block[0], sh.block[1] = 1 << 30, 1 << 31
    rc = lib.kge_train_rows_sharded(ctypes.byref(m), _lib.TAIL_BATCH, 0, 1.0, None, None, None, None, 8, 4, 16, None, None,
                                    None, None, ctypes.byref(sh), None, None, None)
    assert rc == _lib.ERR_INVALID and b"entity table must live inside the local peer block" in lib.kge_last_error()
    grp = _lib.KgePeerGroup(world=2, rank=0)
    rc = lib.kge_peer_barrier(ctypes.byref(grp), 0, 1, 0, 0, None, None)
    assert rc == _lib.ERR_INVALID and b"channels 2 and 3" in lib.kge_last_error()


def test_filter_index_build_argument_errors_without_a_gpu():
    from knowledgegraphembedding_b200 import _lib
    lib = _lib.load()
    assert lib.kge_eval_filter_index_scratch_bytes(100, 7) >= 700 * 4 + 4
    rc = lib.kge_eval_filter_index_build(None, 0, _lib.SINGLE, 100, 7, 16, 16, 16, 1 << 20, None, None)
    assert rc == _lib.ERR_INVALID and b"negative batch mode 0 not supported" in lib.kge_last_error()
    rc = lib.kge_eval_filter_index_build(None, 0, _lib.HEAD_BATCH, 1 << 20, 1 << 12, 16, 16, 16, 1 << 20, None, None)
    assert rc == _lib.ERR_INVALID and b"too large for the direct-address index" in lib.kge_last_error()
    rc = lib.kge_eval_filter_index_build(None, 0, _lib.HEAD_BATCH, 100, 7, 16, 16, 16, 8, None, None)
    assert rc == _lib.ERR_INVALID and b"scratch too small" in lib.kge_last_error()


def test_peer_abi_argument_errors_without_a_gpu():
    import ctypes
    from knowledgegraphembedding_b200 import _lib
    lib = _lib.load()
    grp = _lib.KgePeerGroup(world=1, rank=0)
    t = (_lib.KgeAdamTensor * 1)(_lib.KgeAdamTensor(16, 16, 16, 16, 4, 1, 0))
    rc = lib.kge_peer_reduce_adam(ctypes.byref(grp), 1, t, 1, 8, 0, 2, 0, 1, 8, 0, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, None, None)
    assert rc == _lib.ERR_INVALID and b"peer group of 1 ranks not supported" in lib.kge_last_error()
    grp = _lib.KgePeerGroup(world=2, rank=0)
    rc = lib.kge_peer_reduce_adam(ctypes.byref(grp), 1, t, 1, 6, 0, 1, 0, 1, 8, 0, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, None, None)
    assert rc == _lib.ERR_INVALID and b"multiple of 4" in lib.kge_last_error()
    rc = lib.kge_peer_reduce_adam(ctypes.byref(grp), 1, t, 1, 8, 0, 2, 1, 3, 8, 0, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, None, None)
    assert rc == _lib.ERR_INVALID and b"bad region / slice" in lib.kge_last_error()
    rc = lib.kge_peer_reduce_adam(ctypes.byref(grp), 1, t, 1, 8, 0, 2, 0, 1, 8, 0, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, None, None)
    assert rc == _lib.ERR_INVALID and b"peer 0 is not mapped" in lib.kge_last_error()


def test_metrics_from_ranks_bit_identical_to_the_reference_loop():
    """model.py:412-427 builds one dict per query and averages with sum()/len(); ours must give the same bits."""
    from knowledgegraphembedding_b200.model import metrics_from_ranks
    rng = np.random.RandomState(3)
    for n in (1, 7, 2 * 3134, 100001):
        ranks = rng.randint(1, 15000, size=n)
        ranks[rng.rand(n) < 0.2] = rng.randint(1, 12, size=int((rng.rand(n) < 0.2).sum()) or 1)[0]
        logs = []
        for ranking in ranks.tolist():                       # verbatim shape of the reference's loop body
            logs.append({'MRR': 1.0 / ranking, 'MR': float(ranking), 'HITS@1': 1.0 if ranking <= 1 else 0.0,
                         'HITS@3': 1.0 if ranking <= 3 else 0.0, 'HITS@10': 1.0 if ranking <= 10 else 0.0})
        want = {m: sum([log[m] for log in logs]) / len(logs) for m in logs[0].keys()}
        got = metrics_from_ranks(ranks)
        assert list(got.keys()) == list(want.keys())
        for m in want:
            assert got[m] == want[m], (n, m, got[m], want[m])
