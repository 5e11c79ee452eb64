"""Worker of tests/test_multi_gpu.py (one process per GPU, launched with torch.distributed.run).

Checks on N >= 2 B200s, against the numpy oracle of the reference's single-device train step on the WHOLE batch:
  * batch-sharded training with the entity-sharded optimizer (kge_train_rows_sharded / kge_train_entity_sharded: the
    multi-GPU default) and through the dense NVLink peer-memory exchange (kge_peer_reduce_adam, KGE_PEER_DENSE=1):
    losses, tables and gathered Adam moments within 1e-5, replicas bit-identical on every rank,
    optimizer.state_dict() whole on every rank;
  * the same through the NCCL all-reduce path (KGE_NO_PEER behaviour), and that the two paths agree;
  * switching an optimizer from the peer path to the NCCL path mid-run (moments gathered automatically);
  * entity-sharded filtered ranking: ranks equal the oracle's;
  * at BASELINE.json's full RotatE FB15k shapes: peer path == NCCL path (tables, losses, gathered moments).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import outlier_fraction, relinf          # noqa: E402
from knowledgegraphembedding_b200 import KGEModel      # noqa: E402
from oracle import kge_oracle as O                     # noqa: E402

FLAGS = {"TransE": (False, False), "RotatE": (True, False), "pRotatE": (False, False), "ComplEx": (True, True),
         "DistMult": (False, False)}
TOL = 1e-5


def build(model, nentity, nrel, d, gamma, st, dev):
    de, dr = FLAGS[model]
    m = KGEModel(model, nentity, nrel, d, gamma, double_entity_embedding=de, double_relation_embedding=dr)
    with torch.no_grad():
        m.entity_embedding.copy_(torch.from_numpy(st["entity_embedding"]))
        m.relation_embedding.copy_(torch.from_numpy(st["relation_embedding"]))
        if model == "pRotatE":
            m.modulus.copy_(torch.from_numpy(st["modulus"]))
    return m.to(dev)


def batches(nentity, nrel, B, N, count, seed):
    """B: rows per batch, or a list of per-step sizes (a short last batch can leave some ranks without rows)."""
    rng = np.random.RandomState(seed)
    out = []
    sizes = list(B) if isinstance(B, (list, tuple)) else [B] * count
    for i in range(count):
        B = sizes[i]
        pos = np.stack([rng.randint(nentity, size=B), rng.randint(nrel, size=B), rng.randint(nentity, size=B)], 1)
        neg = rng.randint(nentity, size=(B, N))
        w = np.sqrt(1.0 / rng.randint(8, 200, size=B)).astype(np.float32)
        out.append((pos.astype(np.int64), neg.astype(np.int64), w, "tail-batch" if i % 2 == 0 else "head-batch"))
    return out


def as_torch(b):
    return (torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2]), b[3])


def identical_on_all_ranks(t, what):
    mine = t.detach().contiguous().view(torch.int32).to(torch.int64).sum().reshape(1)
    if dist.get_backend() == "gloo":
        mine = mine.cpu()
    got = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(got, mine)
    assert all(int(g) == int(got[0]) for g in got), what + ": replicas differ"


def check_case(model, nentity, nrel, d, gamma, B, N, steps, dev, lr=1e-3, adversarial=True, uni_weight=False, reg=0.0):
    rank, world = dist.get_rank(), dist.get_world_size()
    st = O.init_tables(model, nentity, nrel, d, gamma, *FLAGS[model], seed=3)
    pool = batches(nentity, nrel, B, N, steps, seed=7)
    sharded_expected = N >= 8 and nentity >= world and (d * (2 if FLAGS[model][0] else 1)) % 4 == 0 and \
        (d if FLAGS[model][0] else d) % 4 == 0
    args = types.SimpleNamespace(cuda=True, negative_adversarial_sampling=adversarial, adversarial_temperature=0.5,
                                 uni_weight=uni_weight, regularization=reg)
    ref = O.TrainState(model, st, gamma, d)
    ref_logs = [O.train_step(ref, b, lr=lr, adversarial=adversarial, alpha=0.5, uni_weight=uni_weight, regularization=reg)
                for b in pool]

    results = {}
    for path in ("sharded", "peer", "nccl", "switch"):
        m = build(model, nentity, nrel, d, gamma, st, dev)
        os.environ.pop("KGE_PEER_DENSE", None)
        if path == "peer":
            os.environ["KGE_PEER_DENSE"] = "1"
        if path == "nccl":
            m._ws['peer'] = False
        opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=lr)
        logs = []
        for i, b in enumerate(pool):
            if path == "switch" and i == steps // 2:
                m._ws['peer'] = False            # from here on the NCCL path: the sliced moments must be gathered first
                m._ws.pop('grad_views', None)
            logs.append(KGEModel.train_step(m, opt, iter([as_torch(b)]), args))
        if path == "peer":
            assert m._ws.get('peer') not in (None, False), "peer exchange was not active"
            if rank == 0:
                print("peer backend:", m._ws['peer'].backend, "multicast" if m._ws['peer'].multicast else "unicast", flush=True)
            assert getattr(opt, '_kge_sliced_moments', None) is not None
        if path == "sharded" and sharded_expected:
            assert isinstance(m._ws.get('shard'), dict), "entity-sharded optimizer was not active"
            assert m.entity_embedding.data_ptr() == m._ws['shard']['e_view'].data_ptr()
            assert getattr(opt, '_kge_sliced_moments', None) is not None and m.entity_embedding.grad is None
        sd = opt.state_dict()                    # run.py:106 -- gathers the sliced moments on the peer path
        assert getattr(opt, '_kge_sliced_moments', None) is None
        for log, want in zip(logs, ref_logs):
            for k, v in want.items():
                assert abs(log[k] - v) <= 2e-5 * max(abs(v), 1e-3), (model, path, k, log[k], v)
        names = ["entity_embedding", "relation_embedding"] + (["modulus"] if model == "pRotatE" else [])
        for idx, name in enumerate(names):
            got = getattr(m, name).detach().cpu().numpy()
            want = ref.state[name]
            assert outlier_fraction(got, want, TOL) <= 1e-3 and np.max(np.abs(got - want)) <= 2.5 * lr * steps, \
                (model, path, name, relinf(got, want))
            identical_on_all_ranks(getattr(m, name), f"{model}/{path}/{name}")
            mom = sd['state'][idx]
            # free-running steps: the kinked models (|x|, |sin x|) amplify summation-order noise from the second step on
            # (tests/test_gpu_fullshape.py::sync_state explains; same bound as the single-GPU fused-vs-dense test); the
            # entity-sharded pass sums an entity's pairs in the order its atomic cursor hands out
            mtol = 1e-4 if model in ("RotatE", "ComplEx", "DistMult") else 2e-3
            for key, okey in (("exp_avg", "m"), ("exp_avg_sq", "v")):
                g, w = mom[key].cpu().numpy(), ref.adam[name][okey]
                assert relinf(g, w) <= mtol, (model, path, name, key, relinf(g, w))
                identical_on_all_ranks(mom[key], f"{model}/{path}/{name}/{key}")
        results[path] = {n: getattr(m, n).detach().clone() for n in names}
    os.environ.pop("KGE_PEER_DENSE", None)
    for other in ("peer", "sharded"):
        for n in results[other]:
            a, b = results[other][n].cpu().numpy(), results["nccl"][n].cpu().numpy()
            assert outlier_fraction(a, b, TOL) <= 1e-3, (model, other, n)
    if rank == 0:
        print(f"train ok: {model} nentity={nentity} d={d} B={B} N={N} world={world}", flush=True)


def check_full_size(dev):
    """BASELINE.json configs[2] shapes (RotatE FB15k: 14,951 x 2000 table, 1024 rows per rank, 256 negatives): the
    peer-memory exchange and the NCCL all-reduce + replicated Adam both agree with the C oracle's single-device step on
    the WHOLE global batch (losses 1e-5; tables: outlier bound, Adam's sign-like first steps bounded by lr per step;
    gathered moments 1e-4), and with each other; replicas stay bit-identical."""
    rank, world = dist.get_rank(), dist.get_world_size()
    model, nentity, nrel, d, gamma, N, lr, steps = "RotatE", 14951, 1345, 1000, 24.0, 256, 1e-4, 3
    B = 1024 * world
    st = O.init_tables(model, nentity, nrel, d, gamma, True, False, seed=0)
    pool = batches(nentity, nrel, B, N, steps, seed=11)
    args = types.SimpleNamespace(cuda=True, negative_adversarial_sampling=True, adversarial_temperature=1.0,
                                 uni_weight=False, regularization=0.0)
    oracle = None
    if rank == 0:                                # one rank runs the CPU oracle (all host cores), the others wait
        from oracle import c_oracle as C
        C.set_num_threads(len(os.sched_getaffinity(0)))       # (torchrun exports OMP_NUM_THREADS=1)
        ts = C.TrainState(model, st, gamma, d)
        oracle = ([C.train_step(ts, b, lr=lr, adversarial=True, alpha=1.0) for b in pool], ts)
    dist.barrier()
    out = {}
    for path in ("sharded", "peer", "nccl"):
        m = build(model, nentity, nrel, d, gamma, st, dev)
        os.environ.pop("KGE_PEER_DENSE", None)
        if path == "peer":
            os.environ["KGE_PEER_DENSE"] = "1"
        if path == "nccl":
            m._ws['peer'] = False
        opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=lr)
        logs = [KGEModel.train_step(m, opt, iter([as_torch(b)]), args) for b in pool]
        if path == "sharded":
            assert isinstance(m._ws.get('shard'), dict), "entity-sharded optimizer was not active at full size"
        sd = opt.state_dict()
        for name in ("entity_embedding", "relation_embedding"):
            identical_on_all_ranks(getattr(m, name), f"full/{path}/{name}")
        out[path] = (logs, m.entity_embedding.detach().cpu().numpy(), m.relation_embedding.detach().cpu().numpy(),
                     sd['state'][0]['exp_avg'].cpu().numpy(), sd['state'][0]['exp_avg_sq'].cpu().numpy())
        if oracle is not None:
            ologs, ts = oracle
            for a, b in zip(logs, ologs):
                for k in b:
                    assert abs(a[k] - b[k]) <= 1e-5 * abs(b[k]), (path, k, a[k], b[k])
            for i, name in ((1, "entity_embedding"), (2, "relation_embedding")):
                got, want = out[path][i], ts.state[name]
                assert outlier_fraction(got, want, TOL) <= 1e-4 and np.max(np.abs(got - want)) <= 2.5 * lr * steps, (path, name)
            assert relinf(out[path][3], ts.m["entity_embedding"]) <= 1e-4, path
            assert relinf(out[path][4], ts.v["entity_embedding"]) <= 1e-4, path
        m._release_shard()
        del m, opt
        torch.cuda.empty_cache()
    os.environ.pop("KGE_PEER_DENSE", None)
    for other in ("peer", "sharded"):
        for a, b in zip(out[other][0], out["nccl"][0]):
            for k in a:
                assert abs(a[k] - b[k]) <= 1e-5 * abs(b[k]), (other, k, a[k], b[k])
        for i, what in ((1, "E"), (2, "R")):
            got, want = out[other][i], out["nccl"][i]
            assert outlier_fraction(got, want, TOL) <= 1e-4 and np.max(np.abs(got - want)) <= 2.5 * lr * steps, (other, what)
        assert relinf(out[other][3], out["nccl"][3]) <= 1e-4 and relinf(out[other][4], out["nccl"][4]) <= 1e-4, other
    if rank == 0:
        print(f"full-size ok: world={world}", flush=True)


def check_eval(dev):
    model, nentity, nrel, d, gamma = "RotatE", 3001, 5, 32, 12.0
    st = O.init_tables(model, nentity, nrel, d, gamma, True, False, seed=4)
    m = build(model, nentity, nrel, d, gamma, st, dev)
    rng = np.random.RandomState(0)
    all_true = sorted({(int(rng.randint(nentity)), int(rng.randint(nrel)), int(rng.randint(40))) for _ in range(2000)})
    test = all_true[:24]
    want = O.filtered_ranks(model, st, test, all_true, nentity, gamma, d)
    got = np.concatenate([m.filtered_ranks(test, all_true, mode, exact=True) for mode in ("head-batch", "tail-batch")])
    fast = np.concatenate([m.filtered_ranks(test, all_true, mode) for mode in ("head-batch", "tail-batch")])
    assert np.array_equal(got, fast)
    assert np.mean(got == np.asarray(want)) >= 0.95 and np.max(np.abs(got - np.asarray(want))) <= 1, (got, want)
    if dist.get_rank() == 0:
        print("eval ok", flush=True)


def main():
    # KGE_TEST_SAME_GPU=1: every rank uses cuda:0 (the driver's 1-GPU test box).  The ranks' contexts time-slice the GPU,
    # peer memory is cudaIpc between processes on one device, torch.distributed runs over gloo (NCCL refuses two ranks on
    # one device): the same kernels, barriers and host logic as on N GPUs, on a reduced list of cases.
    same = os.environ.get("KGE_TEST_SAME_GPU") == "1"
    local = 0 if same else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if same:
        dist.init_process_group("gloo")
        check_case("RotatE", 2003, 7, 64, 12.0, 64, 32, 3, dev)
        check_case("pRotatE", 517, 3, 12, 6.0, 33, 8, 3, dev)
        check_case("ComplEx", 1001, 7, 32, 20.0, 64, 32, 3, dev, reg=1e-3)
        check_case("TransE", 200, 3, 8, 6.0, 1, 8, 2, dev)
        check_case("RotatE", 301, 5, 16, 6.0, [48, 1, 33], 16, 3, dev)   # batch sizes change: a rank goes empty and back
        check_eval(dev)
        check_full_size(dev)                    # BASELINE configs[2] shape, 1024 rows per rank, against the C oracle
        dist.barrier()
        dist.destroy_process_group()
        if int(os.environ.get("RANK", "0")) == 0:
            print("ok", flush=True)
        return
    dist.init_process_group("nccl", device_id=dev)
    check_case("RotatE", 2003, 7, 64, 12.0, 64, 32, 4, dev)            # single-read path sizes (pairs/entity small)
    check_case("RotatE", 301, 5, 16, 6.0, 512, 64, 4, dev)             # many pairs per entity -> entity-major backward,
    #                                                                    exchange cut into regions overlapping it
    check_case("pRotatE", 517, 3, 10, 6.0, 33, 8, 3, dev)              # ragged: odd rows per rank, modulus, tensor tails
    check_case("TransE", 1000, 11, 50, 9.0, 16, 16, 3, dev, adversarial=False, uni_weight=True)   # model.py:274-275,281-283
    check_case("TransE", 200, 3, 8, 6.0, 1, 4, 2, dev)                 # fewer rows than ranks: some ranks hold no row
    check_case("ComplEx", 1001, 7, 32, 20.0, 64, 32, 3, dev, reg=1e-3)   # -r: L3 gradient inside the peer exchange
    # shapes the entity-sharded optimizer takes (rows of 16-byte multiples, >= 8 candidates): ragged rows per rank with
    # the modulus, uniform weights, fewer rows than ranks, -r on a real-valued model
    check_case("pRotatE", 517, 3, 12, 6.0, 33, 8, 3, dev)
    check_case("TransE", 1000, 11, 48, 9.0, 16, 16, 3, dev, adversarial=False, uni_weight=True)
    check_case("TransE", 200, 3, 8, 6.0, 1, 8, 2, dev)
    check_case("DistMult", 777, 5, 20, 10.0, 40, 24, 3, dev, reg=1e-3)
    check_case("RotatE", 301, 5, 16, 6.0, [48, 1, 33], 16, 3, dev)     # batch sizes change: ranks go empty and back
    check_eval(dev)
    check_full_size(dev)
    dist.barrier()
    dist.destroy_process_group()
    if int(os.environ.get("RANK", "0")) == 0:
        print("ok", flush=True)


if __name__ == "__main__":
    main()
