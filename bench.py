#!/usr/bin/env python
"""bench.py -- RotatE FB15k (BASELINE.json configs[2]: 14,951 entities, 1,345 relations, d=1000, -n 256 -b 1024
-g 24 -adv -de) train-step throughput in negative-sample scores/s, plus filtered-eval queries/s.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun by the driver)
    python bench.py --impl reference ...                     (CPU oracle port of the reference path, host cores)

One JSON line on stdout (rank 0).  A "step" is one full KGEModel.train_step: gather+score+loss+backward, dense
Adam, loss read-back excluded for `value` (inputs resident in HBM) and included for `e2e` (host batches).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, nentity, nrelation, d, gamma, B, N, lr, de, dr)
    "rotate_fb15k": ("RotatE", 14951, 1345, 1000, 24.0, 1024, 256, 1e-4, True, False),
    "transe_fb15k237": ("TransE", 14541, 237, 1000, 9.0, 1024, 256, 5e-5, False, False),
    "rotate_yago310": ("RotatE", 123182, 37, 500, 24.0, 1024, 400, 2e-4, True, False),
    "complex_wn18rr": ("ComplEx", 40943, 11, 500, 200.0, 512, 1024, 2e-3, True, True),
    "distmult_fb15k": ("DistMult", 14951, 1345, 2000, 500.0, 1024, 256, 1e-3, False, False),
}
METRIC = "rotate_negative_sample_scores_per_sec_train_step"


def make_batches(nentity, nrel, B, N, count, seed):
    rng = np.random.RandomState(seed)
    out = []
    for i in range(count):
        pos = np.stack([rng.randint(nentity, size=B), rng.randint(nrel, size=B), rng.randint(nentity, size=B)], 1)
        neg = rng.randint(nentity, size=(B, N))
        w = np.sqrt(1.0 / rng.randint(8, 200, size=B)).astype(np.float32)
        out.append((pos.astype(np.int64), neg.astype(np.int64), w, "tail-batch" if i % 2 == 0 else "head-batch"))
    return out


def train_bytes(B, N, De, Dr):
    """Algorithmic bytes of the negative-pass row kernel (DESIGN.md section 5 / SURVEY 8d)."""
    return B * N * De * 4 * 2 + B * N * 8 + 2 * B * (De + Dr) * 4 + B * 28


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def ncu_traffic(wl):
    """DRAM bytes per train-path launch from the committed `ncu --set full` capture (profiles/prof_train_r1l.raw.csv:
    dram__bytes_read.sum + dram__bytes_write.sum of row_kernel_split + entity_kernel); only for the captured workload."""
    path = os.path.join(ROOT, "profiles", "prof_train_r1l.raw.csv")
    if wl != "rotate_fb15k" or not os.path.exists(path):
        return None
    import csv
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    total = 0.0
    for r in rows[2:]:
        if "row_kernel_split" not in r[0] and "entity_kernel" not in r[0]:
            continue
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            total += float(r[i]) * scale
    return total


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------- CPU oracle arm
def cpu_train_sample(wl, rows, steps, warmup):
    """The reference's train_step restated by the oracle on the host cores, on the first `rows` positive rows of
    each batch at full width (N negatives, full tables incl. the dense Adam)."""
    from oracle import c_oracle as C
    from oracle import kge_oracle as O
    model, nentity, nrel, d, gamma, B, N, lr, de, dr = WORKLOADS[wl]
    rows = rows or B
    st = O.init_tables(model, nentity, nrel, d, gamma, de, dr, seed=0)
    batches = make_batches(nentity, nrel, rows, N, steps + warmup, seed=1)
    if hasattr(C, "train_step"):
        state = C.TrainState(model, st, gamma, d)
        step_fn = lambda b: C.train_step(state, b, lr=lr, adversarial=True, alpha=1.0)     # noqa: E731
        cores, kind = C.num_threads(), "port (oracle/kge_oracle.c, OpenMP)"
    else:
        state = O.TrainState(model, st, gamma, d)
        step_fn = lambda b: O.train_step(state, b, lr=lr, adversarial=True, alpha=1.0)     # noqa: E731
        cores, kind = 1, "port (oracle/kge_oracle.py, numpy)"
    for b in batches[:warmup]:
        step_fn(b)
    t0 = time.perf_counter()
    for b in batches[warmup:]:
        step_fn(b)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rows * N / dt, dt, cores, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    model, nentity, nrel, d, gamma, B, N, lr, de, dr = WORKLOADS[wl]
    rows = args.cpu_rows or B
    value, dt, cores, kind = cpu_train_sample(wl, rows, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scores/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "score_function": model, "nentity": nentity, "nrelation": nrel, "hidden_dim": d,
                   "negative_sample_size": N, "batch_size": B, "gamma": gamma, "adversarial": True,
                   "sample_rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": "scores/s", "cores": cores, "kind": "port",
                         "sample": f"{rows} of {B} positive rows x {N} negatives per step at full width, "
                                   f"full-table dense Adam included; {kind}"},
        "e2e": {"value": value, "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from knowledgegraphembedding_b200 import KGEModel
    from oracle import kge_oracle as O          # only for the portable synthetic table initialiser + cpu_baseline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload
    model, nentity, nrel, d, gamma, B, N, lr, de, dr = WORKLOADS[wl]
    Bg = B * world                                   # weak scaling: every rank keeps B rows of the global batch

    st = O.init_tables(model, nentity, nrel, d, gamma, de, dr, seed=0)
    m = KGEModel(model, nentity, nrel, d, gamma, double_entity_embedding=de, double_relation_embedding=dr)
    with torch.no_grad():
        m.entity_embedding.copy_(torch.from_numpy(st["entity_embedding"]))
        m.relation_embedding.copy_(torch.from_numpy(st["relation_embedding"]))
    m = m.to(dev)
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, m.parameters()), lr=lr)
    targs = types.SimpleNamespace(cuda=True, negative_adversarial_sampling=True, adversarial_temperature=1.0,
                                  uni_weight=False, regularization=0.0)
    pool = make_batches(nentity, nrel, Bg, N, 8, seed=1)            # same batches on every rank
    dev_pool = [(torch.from_numpy(p).to(dev), torch.from_numpy(n).to(dev), torch.from_numpy(w).to(dev), md)
                for p, n, w, md in pool]
    pin_pool = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(n).pin_memory(), torch.from_numpy(w).pin_memory(), md)
                for p, n, w, md in pool]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)               # max over ranks
        return float(ms.item())

    # ---- value: inputs resident in HBM, no host read-back inside the timed region --------------------------------
    m.train()

    def step_device(i):
        m.train_step_async(opt, dev_pool[i % len(dev_pool)], targs)

    sampler = ClockSampler(local)
    sampler.start()
    m._ws['kernel_events'] = None
    for i in range(args.warmup):
        step_device(i)
    m._ws['kernel_events'] = events = []
    m._ws['exchange_events'] = xevents = []
    ms_total = timed(step_device, args.steps, 0)
    m._ws['kernel_events'] = m._ws['exchange_events'] = None
    exchange_ms = float(np.mean([a.elapsed_time(b) for a, b in xevents])) if xevents else None
    px = m._ws.get('peer')
    nreg = int(m._ws.get('exchange_regions', 1)) if px else 1
    exchange_path = ("nvlink_peer_memory/%s%s (kge_peer_reduce_adam)" % (px.backend, "+nvswitch_multicast" if px.multicast else "")
                     if px else ("nccl_allreduce + kge_adam_step" if world > 1 else "kge_adam_step"))
    clocks = sampler.stop()
    row_ms = float(np.mean([a.elapsed_time(b) for a, b in events])) if events else None
    ms_per_step = ms_total / args.steps
    value = Bg * N / (ms_per_step * 1e-3)

    # ---- e2e: the public call (KGEModel.train_step) on pinned host batches, loss read back every step ----------
    it_state = {"i": 0}

    class HostIterator:
        def __next__(self):
            b = pin_pool[it_state["i"] % len(pin_pool)]
            it_state["i"] += 1
            return b

    host_it = HostIterator()
    last = {}

    def step_host(i):
        last.update(KGEModel.train_step(m, opt, host_it, targs))

    ms_e2e = timed(step_host, args.steps, args.warmup) / args.steps
    e2e_value = Bg * N / (ms_e2e * 1e-3)
    h2d = world * (B * 3 * 8 + B * N * 8 + Bg * 4)      # every rank copies its own rows (+ the whole weight vector)
    d2h = 8 * 4

    # ---- filtered evaluation throughput (entity-sharded over the ranks) ------------------------------------------
    rng = np.random.RandomState(2)
    all_true = sorted({(int(rng.randint(nentity)), int(rng.randint(nrel)), int(rng.randint(nentity)))
                       for _ in range(483142 if wl == "rotate_fb15k" else 200000)})
    nq = args.eval_queries
    test = [all_true[i] for i in rng.choice(len(all_true), nq, replace=False)]
    for mode in ("head-batch", "tail-batch"):                      # warm-up: filter index, workspaces, clocks
        m.filtered_ranks(test, all_true, mode)
    reps = 3
    m._ws['eval_events'] = eval_events = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for mode in ("head-batch", "tail-batch"):
            m.filtered_ranks(test, all_true, mode)                 # host triples in, host ranks out
    barrier()
    eval_s = (time.perf_counter() - t0) / reps
    m._ws['eval_events'] = None
    eval_qps = 2 * nq / eval_s
    eval_kernel_ms = sum(a.elapsed_time(b) for a, b in eval_events) / reps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # what bounds the distance-model ranking kernel: one square root per (query, entity, complex dimension) on the MUFU
    # pipe, 16 results per clock and SM (DESIGN.md section 3.3 / SURVEY 8d); peak at the SM clock sampled under load
    eval_roofline = None
    if model in ("RotatE",) and eval_kernel_ms > 0:
        sm_mhz = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        sqrt_per_s = 2 * nq * nentity * d / (eval_kernel_ms * 1e-3) / world      # per GPU: its entity slice
        peak_sqrt = 148 * 16 * sm_mhz * 1e6
        eval_roofline = {"bound": "mufu (1 sqrt per query x entity x complex dim, 16/clk/SM)",
                         "achieved": sqrt_per_s / 1e9, "peak": peak_sqrt / 1e9, "unit": "Gsqrt/s per GPU",
                         "frac": sqrt_per_s / peak_sqrt}
    peak, peak_src = peaks()
    De, Dr = m.entity_dim, m.relation_dim
    a_bytes = train_bytes(B, N, De, Dr)
    roofline = None
    if row_ms:
        ach = a_bytes / (row_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "kge_train_rows: row_kernel_split + counting sort + entity_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": ncu_traffic(wl), "peak_source": peak_src, "algorithmic_bytes_per_launch": a_bytes,
                    "avg_launch_ms": row_ms}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt, cores, kind = cpu_train_sample(wl, args.cpu_rows, 3, 1)
        cpu = {"value": v, "unit": "scores/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_rows or B} of {B} positive rows x {N} negatives per step at full width, full-table "
                         f"dense Adam included, 3 steps after 1 warm-up; {kind}"}
    line = {
        "metric": METRIC, "value": value, "unit": "scores/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "score_function": model, "nentity": nentity, "nrelation": nrel, "hidden_dim": d,
                   "negative_sample_size": N, "batch_size_per_gpu": B, "positives_per_step": Bg, "gamma": gamma,
                   "adversarial": True, "double_entity_embedding": de,
                   "sharding": f"positive rows over {world} rank(s), tables replicated; eval: entity slices",
                   "l2_policy": "no flush: each step streams 0.6 GB of tables+moments+grads (> 126 MB L2)"},
        # kernels of libkge_b200.so per step: weight_sum, row_kernel_split, scan_tiles, scan_apply, scatter_pairs,
        # entity_kernel per region, loss_finalize, and adam_kernel (1 GPU / NCCL path) or peer_reduce_adam + peer_finish
        # per region
        "clocks": clocks, "gpu_launches": (6 + nreg + (2 * nreg if px else 1)) * args.steps,
        "e2e": {"value": e2e_value, "unit": "scores/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "last_loss": last.get("loss")},
        "roofline": roofline, "cpu_baseline": cpu,
        "exchange": {"path": exchange_path, "exposed_exchange_and_optimizer_ms": exchange_ms,
                     "regions_per_step": nreg,
                     "bytes_reduced_per_rank": 4 * (m.entity_embedding.numel() + m.relation_embedding.numel())},
        "eval": {"metric": "filtered_eval_queries_per_sec", "value": eval_qps, "queries": 2 * nq, "seconds": eval_s,
                 "what": "KGEModel.filtered_ranks end to end: host triples -> CSR filter -> H2D -> kernels -> host ranks",
                 "count_kernel_ms": eval_kernel_ms, "count_kernel_queries_per_sec": 2 * nq / (eval_kernel_ms * 1e-3),
                 "one_table_pass_per_query_equiv_gbs": 2 * nq * nentity * De * 4 / (eval_kernel_ms * 1e-3) / 1e9 / world,
                 "sharding": f"entities/{world}", "filter_triples": len(all_true), "roofline": eval_roofline},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL's version banner, for one) print to stdout; the driver wants exactly one JSON line there.
    Route fd 1 to stderr for the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rotate_fb15k", choices=list(WORKLOADS))
    ap.add_argument("--cpu-rows", type=int, default=0, help="positive rows per CPU step (0 = the full batch)")
    ap.add_argument("--eval-queries", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()
