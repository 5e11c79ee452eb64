#!/usr/bin/env python
"""Per-kernel CUDA times of the train step, single process or under torchrun (rank 0 reports): the breakdown ncu
cannot give for multi-rank runs (it replays kernels, and the cross-GPU barriers would time out).

    python tools/profile_step.py [--workload rotate_fb15k] [--steps 10]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/profile_step.py

Uses torch.profiler (CUPTI activity records: real kernel durations of the un-replayed run, overlapped kernels included)
around `steps` device-resident train steps after a warm-up.  Output: one JSON line {kernel: [launches/step, avg us]}.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rotate_fb15k", choices=list(B.WORKLOADS))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    b = B.Bench(args)
    wl = args.workload
    model, nentity, nrel, d, gamma, rows, N, lr, de, dr, reg, _ = B.WORKLOADS[wl]
    m, opt, targs, _ = b.build_model(wl)
    pool = B.make_batches(nentity, nrel, rows * b.world, N, 8, seed=1)
    dev_pool = [(torch.from_numpy(p).to(b.dev), torch.from_numpy(n).to(b.dev), torch.from_numpy(w).to(b.dev), md)
                for p, n, w, md in pool]
    m.train()
    for i in range(args.warmup):
        m.train_step_async(opt, dev_pool[i % 8], targs)
    b.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            m.train_step_async(opt, dev_pool[i % 8], targs)
        b.barrier()
    if b.rank == 0:
        out, total = {}, 0.0
        for ev in prof.key_averages():
            t = getattr(ev, "device_time_total", None)
            if t is None:
                t = getattr(ev, "cuda_time_total", 0.0)
            if t <= 0:
                continue
            name = ev.key.split("(")[0].replace("void kge::", "").replace("kge::", "")[:70]
            out[name] = [round(ev.count / args.steps, 2), round(t / ev.count, 2)]
            total += t / args.steps
        print(json.dumps({"world": b.world, "workload": wl, "sum_kernel_us_per_step": round(total, 1),
                          "kernels": dict(sorted(out.items(), key=lambda kv: -kv[1][0] * kv[1][1]))}))
    if b.world > 1:
        b.dist.destroy_process_group()


if __name__ == "__main__":
    main()
