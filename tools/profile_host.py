#!/usr/bin/env python
"""Where does the host time of one KGEModel.train_step go?  (the `e2e` number pays it on every step: the log dict is
read back, so host work and GPU work do not overlap across steps)

  python tools/profile_host.py            # cfg-3 shape, pinned host batches; prints phase timings + cProfile top entries
"""
import cProfile
import os
import pstats
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from knowledgegraphembedding_b200 import KGEModel  # noqa: E402


def main():
    wl = "rotate_fb15k"
    model, nentity, nrel, d, gamma, B, N, lr, de, dr, reg, _ = bench.WORKLOADS[wl]
    b = bench.Bench(types.SimpleNamespace())
    m, opt, targs, _ = b.build_model(wl)
    pool = bench.make_batches(nentity, nrel, B, N, 8, seed=1)
    pin = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(n).pin_memory(), torch.from_numpy(w).pin_memory(), md)
           for p, n, w, md in pool]
    dev = [(p.cuda(), n.cuda(), w.cuda(), md) for p, n, w, md in pin]
    for i in range(10):
        KGEModel.train_step(m, opt, iter([pin[i % 8]]), targs)
    torch.cuda.synchronize()
    # (1) host-only cost of launching one step (device-resident inputs, no sync between steps)
    steps = 300
    t0 = time.perf_counter()
    for i in range(steps):
        m.train_step_async(opt, dev[i % 8], targs)
    host_launch = (time.perf_counter() - t0) / steps
    torch.cuda.synchronize()
    # (2) full e2e step
    t0 = time.perf_counter()
    for i in range(steps):
        KGEModel.train_step(m, opt, iter([pin[i % 8]]), targs)
    e2e = (time.perf_counter() - t0) / steps
    # (3) the pure H2D of one batch + a sync
    t0 = time.perf_counter()
    for i in range(steps):
        p, n, w, _ = pin[i % 8]
        p.cuda(non_blocking=True); n.cuda(non_blocking=True); w.cuda(non_blocking=True)
        torch.cuda.synchronize()
    h2d = (time.perf_counter() - t0) / steps
    # (4) an empty round trip: tiny kernel + .cpu()
    z = torch.zeros(8, device="cuda")
    t0 = time.perf_counter()
    for i in range(steps):
        z.cpu()
    rt = (time.perf_counter() - t0) / steps
    # (5) e2e in windows of 20 steps right after a device-resident burst (what bench.py's 20-step e2e region sees), with
    # and without an nvidia-smi poller running next to it (bench.py samples clocks during its timed regions)
    def windows(label, n=6):
        for i in range(20):
            m.train_step_async(opt, dev[i % 8], targs)
        torch.cuda.synchronize()
        out = []
        for w in range(n):
            t0 = time.perf_counter()
            for i in range(20):
                KGEModel.train_step(m, opt, iter([pin[i % 8]]), targs)
            out.append((time.perf_counter() - t0) / 20 * 1e6)
        print(f"e2e per 20-step window [{label}]: " + " ".join(f"{x:.0f}" for x in out) + " us")
    windows("no poller")
    # (6) the same with pinned batches that the device has never read before (first DMA from a freshly pinned region)
    fresh = [(torch.from_numpy(p_).pin_memory(), torch.from_numpy(n_).pin_memory(), torch.from_numpy(w_).pin_memory(), md)
             for p_, n_, w_, md in bench.make_batches(nentity, nrel, B, N, 40, seed=5)]
    torch.cuda.synchronize()
    for lo in (0, 20):
        t0 = time.perf_counter()
        for i in range(20):
            KGEModel.train_step(m, opt, iter([fresh[lo + i]]), targs)
        print(f"e2e, 20 steps over never-copied pinned batches: {(time.perf_counter() - t0) / 20 * 1e6:.0f} us")
    t0 = time.perf_counter()
    for i in range(40):
        KGEModel.train_step(m, opt, iter([fresh[i]]), targs)
    print(f"e2e, the same 40 batches again:                 {(time.perf_counter() - t0) / 40 * 1e6:.0f} us")
    sampler = bench.ClockSampler(0)
    sampler.start()
    windows("nvidia-smi -lms 100 running")
    sampler.stop()
    print(f"host time to enqueue one step (no sync): {host_launch * 1e6:.1f} us   [GPU-bound loop if > kernel time]")
    print(f"e2e step (pinned batch, log read back):     {e2e * 1e6:.1f} us")
    print(f"H2D of one batch + sync:                    {h2d * 1e6:.1f} us")
    print(f"device->host read of 32 B (idle GPU):       {rt * 1e6:.1f} us")
    pr = cProfile.Profile()
    pr.enable()
    for i in range(steps):
        KGEModel.train_step(m, opt, iter([pin[i % 8]]), targs)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
