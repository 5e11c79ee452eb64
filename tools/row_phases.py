#!/usr/bin/env python
"""Where do the cycles of the persistent row kernel go?  Runs cfg-3 train steps with KGE_ROW_PHASES=1 and prints the
per-phase cycle counters of thread 0 (csrc/kge_train_split.cuh), averaged per CTA and step.
    python tools/row_phases.py [--negatives N]"""
import ctypes
import os
import sys
import types

os.environ["KGE_ROW_PHASES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from knowledgegraphembedding_b200 import _lib  # noqa: E402


def main():
    wl = "rotate_fb15k"
    if "--negatives" in sys.argv:
        w = list(bench.WORKLOADS[wl])
        w[6] = int(sys.argv[sys.argv.index("--negatives") + 1])
        bench.WORKLOADS[wl] = tuple(w)
    model, nentity, nrel, d, gamma, B, N, lr, de, dr, reg, _ = bench.WORKLOADS[wl]
    b = bench.Bench(types.SimpleNamespace())
    m, opt, targs, _ = b.build_model(wl)
    dev = [(torch.from_numpy(p).cuda(), torch.from_numpy(n).cuda(), torch.from_numpy(w_).cuda(), md)
           for p, n, w_, md in bench.make_batches(nentity, nrel, B, N, 8, seed=1)]
    for i in range(5):
        m.train_step_async(opt, dev[i % 8], targs)
    torch.cuda.synchronize()
    out = (ctypes.c_uint64 * 8)()
    _lib.call("kge_debug_row_phase_cycles", out, 1)
    steps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m.train_step_async(opt, dev[i % 8], targs)
    e1.record()
    torch.cuda.synchronize()
    _lib.call("kge_debug_row_phase_cycles", out, 1)
    names = ["query vector", "candidate loop (thread 0's warp)", "row loss (incl. waiting for the slowest warp)", "fold",
             "chain rule", "positive triple", "  of the loop: waiting on the TMA mbarrier"]
    ctas = 148
    total = sum(out[i] for i in range(6))
    print(f"N={N}: {e0.elapsed_time(e1) / steps:.4f} ms per step; row kernel cycles per CTA and step (thread 0): "
          f"{total / ctas / steps:.0f} = {total / ctas / steps / 1.92e3:.1f} us at 1.92 GHz")
    for i, name in enumerate(names):
        print(f"  {name:50s} {out[i] / ctas / steps:12.0f} cycles  {100.0 * out[i] / total:5.1f} %")


if __name__ == "__main__":
    main()
