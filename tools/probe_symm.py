"""Probe (run under torchrun on >= 2 GPUs): which peer-memory backends this box offers and what the exchange costs."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from knowledgegraphembedding_b200.peer import PeerExchange
    for backend in ("symm", "ipc"):
        os.environ["KGE_PEER_BACKEND"] = backend
        try:
            px = PeerExchange(dev, 1 << 20)
            if dist.get_rank() == 0:
                print(backend, "ok: multicast", hex(px.multicast), "ptrs", [hex(px.struct.grad[r] or 0) for r in range(px.world)],
                      flush=True)
            px.close()
        except Exception as exc:        # noqa: BLE001
            if dist.get_rank() == 0:
                print(backend, "failed:", exc, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
