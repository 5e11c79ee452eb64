"""Host side of the NVLink peer-memory training exchange (csrc/kge_peer.cu, include/kge_b200.h `kge_peer_*`).

One `PeerExchange` per model and process group: it owns this rank's peer-visible block (flag block + gradient workspace),
exchanges the cudaIpc handles over torch.distributed, maps the other ranks' blocks and issues the fused
reduce-scatter + Adam + parameter-broadcast call.  The reference is single-device (run.py:241-242); this is the data
path of the batch-sharded multi-GPU extension.  torch.distributed is used for the one-time
handle exchange only -- there is no collective call per step.
"""
import ctypes
import socket

import torch

from . import _lib

FLAG_BYTES = 1024          # 2 channels x KGE_PEER_MAX_RANKS uint32, padded; the workspace starts 16-byte aligned after it


def param_slices(param_floats, world):
    """[(begin4, end4)] per rank: contiguous balanced slices of the parameter region in float4 units.  Rank r owns
    (and keeps the Adam moments of) workspace floats [4*begin4, 4*end4)."""
    units = int(param_floats) // 4
    base, rem = divmod(units, int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class _DeviceBlock:
    """Exposes a raw device allocation to torch (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange:
    def __init__(self, device, workspace_floats, group=None):
        import torch.distributed as dist
        self.device = device
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not (2 <= self.world <= _lib.PEER_MAX_RANKS):
            raise _lib.KgeError("peer exchange supports 2..%d ranks" % _lib.PEER_MAX_RANKS)
        self.capacity = int(workspace_floats)
        nbytes = FLAG_BYTES + 4 * self.capacity
        lib = _lib.load()
        self._local = ctypes.c_void_p()
        self._mapped = []
        self.epoch = 0
        ok, handle, why = 1, bytes(_lib.PEER_HANDLE_BYTES), ""
        try:
            _lib.call("kge_peer_alloc", device.index, nbytes, ctypes.byref(self._local))
            buf = ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
            _lib.call("kge_peer_export", self._local, buf)
            handle = buf.raw
        except Exception as exc:            # noqa: BLE001 -- any failure makes the whole group fall back together
            ok, why = 0, str(exc)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, socket.gethostname(), device.index, handle, why), group=group)
        same_node = len({g[1] for g in gathered}) == 1
        if not all(g[0] for g in gathered) or not same_node:
            self.close()
            raise _lib.KgeError("peer memory unavailable: " +
                                ("; ".join(g[4] for g in gathered if g[4]) or "ranks on different nodes"))
        ptrs, ok, why = [], 1, ""
        for r, g in enumerate(gathered):
            if r == self.rank:
                ptrs.append(self._local.value)
                continue
            p = ctypes.c_void_p()
            try:
                _lib.call("kge_peer_open", device.index, g[3], ctypes.byref(p))
                self._mapped.append(p)
                ptrs.append(p.value)
            except Exception as exc:        # noqa: BLE001
                ok, why = 0, str(exc)
                ptrs.append(0)
        oks = [None] * self.world
        dist.all_gather_object(oks, (ok, why), group=group)
        if not all(o[0] for o in oks):
            self.close()
            raise _lib.KgeError("peer memory unavailable: " + "; ".join(o[1] for o in oks if o[1]))
        self.struct = _lib.KgePeerGroup(world=self.world, rank=self.rank)
        for r, p in enumerate(ptrs):
            self.struct.flags[r] = p
            self.struct.grad[r] = p + FLAG_BYTES
        block = torch.as_tensor(_DeviceBlock(self._local.value, nbytes), device=device)
        self._block = block
        self.workspace = block[FLAG_BYTES:].view(torch.float32)        # [capacity] fp32, zero-initialised
        dist.barrier(group=group)            # nobody signals before every mapping exists

    def reduce_adam(self, entries, hyper, param_floats, row_offset, row_floats, rows_out, err, stream):
        """entries: [(param_ptr, grad_ptr, exp_avg_ptr, exp_avg_sq_ptr, numel, step, 0)], as for kge_adam_step."""
        self.epoch += 1
        lo4, hi4 = param_slices(param_floats, self.world)[self.rank]
        tensors = (_lib.KgeAdamTensor * len(entries))(*[_lib.KgeAdamTensor(*c) for c in entries])
        _lib.call("kge_peer_reduce_adam", ctypes.byref(self.struct), self.epoch, tensors, len(entries), param_floats,
                  lo4, hi4, row_offset, row_floats, ctypes.c_void_p(rows_out.data_ptr()), *hyper,
                  ctypes.c_void_p(err.data_ptr()), stream)

    def close(self):
        lib = _lib.load()
        for p in self._mapped:
            lib.kge_peer_close(p)
        self._mapped = []
        if self._local:
            lib.kge_peer_free(self._local)
            self._local = ctypes.c_void_p()


def gather_sliced_moments(tensors, offsets, param_floats, group=None):
    """After peer-exchange steps exp_avg / exp_avg_sq are current only on the rank owning each slice.  Broadcast every
    owned piece from its owner so that all ranks hold the full moments (before optimizer.state_dict(), or before
    switching to the replicated Adam).  tensors: [(exp_avg, exp_avg_sq)] per parameter, offsets: the parameter's first
    float inside the workspace."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    slices = param_slices(param_floats, world)
    for (m, v), off in zip(tensors, offsets):
        n = m.numel()
        fm, fv = m.view(-1), v.view(-1)
        for r, (lo4, hi4) in enumerate(slices):
            a, b = max(4 * lo4, off) - off, min(4 * hi4, off + n) - off
            if a < b:
                src = dist.get_global_rank(group, r) if group is not None else r
                dist.broadcast(fm[a:b], src=src, group=group)
                dist.broadcast(fv[a:b], src=src, group=group)


def owned_ranges(param_floats, world, rank, offset, numel):
    """[a, b) of a parameter's flat elements owned by `rank` (empty -> a >= b); numpy-free helper for tests."""
    lo4, hi4 = param_slices(param_floats, world)[rank]
    return max(4 * lo4, offset) - offset, min(4 * hi4, offset + numel) - offset


__all__ = ["PeerExchange", "param_slices", "gather_sliced_moments", "owned_ranges"]
