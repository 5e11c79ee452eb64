"""Host side of the NVLink peer-memory training exchange (csrc/kge_peer.cu, include/kge_b200.h `kge_peer_*`).

One `PeerExchange` per model and process group: it owns this rank's peer-visible block (flag block + gradient workspace),
exchanges the cudaIpc handles over torch.distributed, maps the other ranks' blocks and issues the fused
reduce-scatter + Adam + parameter-broadcast call.  The reference is single-device (run.py:241-242); this is the data
path of the batch-sharded multi-GPU extension.  torch.distributed is used for the one-time
handle exchange only -- there is no collective call per step.
"""
import ctypes
import os
import socket

import torch

from . import _lib

FLAG_BYTES = 1024          # 8 channels x KGE_PEER_MAX_RANKS uint32 (0/1: kge_peer_reduce_adam, 2/3: kge_peer_barrier, 6/7: its
#                            error words), padded; the workspace starts 16-byte aligned after it


def _align256(n):
    return (int(n) + 255) // 256 * 256


def region_slices(region, world):
    """[(begin4, end4)] per rank: contiguous balanced slices of one exchange region (float4 units).  Rank r owns (and
    keeps the Adam moments of) workspace floats [4*begin4, 4*end4)."""
    rlo, rhi = int(region[0]), int(region[1])
    base, rem = divmod(rhi - rlo, int(world))
    out, lo = [], rlo
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def exchange_regions(param_floats, nentity, entity_dim, nslices):
    """Cut the parameter part of the workspace [dE | dR | dM] into `nslices` regions (float4 units) along entity
    ranges, so that the exchange of a finished slice of dE overlaps the entity-major backward of the next slice; the
    last region also takes dR and dM (final only when every entity pass has run)."""
    total4 = int(param_floats) // 4
    if nslices <= 1 or entity_dim % 4:
        return [(0, total4)], [(0, int(nentity))]
    regions, entities, lo4 = [], [], 0
    for k in range(nslices):
        base, rem = divmod(int(nentity), nslices)
        eb = k * base + min(k, rem)
        ee = eb + base + (1 if k < rem else 0)
        hi4 = total4 if k == nslices - 1 else ee * entity_dim // 4
        regions.append((lo4, hi4))
        entities.append((eb, ee))
        lo4 = hi4
    return regions, entities


class _DeviceBlock:
    """Exposes a raw device allocation to torch (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange:
    """backend 'symm': the block comes from torch's symmetric-memory allocator (cuMem + an NVSwitch multicast mapping
    when the fabric supports it -> multimem.ld_reduce / multimem.st in the kernel); backend 'ipc': cudaMalloc + cudaIpc
    handles through kge_peer_alloc/export/open (unicast NVLink loads and stores).  KGE_PEER_BACKEND=symm|ipc forces one,
    KGE_PEER_MULTICAST=0|1 overrides whether the multicast mapping is used (default: only above 4 ranks)."""

    def __init__(self, device, workspace_floats, group=None, extra_bytes=0):
        import torch.distributed as dist
        self.device = device
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not (2 <= self.world <= _lib.PEER_MAX_RANKS):
            raise _lib.KgeError("peer exchange supports 2..%d ranks" % _lib.PEER_MAX_RANKS)
        self.capacity = int(workspace_floats)
        # [flags | workspace | extra]: `extra` (entity-sharded step) holds the entity table and the gather area
        self.extra_offset = _align256(FLAG_BYTES + 4 * self.capacity)
        self.extra_bytes = int(extra_bytes)
        nbytes = self.extra_offset + self.extra_bytes if extra_bytes else FLAG_BYTES + 4 * self.capacity
        self.block_bytes = nbytes
        self._epochs = {}
        self.mc_block = 0
        self._local = ctypes.c_void_p()
        self._mapped = []
        self._symm = None
        self.epoch = 0
        self.multicast = 0
        want = os.environ.get("KGE_PEER_BACKEND", "auto")
        ptrs, errors = None, []
        if want in ("auto", "symm"):
            ptrs, why = self._agree(self._try(self._init_symm, nbytes))
            errors += why
        if ptrs is None and want in ("auto", "ipc"):
            self.multicast = self.mc_block = 0
            ptrs, why = self._init_ipc(nbytes)
            errors += why
        if ptrs is None:
            self.close()
            raise _lib.KgeError("peer memory unavailable: " + ("; ".join(errors) or "no backend"))
        self.backend = "symm" if self._symm is not None else "ipc"
        self.block_ptrs = [int(p) for p in ptrs]
        self.extra = self._block[self.extra_offset:self.extra_offset + self.extra_bytes] if extra_bytes else None
        self.struct = _lib.KgePeerGroup(world=self.world, rank=self.rank)
        for r, p in enumerate(ptrs):
            self.struct.flags[r] = p
            self.struct.grad[r] = p + FLAG_BYTES
        # the in-switch reduction halves the NVLink bytes but has a flat cost (0.37-0.40 ms per 125 MB exchange at any G on
        # B200/NVSwitch); unicast loads/stores are faster up to 4 ranks (0.25 ms at 2, 0.35 at 4, 0.44 at 8)
        default_mc = "1" if self.world > 4 else "0"
        if self.multicast and os.environ.get("KGE_PEER_MULTICAST", default_mc) != "0":
            self.struct.multicast = self.multicast + FLAG_BYTES
        else:
            self.multicast = 0
        dist.barrier(group=group)            # nobody signals before every mapping exists and every block is zeroed

    @staticmethod
    def _try(fn, *args):
        try:
            return fn(*args), ""
        except Exception as exc:            # noqa: BLE001 -- any failure makes the whole group fall back together
            return None, "%s: %s" % (fn.__name__, exc)

    def _agree(self, result):
        """All ranks keep a backend only if it worked on every rank."""
        import torch.distributed as dist
        value, why = result
        oks = [None] * self.world
        dist.all_gather_object(oks, (value is not None, why), group=self.group)
        if all(o[0] for o in oks):
            return value, []
        return None, [o[1] for o in oks if o[1]][:1]

    def _init_symm(self, nbytes):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        block = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        handle = symm.rendezvous(block, self.group if self.group is not None else dist.group.WORLD)
        block.zero_()
        torch.cuda.synchronize(self.device)
        ptrs = [int(p) for p in handle.buffer_ptrs]
        shift = block.data_ptr() - ptrs[self.rank]           # the tensor may sit at an offset inside the allocation
        if shift < 0 or shift % 16:
            raise RuntimeError("unexpected symmetric-memory layout")
        mc = int(getattr(handle, "multicast_ptr", 0) or 0)
        self._symm, self._block = handle, block
        self.multicast = mc + shift if mc else 0
        self.mc_block = self.multicast                       # multicast address of the block base (entity-sharded step)
        self.workspace = block[FLAG_BYTES:FLAG_BYTES + 4 * self.capacity].view(torch.float32)
        return [p + shift for p in ptrs]

    def _init_ipc(self, nbytes):
        import torch.distributed as dist
        device = self.device

        def export():
            _lib.call("kge_peer_alloc", device.index, nbytes, ctypes.byref(self._local))
            buf = ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
            _lib.call("kge_peer_export", self._local, buf)
            return buf.raw

        handle, why = self._try(export)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (handle, socket.gethostname(), why), group=self.group)
        if not all(g[0] is not None for g in gathered):
            return None, [g[2] for g in gathered if g[2]][:1]
        if len({g[1] for g in gathered}) != 1:
            return None, ["ranks on different nodes"]

        def open_all():
            ptrs = []
            for r, g in enumerate(gathered):
                if r == self.rank:
                    ptrs.append(self._local.value)
                    continue
                p = ctypes.c_void_p()
                _lib.call("kge_peer_open", device.index, g[0], ctypes.byref(p))
                self._mapped.append(p)
                ptrs.append(p.value)
            return ptrs

        ptrs, why = self._agree(self._try(open_all))
        if ptrs is None:
            return None, why
        block = torch.as_tensor(_DeviceBlock(self._local.value, nbytes), device=device)
        self._block = block
        self.workspace = block[FLAG_BYTES:FLAG_BYTES + 4 * self.capacity].view(torch.float32)    # zero-initialised
        return ptrs, []

    def reduce_adam(self, entries, hyper, param_floats, region, row_offset, row_floats, rows_out, err, stream, l3=0.0):
        """Exchange one region (begin4, end4) of the parameter part.  entries: [(param_ptr, grad_ptr, exp_avg_ptr,
        exp_avg_sq_ptr, numel, step, 0)], as for kge_adam_step; row_floats > 0 also sums the loss rows."""
        self.epoch += 1
        lo4, hi4 = region_slices(region, self.world)[self.rank]
        tensors = (_lib.KgeAdamTensor * len(entries))(*[_lib.KgeAdamTensor(*c) for c in entries])
        _lib.call("kge_peer_reduce_adam", ctypes.byref(self.struct), self.epoch, tensors, len(entries), param_floats,
                  region[0], region[1], lo4, hi4, row_offset, row_floats,
                  ctypes.c_void_p(rows_out.data_ptr()) if row_floats else None, *hyper, float(l3),
                  ctypes.c_void_p(err.data_ptr()), stream)

    def barrier(self, channel, err, stream, exchange_err=False, phase=0):
        """Cross-GPU barrier on `stream` (kge_peer_barrier; channel 2 or 3, every rank calls it in the same order).
        phase 1 publishes the arrival only, a later phase 2 waits for everybody's."""
        if phase != 2:
            self._epochs[channel] = self._epochs.get(channel, 0) + 1
        _lib.call("kge_peer_barrier", ctypes.byref(self.struct), channel, self._epochs[channel], 1 if exchange_err else 0,
                  phase, ctypes.c_void_p(err.data_ptr()) if err is not None else None, stream)

    def shard(self, gather_offset, rows_max, rows_of):
        """kge_shard_t of the entity-sharded step for this group's blocks."""
        sh = _lib.KgeShard(world=self.world, rank=self.rank, block_bytes=self.block_bytes,
                           gather_offset=int(gather_offset), rows_max=int(rows_max))
        # KGE_SHARD_MULTICAST=0|1: mirror through the NVSwitch multicast mapping (symmetric-memory backend only)
        if self.mc_block and os.environ.get("KGE_SHARD_MULTICAST", "1") != "0":
            sh.multicast = self.mc_block
        for r in range(self.world):
            sh.block[r] = self.block_ptrs[r]
            sh.rows_of[r] = int(rows_of[r])
        return sh

    def close(self):
        """Unmap the peers' blocks and release the local one (the caller makes sure no exchange is in flight)."""
        self._symm = None
        lib = _lib.load()
        for p in self._mapped:
            lib.kge_peer_close(p)
        self._mapped = []
        if self._local:
            lib.kge_peer_free(self._local)
            self._local = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:                   # noqa: BLE001 -- interpreter shutdown: the driver reclaims the mappings
            pass


def moment_ranges(offsets, numels, regions, world):
    """Who keeps which part of the Adam moments current after kge_peer_reduce_adam steps over `regions`: per tensor
    (first workspace float `offsets[i]`, `numels[i]` elements) a tuple of (rank, begin, end) element ranges."""
    out = []
    for off, n in zip(offsets, numels):
        rs = []
        for region in regions:
            for r, (lo4, hi4) in enumerate(region_slices(region, world)):
                a, b = max(4 * lo4, off) - off, min(4 * hi4, off + n) - off
                if a < b:
                    rs.append((r, int(a), int(b)))
        out.append(tuple(rs))
    return tuple(out)


def entity_ranges(nentity, entity_dim, world):
    """Moment ownership of the entity table in the entity-sharded step: rank r owns the rows
    [r*base + min(r, rem), + base (+1 if r < rem)) -- the same split as kge_shard_t / Mirror::ent_base, ent_rem."""
    base, rem = divmod(int(nentity), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        if hi > lo:
            out.append((r, lo * int(entity_dim), hi * int(entity_dim)))
        lo = hi
    return tuple(out)


def gather_moment_ranges(tensors, ranges, group=None):
    """After sliced-optimizer steps exp_avg / exp_avg_sq are current only on the rank owning each range.  Broadcast
    every owned piece from its owner so that all ranks hold the full moments (before optimizer.state_dict(), or before
    switching to the replicated Adam / another ownership layout).  tensors: [(exp_avg, exp_avg_sq)] per parameter,
    ranges: per parameter the (rank, begin, end) element ranges (moment_ranges / entity_ranges)."""
    import torch.distributed as dist
    for (m, v), rs in zip(tensors, ranges):
        fm, fv = m.view(-1), v.view(-1)
        for r, a, b in rs:
            src = dist.get_global_rank(group, r) if group is not None else r
            dist.broadcast(fm[a:b], src=src, group=group)
            dist.broadcast(fv[a:b], src=src, group=group)


def gather_sliced_moments(tensors, offsets, regions, group=None):
    """gather_moment_ranges for the region layout of kge_peer_reduce_adam (offsets: each parameter's first float
    inside the workspace, regions: the exchange regions of the steps so far)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gather_moment_ranges(tensors, moment_ranges(offsets, [m.numel() for m, _ in tensors], regions, world), group)


__all__ = ["PeerExchange", "region_slices", "exchange_regions", "gather_sliced_moments", "moment_ranges",
           "entity_ranges", "gather_moment_ranges"]
