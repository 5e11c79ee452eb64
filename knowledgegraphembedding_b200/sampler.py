"""Device-side replacement for the reference's training input pipeline (codes/dataloader.py:13-119, 165-186):
`TrainDataset` + `DataLoader(shuffle=True)` + `BidirectionalOneShotIterator`, producing the same 4-tuple
`(positive [B,3] int64, negative [B,N] int64, subsampling_weight [B] f32, mode)` -- already on the GPU.

The one-off preprocessing (frequency counts, true-head / true-tail lists) is vectorised numpy on the host; every
batch then costs one gather and one launch of `kge_sample_negatives`.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .model import _ptr, _stream


def subsampling_weights(triples, start=4):
    """sqrt(1 / (count(h,r) + count(t,-r-1))) with both counts starting at `start` (dataloader.py:31-34,77-93),
    evaluated in fp32 like `torch.sqrt(1 / torch.Tensor([c]))`."""
    tri = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    h, r, t = tri[:, 0], tri[:, 1], tri[:, 2]
    nrel = int(r.max()) + 1 if tri.size else 1

    def occurrences(a, b):
        key = a * nrel + b
        _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        return cnt[inv]

    c = (start - 1 + occurrences(h, r)) + (start - 1 + occurrences(t, r))
    return np.sqrt(np.float32(1.0) / c.astype(np.float32)).astype(np.float32)


def true_lists(triples, nentity, nrelation, mode):
    """Per-triple (start, len) into a concatenation of sorted unique true entities: the true heads of (r,t) for
    head-batch, the true tails of (h,r) for tail-batch (dataloader.py:95-119)."""
    tri = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    h, r, t = tri[:, 0], tri[:, 1], tri[:, 2]
    if mode == 'head-batch':
        key, val = r * nentity + t, h
    elif mode == 'tail-batch':
        key, val = h * nrelation + r, t
    else:
        raise ValueError('Training batch mode %s not supported' % mode)          # dataloader.py:56
    pairs = np.unique(np.stack([key, val], axis=1), axis=0)                       # sorted by key, then value; unique
    ukeys, starts, counts = np.unique(pairs[:, 0], return_index=True, return_counts=True)
    which = np.searchsorted(ukeys, key)
    return starts[which].astype(np.int32), counts[which].astype(np.int32), pairs[:, 1].astype(np.int32)


class GpuTrainDataset:
    """One mode's worth of the reference's TrainDataset + shuffling DataLoader, on the device."""

    def __init__(self, triples, nentity, nrelation, negative_sample_size, mode, batch_size, device, seed=0):
        self.nentity, self.nrelation, self.N = int(nentity), int(nrelation), int(negative_sample_size)
        self.mode, self.batch_size, self.seed = mode, int(batch_size), int(seed)
        self.device = torch.device(device)
        tri = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        self.len = tri.shape[0]
        start, length, ents = true_lists(tri, nentity, nrelation, mode)
        self.triples = torch.from_numpy(tri).to(self.device)
        self.weights = torch.from_numpy(subsampling_weights(tri)).to(self.device)
        self.key_start = torch.from_numpy(start).to(self.device)
        self.key_len = torch.from_numpy(length).to(self.device)
        self.true_entities = torch.from_numpy(ents).to(self.device)
        self.step = 0
        self.epoch = 0
        self._perm = None
        self._cursor = 0
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(self.seed * 2 + (1 if mode == 'head-batch' else 0))

    def __len__(self):
        return self.len

    def sample(self, index, step=None):
        """Batch for the given train-triple indices (int64 device tensor)."""
        index = index.to(self.device, torch.int64).contiguous()
        B = index.shape[0]
        negative = torch.empty((B, self.N), dtype=torch.int64, device=self.device)
        step = self.step if step is None else step
        seed = (self.seed << 1) | (1 if self.mode == 'head-batch' else 0)
        _lib.call("kge_sample_negatives", _ptr(index), _ptr(self.key_start), _ptr(self.key_len),
                  _ptr(self.true_entities), B, self.N, self.nentity, ctypes.c_uint64(seed), ctypes.c_uint64(step),
                  _ptr(negative), _stream(self.device))
        return self.triples[index], negative, self.weights[index], self.mode

    def __iter__(self):
        return self

    def __next__(self):
        """Endless stream of shuffled epochs; the last batch of an epoch may be smaller (no drop_last, run.py:246)."""
        if self._perm is None or self._cursor >= self.len:
            self._perm = torch.randperm(self.len, device=self.device, generator=self._gen)
            self._cursor = 0
            self.epoch += 1
        index = self._perm[self._cursor:self._cursor + self.batch_size]
        self._cursor += self.batch_size
        self.step += 1
        return self.sample(index)


class BidirectionalGpuIterator:
    """Same alternation as BidirectionalOneShotIterator (dataloader.py:165-177): tail-batch on odd steps (the first
    call), head-batch on even steps."""

    def __init__(self, train_triples, nentity, nrelation, negative_sample_size, batch_size, device='cuda', seed=0):
        args = (train_triples, nentity, nrelation, negative_sample_size)
        self.iterator_head = GpuTrainDataset(*args, 'head-batch', batch_size, device, seed)
        self.iterator_tail = GpuTrainDataset(*args, 'tail-batch', batch_size, device, seed)
        self.step = 0

    def __iter__(self):
        return self

    def __next__(self):
        self.step += 1
        return next(self.iterator_head) if self.step % 2 == 0 else next(self.iterator_tail)
