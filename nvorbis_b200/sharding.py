"""Sharding of a frame batch across GPUs: contiguous frame ranges, each shard re-decodes the
frame before its range as a halo whose output is dropped -- exactly the reference's "first block of a stream emits
nothing, it only leaves its tail" rule (StreamDecoder.cs:446-450) -- so shards need no data-path communication."""
from __future__ import annotations

import numpy as np

from . import capi


def shard_cuts(frames: np.ndarray, world: int) -> list:
    """world+1 cut points; a shard may only start after a decoded block (status OK): a failed packet drains the previous
    tail (StreamDecoder.cs:352-356), which a fresh decoder state cannot reproduce."""
    n = len(frames)
    cuts = [0]
    for k in range(1, world):
        c = max(cuts[-1], (n * k) // world)
        while 0 < c < n and frames["status"][c - 1] != capi.FRAME_OK:
            c += 1
        cuts.append(min(c, n))
    cuts.append(n)
    return cuts


def slice_batch(b: capi.HostBatch, lo: int, hi: int, channels: int) -> capi.HostBatch:
    """Frames [lo, hi) of a batch with the class / entry offsets rebased."""
    n = len(b.frames)
    fr = b.frames[lo:hi].copy()
    stride = b.posts.size // max(n * channels, 1)
    if hi <= lo:
        return capi.HostBatch(fr, np.zeros(0, np.int16), np.zeros(0, np.uint8), np.zeros(0, np.uint16))
    c0, e0 = int(b.frames["classes_off"][lo]), int(b.frames["entries_off"][lo])
    c1 = int(b.frames["classes_off"][hi]) if hi < n else b.classes.size
    e1 = int(b.frames["entries_off"][hi]) if hi < n else b.entries.size
    fr["classes_off"] -= c0; fr["entries_off"] -= e0
    return capi.HostBatch(fr, b.posts[lo * channels * stride: hi * channels * stride], b.classes[c0:c1], b.entries[e0:e1])


def take_shard(b: capi.HostBatch, cuts: list, rank: int, channels: int) -> capi.HostBatch:
    """The sub-batch rank `rank` decodes: its range plus one halo frame in front (none for the first shard)."""
    lo, hi = cuts[rank], cuts[rank + 1]
    return slice_batch(b, lo - 1 if lo > 0 else 0, hi, channels)
