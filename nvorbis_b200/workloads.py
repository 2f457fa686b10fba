"""Synthetic batches for the BASELINE.json configs, built by re-sampling real unpacked frames.

A FramePool holds the boundary records (the C-ABI inputs: nvb_frame + posts + classes + entries) of a
real stream; the generators draw frames from it with a seeded numpy Generator so that the GPU path and
the CPU baseline decode identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class FramePool:
    channels: int
    post_stride: int
    block_size: tuple
    frames: np.ndarray        # capi.FRAME_DTYPE, offsets into classes / entries below
    posts: np.ndarray         # int16 [n, channels, post_stride]
    classes: np.ndarray       # uint8
    entries: np.ndarray       # uint16
    class_len: np.ndarray     # int64 [n]: classes each frame owns
    long_flag: np.ndarray     # bool [n]: block is a long block

    @staticmethod
    def from_npz(desc: dict, z: dict) -> "FramePool":
        frames = np.ascontiguousarray(z["pool_frames"]).view(capi.FRAME_DTYPE).reshape(-1)
        n = len(frames)
        posts = np.ascontiguousarray(z["pool_posts"], np.int16)
        stride = posts.size // max(n * desc["channels"], 1)
        return FramePool(desc["channels"], stride, tuple(desc["block_size"]), frames, posts.reshape(n, desc["channels"], stride),
                         np.ascontiguousarray(z["pool_classes"], np.uint8), np.ascontiguousarray(z["pool_entries"], np.uint16),
                         np.ascontiguousarray(z["pool_class_len"], np.int64), np.ascontiguousarray(z["pool_long"], bool))

    def gather(self, idx: np.ndarray) -> capi.HostBatch:
        """The batch made of pool frames idx[0], idx[1], ... in that order."""
        idx = np.asarray(idx, np.int64)
        fr = self.frames[idx].copy()
        ecount = np.where(fr["res_decoded"] != 0, fr["entry_count"], 0).astype(np.int64)
        ccount = self.class_len[idx]
        eoff = np.concatenate([[0], np.cumsum(ecount)])
        coff = np.concatenate([[0], np.cumsum(ccount)])

        def take(src, starts, counts, offs):
            total = int(offs[-1])
            if total == 0:
                return src[:0].copy()
            pos = np.arange(total, dtype=np.int64) - np.repeat(offs[:-1], counts) + np.repeat(starts, counts)
            return src[pos]

        entries = take(self.entries, self.frames["entries_off"][idx].astype(np.int64), ecount, eoff)
        classes = take(self.classes, self.frames["classes_off"][idx].astype(np.int64), ccount, coff)
        if eoff[-1] >= 2 ** 32 or coff[-1] >= 2 ** 32:
            raise ValueError("batch too large for 32-bit offsets")
        fr["entries_off"] = eoff[:-1]; fr["classes_off"] = coff[:-1]
        return capi.HostBatch(fr, self.posts[idx].reshape(-1), classes, entries)

    # ---- frame classes ------------------------------------------------------------------------------
    def long_long(self) -> np.ndarray:
        """Indices of long blocks whose both neighbours are long (window index 3, Mode.cs:44-50,135)."""
        return np.nonzero(self._ll_mask())[0]

    def _ll_mask(self) -> np.ndarray:
        f = self.frames
        half = self.block_size[1] // 2           # untrimmed long/long block: start 0, valid N/2 (an EOS-trimmed block is left out)
        return (f["status"] == capi.FRAME_OK) & self.long_flag & (f["window"] == 3) & (f["start"] == 0) & (f["valid"] == half)


def config2(pool: FramePool, n_frames: int = 4096, seed: int = 20240002) -> capi.HostBatch:
    """BASELINE configs[1]: stereo 44.1 kHz long-block N=2048 batch: long/long frames drawn with replacement."""
    cand = pool.long_long()
    rng = np.random.Generator(np.random.PCG64(seed))
    return pool.gather(cand[rng.integers(0, len(cand), n_frames)])


def config3(pool: FramePool, n_frames: int = 16384, seed: int = 20240003, min_run: int = 32) -> capi.HostBatch:
    """BASELINE configs[2]: mixed short/long window transitions: contiguous runs of real frames, cut only
    between two long/long blocks so that every block's window flags match its neighbours."""
    f = pool.frames
    ll = pool._ll_mask()
    cut = np.nonzero(ll[:-1] & ll[1:])[0] + 1          # a run may start at i if i-1 and i are long/long
    rng = np.random.Generator(np.random.PCG64(seed))
    out, total = [], 0
    while total < n_frames:
        a = int(cut[rng.integers(0, len(cut))])
        ends = cut[cut >= a + min_run]
        if len(ends) == 0:
            continue
        b = int(ends[rng.integers(0, min(len(ends), 8))])
        if not (f["status"][a:b] == capi.FRAME_OK).all():
            continue
        out.append(np.arange(a, b)); total += b - a
    return pool.gather(np.concatenate(out)[:n_frames])


def output_samples(batch: capi.HostBatch) -> int:
    """Upper bound of samples per channel the batch emits."""
    return capi.sum_output_bound(batch.frames)
