"""ctypes binding of libnvorbis_host.so (include/nvorbis_host.h): the CPU half of the split decoder --
Ogg demux, header parsing and bit unpacking into nvb_setup / nvb_batch arrays.  No GPU, no samples."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

DEFAULT_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnvorbis_host.so")


class Info(C.Structure):
    _fields_ = [("channels", C.c_int32), ("sample_rate", C.c_int32), ("block_size", C.c_int32 * 2), ("n_books", C.c_int32),
                ("n_floors", C.c_int32), ("n_residues", C.c_int32), ("n_mappings", C.c_int32), ("n_modes", C.c_int32),
                ("post_stride", C.c_int32), ("n_packets", C.c_int64), ("n_audio_packets", C.c_int64), ("last_granule", C.c_int64),
                ("has_eos", C.c_int32), ("floor0_stride", C.c_int32)]


_lib = None


def load_library(path: str | None = None):
    global _lib
    if _lib is None:
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing: run __graft_entry__.build()")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.nvh_open_ogg.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.nvh_open_packets.argtypes = [vp, vp, vp, vp, C.c_int64, C.POINTER(vp)]
        L.nvh_open_forward.argtypes = [C.POINTER(vp)]
        L.nvh_feed.restype = C.c_int64; L.nvh_feed.argtypes = [vp, vp, C.c_size_t, C.c_int]
        L.nvh_ogg_stream_count.argtypes = [vp, C.c_size_t]
        L.nvh_open_ogg_stream.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(vp)]
        L.nvh_close.argtypes = [vp]
        L.nvh_last_error.restype = C.c_char_p; L.nvh_last_error.argtypes = [vp]
        L.nvh_get_info.argtypes = [vp, C.POINTER(Info)]
        L.nvh_setup.restype = C.POINTER(capi.SetupStruct); L.nvh_setup.argtypes = [vp]
        L.nvh_packet_size.restype = C.c_int64; L.nvh_packet_size.argtypes = [vp, C.c_int64]
        L.nvh_packet_get.argtypes = [vp, C.c_int64, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.nvh_unpack.restype = C.c_int64
        L.nvh_unpack.argtypes = [vp, C.c_int64, C.c_int, C.POINTER(capi.BatchStruct), C.POINTER(C.c_int32)]
        L.nvh_rewind.argtypes = [vp]
        L.nvh_seek.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64)]
        L.nvh_total_samples.restype = C.c_int64; L.nvh_total_samples.argtypes = [vp]
        L.nvh_unpack_tables.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.nvh_packet_batch.restype = C.c_int64
        L.nvh_packet_batch.argtypes = [vp, C.c_int64, C.POINTER(capi.PacketBatchStruct), C.POINTER(C.c_int32)]
        _lib = L
    return _lib


class HostError(RuntimeError):
    pass


class SetupView:
    """An nvb_setup owned by an nvh_stream, usable wherever a capi.Setup is."""

    def __init__(self, ptr, channels, sample_rate, block_size, owner):
        self.struct = ptr.contents
        self.channels, self.sample_rate, self.block_size = channels, sample_rate, block_size
        self._owner = owner


def ogg_stream_count(data) -> int:
    """Logical streams (serial numbers) of an Ogg container image."""
    a = np.frombuffer(bytes(data), np.uint8)
    n = load_library().nvh_ogg_stream_count(a.ctypes.data, a.size)
    if n < 0:
        raise HostError(f"nvh_ogg_stream_count: status {n}")
    return int(n)


class HostStream:
    """The unpacking half of a StreamDecoder over an Ogg file image or a packet list."""

    def __init__(self, data=None, packets=None, stream_index: int = 0, forward: bool = False):
        L = load_library()
        h = C.c_void_p()
        if forward:                                     # forward-only input: bytes arrive through feed()
            rc = L.nvh_open_forward(C.byref(h))
            if rc != 0:
                raise HostError(f"nvh_open_forward: status {rc}")
            self.lib, self.handle, self.info = L, h, None
            self.channels = self.sample_rate = self.post_stride = self.floor0_stride = self.n_audio_packets = 0
            self.block_size = (0, 0)
            return
        if packets is not None:
            d, sizes, gran, flags = (np.ascontiguousarray(packets[0], np.uint8), np.ascontiguousarray(packets[1], np.int64),
                                     np.ascontiguousarray(packets[2], np.int64), np.ascontiguousarray(packets[3], np.uint8))
            rc = L.nvh_open_packets(d.ctypes.data, sizes.ctypes.data, gran.ctypes.data, flags.ctypes.data, len(sizes), C.byref(h))
        else:
            self._data = np.frombuffer(bytes(data), np.uint8)
            rc = L.nvh_open_ogg_stream(self._data.ctypes.data, self._data.size, int(stream_index), C.byref(h))
        if rc != 0:
            raise HostError(f"open failed: status {rc} ({L.nvh_last_error(None).decode()})")
        self.lib, self.handle = L, h
        self._refresh_info()

    def _refresh_info(self):
        info = Info()
        self.lib.nvh_get_info(self.handle, C.byref(info))
        self.info = info
        self.channels, self.sample_rate = info.channels, info.sample_rate
        self.block_size = (info.block_size[0], info.block_size[1])
        self.post_stride = info.post_stride
        self.floor0_stride = info.floor0_stride
        self.n_audio_packets = int(info.n_audio_packets)

    def feed(self, data, end_of_input: bool = False) -> int:
        """nvh_feed: the next container bytes of a forward-only stream; returns the audio packets demuxed so far."""
        a = np.frombuffer(bytes(data), np.uint8)
        n = self.lib.nvh_feed(self.handle, a.ctypes.data if a.size else None, a.size, 1 if end_of_input else 0)
        if n < 0:
            raise HostError(f"nvh_feed: status {n} ({self.lib.nvh_last_error(self.handle).decode()})")
        if n > 0 or self.channels:
            self._refresh_info()
        return int(n)

    def close(self):
        if self.handle:
            self.lib.nvh_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self) -> SetupView:
        return SetupView(self.lib.nvh_setup(self.handle), self.channels, self.sample_rate, self.block_size, self)

    def rewind(self):
        self.lib.nvh_rewind(self.handle)

    def packet(self, i: int):
        n = self.lib.nvh_packet_size(self.handle, i)
        if n < 0:
            raise IndexError(i)
        buf = np.zeros(max(n, 1), np.uint8)
        g, f = C.c_int64(), C.c_int32()
        self.lib.nvh_packet_get(self.handle, i, buf.ctypes.data, C.byref(g), C.byref(f))
        return bytes(buf[:n]), int(g.value), int(f.value)

    def unpack(self, count: int, threads: int = 0, copy: bool = True):
        """Unpacks the next `count` audio packets.  Returns (HostBatch, end_of_stream)."""
        b = capi.BatchStruct()
        eos = C.c_int32()
        n = self.lib.nvh_unpack(self.handle, count, threads or (os.cpu_count() or 1), C.byref(b), C.byref(eos))
        if n < 0:
            raise HostError(f"nvh_unpack: status {n} ({self.lib.nvh_last_error(self.handle).decode()})")

        def view(ptr, dtype, count_):
            if not ptr or count_ == 0:
                return np.zeros(0, dtype)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count_ * np.dtype(dtype).itemsize,)).view(dtype)
            return a.copy() if copy else a

        frames = view(b.frames, capi.FRAME_DTYPE, n)
        posts = view(b.posts, np.int16, n * self.channels * self.post_stride)
        classes = view(b.classes, np.uint8, b.n_classes)
        entries = view(b.entries, np.uint16, b.n_entries)
        floor0 = view(b.floor0, np.float32, n * self.channels * self.floor0_stride) if b.floor0 else None
        return capi.HostBatch(frames, posts, classes, entries, floor0), bool(eos.value)

    def seek(self, sample_position: int) -> int:
        """nvh_seek: cursor on the pre-roll packet; returns the samples per channel to drop from the output that follows."""
        skip = C.c_int64()
        rc = self.lib.nvh_seek(self.handle, int(sample_position), C.byref(skip))
        if rc != 0:
            raise HostError(f"nvh_seek: status {rc} ({self.lib.nvh_last_error(self.handle).decode()})")
        return int(skip.value)

    def total_samples(self) -> int:
        n = self.lib.nvh_total_samples(self.handle)
        if n < 0:
            raise HostError(f"nvh_total_samples: status {n}")
        return int(n)

    # ---- host half of the GPU-side packet unpack ------------------------------------------------------------------------
    def unpack_tables(self) -> np.ndarray:
        """The unpack tables of this stream's setup (for capi.Context.upload_unpack_tables)."""
        p, n = C.c_void_p(), C.c_size_t()
        rc = self.lib.nvh_unpack_tables(self.handle, C.byref(p), C.byref(n))
        if rc != 0:
            raise HostError(f"nvh_unpack_tables: status {rc} ({self.lib.nvh_last_error(self.handle).decode()})")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)).copy()

    def packet_batch(self, count: int, copy: bool = True):
        """The next `count` audio packets as raw bytes + per-packet header values.  Returns (capi.PacketBatch, end_of_stream)."""
        b = capi.PacketBatchStruct()
        eos = C.c_int32()
        n = self.lib.nvh_packet_batch(self.handle, count, C.byref(b), C.byref(eos))
        if n < 0:
            raise HostError(f"nvh_packet_batch: status {n} ({self.lib.nvh_last_error(self.handle).decode()})")

        def view(ptr, dtype, count_):
            if not ptr or count_ == 0:
                return np.zeros(0, dtype)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count_ * np.dtype(dtype).itemsize,)).view(dtype)
            return a.copy() if copy else a

        offsets = view(b.offsets, np.uint32, n + 1) if n else np.zeros(1, np.uint32)
        frames = view(b.frames, capi.FRAME_DTYPE, n)
        data = view(b.data, np.uint8, int(offsets[-1]))
        return capi.PacketBatch(frames, data, offsets), bool(eos.value)
