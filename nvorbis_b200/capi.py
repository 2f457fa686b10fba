"""ctypes binding of the C ABI in include/nvorbis_b200.h (libnvorbis_b200.so).

This is what a P/Invoke host would bind; the Python layer only marshals
numpy arrays into the plain-pointer structs.  There is no CPU implementation behind it: if the
shared library is missing, or no sm_100 GPU is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

ABI_VERSION = 3
MAX_CHANNELS, MAX_POSTS, MAX_CLASSES, MAX_STAGES, MAX_COUPLING = 32, 64, 64, 8, 256
MAX_IN_FLIGHT = 3                                   # NVB_MAX_IN_FLIGHT: batches between decode_batch_begin and decode_batch_end

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM, ERR_STATE, ERR_CAPACITY, ERR_DATA = 0, -1, -2, -3, -4, -5, -6, -7
FRAME_OK, FRAME_FAILED = 0, 1
RUN_DEFAULT, RUN_EXACT, RUN_NO_CLIP, RUN_CONTINUE, RUN_PCM_S16, RUN_DEVICE_OUT, RUN_ONE_KERNEL, RUN_TWO_KERNELS = 0, 1, 2, 4, 8, 16, 32, 64

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.environ.get("NVB_LIB_PATH") or os.path.join(_HERE, "libnvorbis_b200.so")      # NVB_LIB_PATH: development hook (kernel build variants)


class NvbError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"{what}: status {status}" + (f" ({detail})" if detail else ""))


class Codebook(C.Structure):
    _fields_ = [("dims", C.c_int32), ("entries", C.c_int32), ("map_type", C.c_int32), ("reserved", C.c_int32), ("table_off", C.c_int64)]


class Floor1(C.Structure):
    _fields_ = [("n_posts", C.c_int32), ("multiplier", C.c_int32), ("range", C.c_int32), ("reserved", C.c_int32),
                ("x_list", C.c_uint16 * MAX_POSTS), ("l_neigh", C.c_uint8 * MAX_POSTS), ("h_neigh", C.c_uint8 * MAX_POSTS),
                ("sort_idx", C.c_uint8 * MAX_POSTS)]


class Floor0(C.Structure):
    _fields_ = [("order", C.c_int32), ("rate", C.c_int32), ("bark_map_size", C.c_int32), ("amp_bits", C.c_int32), ("amp_ofs", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class Floor(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("f1", Floor1), ("f0", Floor0)]


class Residue(C.Structure):
    _fields_ = [("type", C.c_int32), ("begin", C.c_int32), ("end", C.c_int32), ("partition_size", C.c_int32),
                ("classifications", C.c_int32), ("max_stages", C.c_int32), ("cascade", C.c_int32 * MAX_CLASSES),
                ("books", (C.c_int16 * MAX_STAGES) * MAX_CLASSES)]


class Mapping(C.Structure):
    _fields_ = [("n_coupling", C.c_int32), ("n_submaps", C.c_int32), ("magnitude", C.c_uint8 * MAX_COUPLING),
                ("angle", C.c_uint8 * MAX_COUPLING), ("floor", C.c_int32), ("residue", C.c_int32)]


class Mode(C.Structure):
    _fields_ = [("block_flag", C.c_int32), ("mapping", C.c_int32)]


class SetupStruct(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("channels", C.c_int32), ("sample_rate", C.c_int32), ("block_size", C.c_int32 * 2),
                ("n_books", C.c_int32), ("n_floors", C.c_int32), ("n_residues", C.c_int32), ("n_mappings", C.c_int32), ("n_modes", C.c_int32),
                ("books", C.POINTER(Codebook)), ("vq_floats", C.POINTER(C.c_float)), ("n_vq_floats", C.c_int64),
                ("floors", C.POINTER(Floor)), ("residues", C.POINTER(Residue)), ("mappings", C.POINTER(Mapping)), ("modes", C.POINTER(Mode)),
                ("window_slope", C.POINTER(C.c_float) * 2), ("mdct_a", C.POINTER(C.c_float) * 2), ("mdct_b", C.POINTER(C.c_float) * 2),
                ("mdct_c", C.POINTER(C.c_float) * 2), ("mdct_bitrev", C.POINTER(C.c_uint16) * 2)]


class BatchStruct(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("reserved", C.c_int32), ("frames", C.c_void_p), ("posts", C.c_void_p),
                ("classes", C.c_void_p), ("n_classes", C.c_int64), ("entries", C.c_void_p), ("n_entries", C.c_int64), ("floor0", C.c_void_p)]


class PacketBatchStruct(C.Structure):
    _fields_ = [("n_packets", C.c_int32), ("reserved", C.c_int32), ("frames", C.c_void_p), ("data", C.c_void_p), ("offsets", C.c_void_p)]


class ResultStruct(C.Structure):
    _fields_ = [("samples_per_channel", C.c_int64), ("has_clipped", C.c_int32), ("n_failed", C.c_int32),
                ("n_floor_range", C.c_int32), ("n_inconsistent", C.c_int32)]


# nvb_frame
FRAME_DTYPE = np.dtype([("status", "u1"), ("mode", "u1"), ("window", "u1"), ("res_decoded", "u1"), ("exec_mask", "<u4"),
                        ("start", "<i4"), ("valid", "<i4"), ("total", "<i4"), ("classes_off", "<u4"), ("entries_off", "<u4"),
                        ("entry_count", "<u4")], align=True)
assert FRAME_DTYPE.itemsize == 32

EXPORTS = [
    "nvb_abi_version", "nvb_strerror", "nvb_last_error", "nvb_create", "nvb_destroy", "nvb_host_alloc", "nvb_host_free",
    "nvb_upload_setup", "nvb_setup_blob_size", "nvb_setup_blob_export", "nvb_setup_blob_import", "nvb_post_stride", "nvb_floor0_stride", "nvb_reset",
    "nvb_decode_batch", "nvb_decode_batch_begin", "nvb_decode_batch_end", "nvb_dbatch_create", "nvb_dbatch_samples", "nvb_dbatch_run", "nvb_dbatch_result", "nvb_dbatch_destroy",
    "nvb_dbatch_run_spectrum", "nvb_dbatch_run_imdct", "nvb_dbatch_spectrum_floats", "nvb_dbatch_launches",
    "nvb_upload_unpack_tables", "nvb_decode_packets", "nvb_decode_packets_begin", "nvb_unpack_strides", "nvb_unpack_packets",
]

_libs: dict = {}


def load_library(path: str | None = None):
    """Loads libnvorbis_b200.so (built by __graft_entry__.build() / make -C nvorbis_b200/csrc)."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback for the synthesis path)")
    L = C.CDLL(path)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    L.nvb_abi_version.restype = i32
    L.nvb_strerror.restype = C.c_char_p; L.nvb_strerror.argtypes = [i32]
    L.nvb_last_error.restype = C.c_char_p; L.nvb_last_error.argtypes = [vp]
    L.nvb_create.argtypes = [i32, C.POINTER(vp)]
    L.nvb_destroy.argtypes = [vp]
    L.nvb_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.nvb_host_free.argtypes = [vp]
    L.nvb_upload_setup.argtypes = [vp, C.POINTER(SetupStruct)]
    L.nvb_setup_blob_size.argtypes = [vp, C.POINTER(sz)]
    L.nvb_setup_blob_export.argtypes = [vp, vp, sz]
    L.nvb_setup_blob_import.argtypes = [vp, vp, sz]
    L.nvb_post_stride.argtypes = [vp]
    L.nvb_floor0_stride.argtypes = [vp]
    L.nvb_reset.argtypes = [vp]
    L.nvb_decode_batch.argtypes = [vp, C.POINTER(BatchStruct), i32, vp, sz, C.POINTER(ResultStruct)]
    L.nvb_decode_batch_begin.argtypes = [vp, C.POINTER(BatchStruct), i32, vp, sz]
    L.nvb_decode_batch_end.argtypes = [vp, C.POINTER(ResultStruct)]
    L.nvb_dbatch_create.argtypes = [vp, C.POINTER(BatchStruct), i32, C.POINTER(vp)]
    L.nvb_dbatch_samples.restype = i64; L.nvb_dbatch_samples.argtypes = [vp]
    L.nvb_dbatch_spectrum_floats.restype = i64; L.nvb_dbatch_spectrum_floats.argtypes = [vp]
    L.nvb_dbatch_launches.argtypes = [vp]
    L.nvb_dbatch_run.argtypes = [vp, vp, vp, vp]
    L.nvb_dbatch_run_spectrum.argtypes = [vp, vp, vp, vp]
    L.nvb_dbatch_run_imdct.argtypes = [vp, vp, vp, vp, vp]
    L.nvb_dbatch_result.argtypes = [vp, vp, vp, C.POINTER(ResultStruct)]
    L.nvb_dbatch_destroy.argtypes = [vp, vp]
    L.nvb_upload_unpack_tables.argtypes = [vp, vp, sz]
    L.nvb_decode_packets.argtypes = [vp, C.POINTER(PacketBatchStruct), i32, vp, sz, C.POINTER(ResultStruct)]
    L.nvb_decode_packets_begin.argtypes = [vp, C.POINTER(PacketBatchStruct), i32, vp, sz]
    L.nvb_unpack_strides.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.nvb_unpack_packets.argtypes = [vp, C.POINTER(PacketBatchStruct), vp, vp, vp, vp]
    _libs[path] = L
    return L


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class Setup:
    """Owns the arrays an nvb_setup points to.  Built from plain Python/numpy descriptions of what
    StreamDecoder.LoadBooks leaves behind (StreamDecoder.cs:226-289)."""

    def __init__(self, channels: int, sample_rate: int, block_size, books, floors, residues, mappings, modes,
                 window_slope=(None, None), mdct_tables=(None, None)):
        self.channels, self.sample_rate, self.block_size = int(channels), int(sample_rate), (int(block_size[0]), int(block_size[1]))
        nb = len(books)
        self._books = (Codebook * max(nb, 1))()
        tables, off = [], 0
        for i, b in enumerate(books):
            t = b.get("table")
            has = b["map_type"] != 0 and t is not None and len(t) > 0
            self._books[i] = Codebook(int(b["dims"]), int(b["entries"]), int(b["map_type"]), 0, off if has else -1)
            if has:
                t = np.ascontiguousarray(t, np.float32)
                tables.append(t); off += t.size
        self._vq = np.concatenate(tables) if tables else np.zeros(1, np.float32)
        self._n_vq = off
        self._floors = (Floor * max(len(floors), 1))()
        for i, f in enumerate(floors):
            fl = Floor(); fl.type = int(f["type"])
            if fl.type == 1:
                n = int(f["n_posts"])
                fl.f1.n_posts, fl.f1.multiplier, fl.f1.range = n, int(f["multiplier"]), int(f["range"])
                for k in range(min(n, MAX_POSTS)):
                    fl.f1.x_list[k] = int(f["x_list"][k]); fl.f1.l_neigh[k] = int(f["l_neigh"][k])
                    fl.f1.h_neigh[k] = int(f["h_neigh"][k]); fl.f1.sort_idx[k] = int(f["sort_idx"][k])
            elif fl.type == 0:
                fl.f0.order, fl.f0.rate, fl.f0.bark_map_size = int(f["order"]), int(f["rate"]), int(f["bark_map_size"])
                fl.f0.amp_bits, fl.f0.amp_ofs = int(f["amp_bits"]), int(f["amp_ofs"])
            self._floors[i] = fl
        self._residues = (Residue * max(len(residues), 1))()
        for i, r in enumerate(residues):
            rs = Residue()
            rs.type, rs.begin, rs.end, rs.partition_size = int(r["type"]), int(r["begin"]), int(r["end"]), int(r["partition_size"])
            rs.classifications, rs.max_stages = int(r["classifications"]), int(r["max_stages"])
            for c in range(MAX_CLASSES):
                for s in range(MAX_STAGES):
                    rs.books[c][s] = -1
            for c in range(min(rs.classifications, MAX_CLASSES)):
                rs.cascade[c] = int(r["cascade"][c])
                row = r["books"][c]
                for s in range(min(len(row), MAX_STAGES)):
                    rs.books[c][s] = int(row[s])
            self._residues[i] = rs
        self._mappings = (Mapping * max(len(mappings), 1))()
        for i, m in enumerate(mappings):
            mp = Mapping()
            mp.n_coupling, mp.n_submaps, mp.floor, mp.residue = int(m["n_coupling"]), int(m["n_submaps"]), int(m["floor"]), int(m["residue"])
            for k in range(min(mp.n_coupling, MAX_COUPLING)):
                mp.magnitude[k] = int(m["magnitude"][k]); mp.angle[k] = int(m["angle"][k])
            self._mappings[i] = mp
        self._modes = (Mode * max(len(modes), 1))()
        for i, m in enumerate(modes):
            self._modes[i] = Mode(int(m["block_flag"]), int(m["mapping"]))
        self._slopes = [None if s is None else np.ascontiguousarray(s, np.float32) for s in window_slope]
        self._mdct = [None if t is None else tuple(np.ascontiguousarray(x, d) for x, d in zip(t, (np.float32, np.float32, np.float32, np.uint16)))
                      for t in mdct_tables]
        s = SetupStruct()
        s.abi_version, s.channels, s.sample_rate = ABI_VERSION, self.channels, self.sample_rate
        s.block_size[0], s.block_size[1] = self.block_size
        s.n_books, s.n_floors, s.n_residues, s.n_mappings, s.n_modes = nb, len(floors), len(residues), len(mappings), len(modes)
        s.books = C.cast(self._books, C.POINTER(Codebook)); s.vq_floats = _fptr(self._vq); s.n_vq_floats = self._n_vq
        s.floors = C.cast(self._floors, C.POINTER(Floor)); s.residues = C.cast(self._residues, C.POINTER(Residue))
        s.mappings = C.cast(self._mappings, C.POINTER(Mapping)); s.modes = C.cast(self._modes, C.POINTER(Mode))
        for i in range(2):
            if self._slopes[i] is not None:
                s.window_slope[i] = _fptr(self._slopes[i])
            if self._mdct[i] is not None:
                a, b, c, br = self._mdct[i]
                s.mdct_a[i], s.mdct_b[i], s.mdct_c[i] = _fptr(a), _fptr(b), _fptr(c)
                s.mdct_bitrev[i] = br.ctypes.data_as(C.POINTER(C.c_uint16))
        self.struct = s


@dataclass
class Result:
    samples_per_channel: int
    has_clipped: bool
    n_failed: int
    n_floor_range: int
    n_inconsistent: int


def _result(r: ResultStruct) -> Result:
    return Result(int(r.samples_per_channel), bool(r.has_clipped), int(r.n_failed), int(r.n_floor_range), int(r.n_inconsistent))


class HostBatch:
    """The four host arrays of an nvb_batch (kept alive while the struct is in use)."""

    def __init__(self, frames, posts, classes, entries, floor0=None):
        self.floor0 = None if floor0 is None else np.ascontiguousarray(floor0, np.float32)
        self.frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        self.posts = np.ascontiguousarray(posts, np.int16)
        self.classes = np.ascontiguousarray(classes, np.uint8)
        self.entries = np.ascontiguousarray(entries, np.uint16)
        b = BatchStruct()
        b.n_frames = len(self.frames)
        b.frames = self.frames.ctypes.data if len(self.frames) else None
        b.posts = self.posts.ctypes.data if self.posts.size else None
        b.classes = self.classes.ctypes.data if self.classes.size else None
        b.n_classes = self.classes.size
        b.entries = self.entries.ctypes.data if self.entries.size else None
        b.n_entries = self.entries.size
        b.floor0 = self.floor0.ctypes.data if self.floor0 is not None and self.floor0.size else None
        self.struct = b

    @property
    def h2d_bytes(self) -> int:
        return self.frames.nbytes + self.posts.nbytes + self.classes.nbytes + self.entries.nbytes + (self.floor0.nbytes if self.floor0 is not None else 0)


class PacketBatch:
    """Owns the arrays an nvb_packet_batch points to: raw audio packets + the per-packet header values the host read
    (Mode.GetPacketInfo, Mode.cs:119-151).  The GPU unpacks them (nvb_decode_packets)."""

    def __init__(self, frames: np.ndarray, data: np.ndarray, offsets: np.ndarray):
        self.frames = np.ascontiguousarray(frames, FRAME_DTYPE)
        self.data = np.ascontiguousarray(data, np.uint8)
        self.offsets = np.ascontiguousarray(offsets, np.uint32)
        assert self.offsets.size == len(self.frames) + 1
        if self.data.size == 0:
            self.data = np.zeros(1, np.uint8)
        self.struct = PacketBatchStruct(len(self.frames), 0, self.frames.ctypes.data, self.data.ctypes.data, self.offsets.ctypes.data)

    @property
    def h2d_bytes(self) -> int:
        return int(self.offsets[-1]) + self.offsets.nbytes + len(self.frames) * 64


class DeviceBatch:
    def __init__(self, ctx: "Context", handle):
        self.ctx, self.handle = ctx, handle

    @property
    def samples(self) -> int:
        return int(self.ctx.lib.nvb_dbatch_samples(self.handle))

    @property
    def spectrum_floats(self) -> int:
        return int(self.ctx.lib.nvb_dbatch_spectrum_floats(self.handle))

    @property
    def launches(self) -> int:
        return int(self.ctx.lib.nvb_dbatch_launches(self.handle))

    def run(self, d_pcm: int, stream: int = 0):
        self.ctx._check(self.ctx.lib.nvb_dbatch_run(self.ctx.handle, self.handle, d_pcm, stream), "nvb_dbatch_run")

    def run_spectrum(self, d_spectrum: int, stream: int = 0):
        self.ctx._check(self.ctx.lib.nvb_dbatch_run_spectrum(self.ctx.handle, self.handle, d_spectrum, stream), "nvb_dbatch_run_spectrum")

    def run_imdct(self, d_spectrum: int, d_pcm: int, stream: int = 0):
        self.ctx._check(self.ctx.lib.nvb_dbatch_run_imdct(self.ctx.handle, self.handle, d_spectrum, d_pcm, stream), "nvb_dbatch_run_imdct")

    def result(self, stream: int = 0) -> Result:
        r = ResultStruct()
        self.ctx._check(self.ctx.lib.nvb_dbatch_result(self.ctx.handle, self.handle, stream, C.byref(r)), "nvb_dbatch_result")
        return _result(r)

    def destroy(self):
        if self.handle:
            self.ctx.lib.nvb_dbatch_destroy(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.destroy()
        except Exception:
            pass


class Context:
    """nvb_ctx: one synthesis context on one GPU (single-threaded, like a reference StreamDecoder)."""

    def __init__(self, device: int = 0, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        h = C.c_void_p()
        rc = self.lib.nvb_create(device, C.byref(h))
        if rc != OK:
            raise NvbError(rc, "nvb_create", self.lib.nvb_last_error(None).decode())
        self.handle = h
        self.channels = 0
        self._setup = None

    def _check(self, rc: int, what: str):
        if rc != OK:
            raise NvbError(rc, what, self.lib.nvb_last_error(self.handle).decode())

    def close(self):
        if self.handle:
            self.lib.nvb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_setup(self, setup: Setup):
        self._check(self.lib.nvb_upload_setup(self.handle, C.byref(setup.struct)), "nvb_upload_setup")
        self._setup = setup
        self.channels = setup.channels

    def export_blob(self) -> np.ndarray:
        n = C.c_size_t()
        self._check(self.lib.nvb_setup_blob_size(self.handle, C.byref(n)), "nvb_setup_blob_size")
        buf = np.zeros(n.value, np.uint8)
        self._check(self.lib.nvb_setup_blob_export(self.handle, buf.ctypes.data, buf.size), "nvb_setup_blob_export")
        return buf

    def import_blob(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, np.uint8)
        self._check(self.lib.nvb_setup_blob_import(self.handle, blob.ctypes.data, blob.size), "nvb_setup_blob_import")
        self.channels = int(np.frombuffer(blob[16:20].tobytes(), "<i4")[0])

    @property
    def floor0_stride(self) -> int:
        r = self.lib.nvb_floor0_stride(self.handle)
        if r < 0:
            self._check(r, "nvb_floor0_stride")
        return int(r)

    @property
    def post_stride(self) -> int:
        r = self.lib.nvb_post_stride(self.handle)
        if r < 0:
            self._check(r, "nvb_post_stride")
        return r

    def reset(self):
        self._check(self.lib.nvb_reset(self.handle), "nvb_reset")

    def decode_batch(self, batch: HostBatch, flags: int = RUN_DEFAULT, out: np.ndarray | None = None, cap_samples: int | None = None):
        """nvb_decode_batch with host buffers.  Returns (interleaved pcm view -- float32, or int16 with RUN_PCM_S16 --, Result)."""
        if flags & RUN_DEVICE_OUT:
            raise ValueError("RUN_DEVICE_OUT needs a device pointer: use decode_batch_ptr / decode_batch_begin")
        if out is None:
            cap = int(cap_samples) if cap_samples is not None else sum_output_bound(batch.frames)
            out = np.empty(max(cap * self.channels, 1), np.int16 if flags & RUN_PCM_S16 else np.float32)
        elif out.dtype != (np.int16 if flags & RUN_PCM_S16 else np.float32):
            raise ValueError("out must be int16 with RUN_PCM_S16, float32 otherwise")
        r = ResultStruct()
        rc = self.lib.nvb_decode_batch(self.handle, C.byref(batch.struct), flags, out.ctypes.data, out.size, C.byref(r))
        self._check(rc, "nvb_decode_batch")
        return out[: int(r.samples_per_channel) * self.channels], _result(r)

    def decode_batch_begin(self, batch: HostBatch, flags: int, out_ptr: int, out_floats: int) -> None:
        """nvb_decode_batch_begin: enqueue H2D + synthesis + D2H and return (up to MAX_IN_FLIGHT = 3 batches in flight; the batch's host
        buffers must stay alive and untouched until its decode_batch_end)."""
        self._check(self.lib.nvb_decode_batch_begin(self.handle, C.byref(batch.struct), flags, out_ptr, out_floats), "nvb_decode_batch_begin")

    def decode_batch_end(self) -> Result:
        """nvb_decode_batch_end: wait for the oldest batch in flight."""
        r = ResultStruct()
        self._check(self.lib.nvb_decode_batch_end(self.handle, C.byref(r)), "nvb_decode_batch_end")
        return _result(r)

    def decode_batch_ptr(self, batch: HostBatch, flags: int, out_ptr: int, out_floats: int) -> Result:
        r = ResultStruct()
        self._check(self.lib.nvb_decode_batch(self.handle, C.byref(batch.struct), flags, out_ptr, out_floats, C.byref(r)), "nvb_decode_batch")
        return _result(r)

    # ---- GPU-side packet unpack -------------------------------------------------------------------------------------
    def upload_unpack_tables(self, blob: np.ndarray):
        """nvb_upload_unpack_tables: the blob comes from the host half (hostlib.HostStream.unpack_tables)."""
        blob = np.ascontiguousarray(blob, np.uint8)
        self._check(self.lib.nvb_upload_unpack_tables(self.handle, blob.ctypes.data, blob.size), "nvb_upload_unpack_tables")

    def unpack_strides(self):
        c, e = C.c_int32(), C.c_int32()
        self._check(self.lib.nvb_unpack_strides(self.handle, C.byref(c), C.byref(e)), "nvb_unpack_strides")
        return int(c.value), int(e.value)

    def decode_packets(self, pb: "PacketBatch", flags: int = RUN_DEFAULT, out: np.ndarray | None = None):
        """nvb_decode_packets: raw packets in, interleaved PCM out (float32, or int16 with RUN_PCM_S16)."""
        if out is None:
            out = np.empty(max(sum_output_bound(pb.frames) * self.channels, 1), np.int16 if flags & RUN_PCM_S16 else np.float32)
        r = ResultStruct()
        self._check(self.lib.nvb_decode_packets(self.handle, C.byref(pb.struct), flags, out.ctypes.data, out.size, C.byref(r)), "nvb_decode_packets")
        return out[: int(r.samples_per_channel) * self.channels], _result(r)

    def decode_packets_begin(self, pb: "PacketBatch", flags: int, out_ptr: int, out_elems: int) -> None:
        self._check(self.lib.nvb_decode_packets_begin(self.handle, C.byref(pb.struct), flags, out_ptr, out_elems), "nvb_decode_packets_begin")

    def unpack_packets(self, pb: "PacketBatch", post_stride: int) -> HostBatch:
        """nvb_unpack_packets: the device-produced boundary records, compacted into the layout nvh_unpack uses (for parity tests)."""
        cs, es = self.unpack_strides()
        n = len(pb.frames)
        frames = np.zeros(n, FRAME_DTYPE); posts = np.zeros(max(n * self.channels * post_stride, 1), np.int16)
        classes = np.zeros(max(n * cs, 1), np.uint8); entries = np.zeros(max(n * es, 1), np.uint16)
        self._check(self.lib.nvb_unpack_packets(self.handle, C.byref(pb.struct), frames.ctypes.data, posts.ctypes.data, classes.ctypes.data, entries.ctypes.data), "nvb_unpack_packets")
        return frames, posts[: n * self.channels * post_stride], classes.reshape(n, cs) if n else classes[:0], entries.reshape(n, es) if n else entries[:0]

    def create_dbatch(self, batch: HostBatch, flags: int = RUN_DEFAULT) -> DeviceBatch:
        h = C.c_void_p()
        self._check(self.lib.nvb_dbatch_create(self.handle, C.byref(batch.struct), flags, C.byref(h)), "nvb_dbatch_create")
        return DeviceBatch(self, h)


def sum_output_bound(frames: np.ndarray) -> int:
    """Upper bound of the samples per channel a batch can emit (every block's full length)."""
    if len(frames) == 0:
        return 0
    return int(np.maximum(frames["total"].astype(np.int64), 0).sum()) + 8192
