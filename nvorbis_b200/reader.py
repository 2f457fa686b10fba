"""VorbisReader / StreamDecoder mirror (VorbisReader.cs, StreamDecoder.cs) on top of the split decoder:
the host library unpacks batches of packets (libnvorbis_host.so), the GPU library synthesises them
(libnvorbis_b200.so), and ReadSamples is served from the decoded batch, as the batching StreamDecoder of
INTEGRATION.md does on the C# side.  One batch is always in flight (nvb_decode_batch_begin / _end): while the caller
consumes batch k, batch k+1 has already been unpacked and is on the GPU.  Same names, argument meaning and end-of-stream behaviour as the
reference's public API for this path; there is no CPU synthesis fallback.
"""
from __future__ import annotations

import numpy as np

from . import capi, hostlib


SEEK_BEGIN, SEEK_CURRENT, SEEK_END = 0, 1, 2          # System.IO.SeekOrigin


class VorbisReader:
    """`new VorbisReader(stream)` ... `ReadSamples(buffer, offset, count)` (VorbisReader.cs:42-64, 336-345)."""

    def __init__(self, source, device: int = 0, batch_packets: int = 4096, clip_samples: bool = True,
                 unpack_threads: int = 0, lib_path: str | None = None, gpu_unpack: bool | None = None, stream_index: int = 0):
        # stream_index: which logical stream of a multi-stream container (VorbisReader.Streams / SwitchStreams, VorbisReader.cs:96-150)
        self._container = None                  # the container's bytes (kept for SwitchStreams)
        if isinstance(source, (bytes, bytearray, memoryview)):
            self._container = bytes(source)
        elif isinstance(source, str):
            with open(source, "rb") as f:
                self._container = f.read()
        if self._container is not None:
            self._host = hostlib.HostStream(data=self._container, stream_index=stream_index)
        else:                                   # (data, sizes, granules, flags): an IPacketProvider's packets
            self._host = hostlib.HostStream(packets=tuple(source))
        self._stream_index = int(stream_index) if self._container is not None else 0
        self._ctx = capi.Context(device, lib_path=lib_path)
        self._want_gpu_unpack = gpu_unpack
        self._install_setup()
        self._batch_packets = int(batch_packets)
        self._skip = 0                                  # floats to drop after a seek (roll-forward, StreamDecoder.cs:626)
        self._total = None
        self._threads = unpack_threads
        self.clip_samples = bool(clip_samples)          # StreamDecoder.ClipSamples (VorbisReader.cs:77 forces true)
        self._pcm = np.zeros(0, np.float32)             # decoded, not yet handed out
        self._pos = 0
        self._eos = False
        self._started = False
        self._has_clipped = False
        self._samples_read = 0
        self._inflight = None                           # (HostBatch, pcm buffer) of the batch begun and not yet ended
        self._unpack_done = False                       # the host unpacker reached the end of the stream

    def _install_setup(self):
        """Uploads the current logical stream's setup (and, where wanted and covered, its unpack tables) to the GPU context."""
        self._ctx.upload_setup(self._host.setup())
        # GPU-side packet unpack (nvb_decode_packets): the host only pages the container and reads each packet's first bits;
        # None = use it whenever the setup is covered by the device tables, False = host unpacker (nvh_unpack) + nvb_decode_batch
        self._gpu_unpack = False
        if self._want_gpu_unpack is None or self._want_gpu_unpack:
            try:
                self._ctx.upload_unpack_tables(self._host.unpack_tables())
                self._gpu_unpack = True
            except (hostlib.HostError, capi.NvbError):
                if self._want_gpu_unpack:
                    raise

    # ---- logical streams of a multiplexed / chained container (VorbisReader.Streams, StreamIndex, SwitchStreams: VorbisReader.cs:116,184-190,291-305)
    @property
    def stream_count(self) -> int:
        return hostlib.ogg_stream_count(self._container) if self._container is not None else 1

    @property
    def stream_index(self) -> int:
        return self._stream_index

    def switch_streams(self, index: int) -> bool:
        """SwitchStreams(index): decode another logical stream from its start; True when its channel count or sample rate differs from
        the stream decoded before.  The clipping setting carries over (VorbisReader.cs:299-300)."""
        if index < 0 or index >= self.stream_count:
            raise IndexError("index")                                            # ArgumentOutOfRangeException
        if index == self._stream_index:
            return False
        old = (self.channels, self.sample_rate)
        if self._inflight is not None:
            self._ctx.decode_batch_end(); self._inflight = None
        self._host.close()
        self._host = hostlib.HostStream(data=self._container, stream_index=index)
        self._stream_index = index
        self._ctx.reset()
        self._install_setup()
        self._pcm, self._pos, self._eos, self._started = np.zeros(0, np.float32), 0, False, False
        self._samples_read, self._skip, self._unpack_done, self._total, self._has_clipped = 0, 0, False, None, False
        return (self.channels, self.sample_rate) != old

    # ---- properties of IVorbisReader / IStreamDecoder used by TestApp ---------------------------------
    @property
    def channels(self) -> int:
        return self._host.channels

    @property
    def sample_rate(self) -> int:
        return self._host.sample_rate

    @property
    def has_clipped(self) -> bool:
        return self._has_clipped

    @has_clipped.setter
    def has_clipped(self, v: bool):
        self._has_clipped = bool(v)

    @property
    def is_end_of_stream(self) -> bool:                 # StreamDecoder.IsEndOfStream
        return self._eos and self._pos >= self._pcm.size

    @property
    def sample_position(self) -> int:                   # VorbisReader.SamplePosition (VorbisReader.cs:237-244): get, or set = SeekTo(value)
        return self._samples_read

    @sample_position.setter
    def sample_position(self, value: int):
        self.seek_to(value)

    @property
    def time_position(self) -> float:                   # TimePosition in seconds (StreamDecoder.cs: currentPosition / sampleRate); set = SeekTo(TimeSpan)
        return self._samples_read / self.sample_rate

    @time_position.setter
    def time_position(self, seconds: float):
        self.seek_to_time(seconds)

    @property
    def total_time(self) -> float:                      # TotalTime in seconds (TotalSamples / SampleRate)
        return self.total_samples / self.sample_rate

    def _begin_next(self) -> bool:
        """Unpacks the next run of packets and puts it on the GPU; False when the stream has no more packets."""
        if self._unpack_done:
            return False
        flags = capi.RUN_DEFAULT | (capi.RUN_CONTINUE if self._started else 0) | (0 if self.clip_samples else capi.RUN_NO_CLIP)
        if self._gpu_unpack:
            hb, eos = self._host.packet_batch(self._batch_packets, copy=True)          # the arrays must outlive the call
            out = np.empty(max(capi.sum_output_bound(hb.frames) * self.channels, 1), np.float32)
            self._ctx.decode_packets_begin(hb, flags, out.ctypes.data, out.size)
        else:
            hb, eos = self._host.unpack(self._batch_packets, self._threads, copy=True)
            out = np.empty(max(capi.sum_output_bound(hb.frames) * self.channels, 1), np.float32)
            self._ctx.decode_batch_begin(hb, flags, out.ctypes.data, out.size)
        self._inflight = (hb, out)
        self._started = True
        self._unpack_done = eos
        return True

    def _refill(self) -> bool:
        if self._eos:
            return False
        if self._inflight is None and not self._begin_next():
            self._eos = True
            return False
        hb, out = self._inflight
        res = self._ctx.decode_batch_end()
        self._inflight = None
        self._has_clipped |= res.has_clipped
        self._pcm, self._pos = out[: res.samples_per_channel * self.channels], 0
        if self._skip:                                  # roll forward to the sample a seek asked for
            drop = min(self._skip, self._pcm.size)
            self._pos, self._skip = drop, self._skip - drop
        if not self._begin_next():                      # batch k+1 goes up while the caller consumes batch k
            self._eos = True
        return self._pcm.size > 0 or not self._eos

    def read_samples(self, buffer: np.ndarray, offset: int, count: int) -> int:
        """Fills buffer[offset : offset+count] with interleaved float PCM; returns the number of floats written
        (a multiple of Channels; 0 at the end of the stream)."""
        if buffer.dtype != np.float32:
            raise TypeError("buffer must be float32")
        if offset < 0 or offset + count > buffer.size:
            raise IndexError("offset/count outside the buffer")                  # ArgumentOutOfRangeException
        count -= count % self.channels                                           # VorbisReader.cs:339-340
        done = 0
        while done < count:
            if self._pos >= self._pcm.size and not self._refill():
                break
            n = min(count - done, self._pcm.size - self._pos)
            buffer[offset + done: offset + done + n] = self._pcm[self._pos: self._pos + n]
            self._pos += n; done += n
        self._samples_read += done // self.channels
        return done

    def read_all(self, chunk_seconds: float = 4.0) -> np.ndarray:
        """TestApp/Program.cs:21-26: 4-second chunks until ReadSamples returns 0."""
        buf = np.zeros(int(self.sample_rate * chunk_seconds) * self.channels, np.float32)
        out = []
        while True:
            n = self.read_samples(buf, 0, buf.size)
            if n <= 0:
                break
            out.append(buf[:n].copy())
        return np.concatenate(out) if out else np.zeros(0, np.float32)

    @property
    def total_samples(self) -> int:
        """TotalSamples: what a decode of the whole stream emits per channel (header-only walk, cached)."""
        if self._total is None:
            if self._inflight is not None or self._started:
                raise RuntimeError("total_samples must be read before decoding starts or after seek_to(0)")
            self._total = self._host.total_samples()
        return self._total

    def seek_to_time(self, seconds: float, origin: int = SEEK_BEGIN):
        """SeekTo(TimeSpan, SeekOrigin) (StreamDecoder.cs:551-554): (long)(SampleRate * TotalSeconds), then the sample form."""
        self.seek_to(int(self.sample_rate * float(seconds)), origin)

    def seek_to(self, sample_position: int, origin: int = SEEK_BEGIN):
        """SeekTo(samplePosition, seekOrigin) (StreamDecoder.cs:562-628): the next read_samples starts at that sample.  Decoding restarts
        one packet early (the pre-roll packet only leaves its overlap tail) and rolls forward inside the next block.  The origins
        follow the reference to the letter: Current means SamplePosition - samplePosition (StreamDecoder.cs:572), End means
        TotalSamples - samplePosition (:575)."""
        if origin == SEEK_CURRENT:
            sample_position = self.sample_position - sample_position
        elif origin == SEEK_END:
            sample_position = self.total_samples - sample_position
        elif origin != SEEK_BEGIN:
            raise IndexError("seekOrigin")                                       # ArgumentOutOfRangeException
        if sample_position < 0:
            raise IndexError("samplePosition")                                   # ArgumentOutOfRangeException
        if self._inflight is not None:
            self._ctx.decode_batch_end(); self._inflight = None
        skip = self._host.seek(int(sample_position))
        self._ctx.reset()
        self._pcm, self._pos, self._eos, self._started = np.zeros(0, np.float32), 0, False, False
        self._samples_read = int(sample_position)
        self._skip = skip * self.channels
        self._unpack_done = False

    def seek_to_start(self):
        """SeekTo(0): restart decoding at the first audio packet (the looping short cut, StreamDecoder.cs:588-592)."""
        self.seek_to(0)

    def close(self):
        if self._inflight is not None:
            try:
                self._ctx.decode_batch_end()
            finally:
                self._inflight = None
        self._ctx.close()
        self._host.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
