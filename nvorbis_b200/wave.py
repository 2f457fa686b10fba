"""WaveWriter mirror (TestApp/WaveWriter.cs:18-62): the only output format the reference ships -- a RIFF/WAVE file with an
18-byte `fmt ` chunk, encoding 3 (IEEE float), 32 bits per sample, filled by WriteSamples and patched on Dispose -- plus the
16-bit PCM form (encoding 1) for the output of NVB_RUN_PCM_S16.

One deliberate difference, switchable: the reference's Dispose writes the data-chunk size at file offset 44
(WaveWriter.cs:56-57), but with its 18-byte `fmt ` chunk that field sits at offset 42, so its files carry a wrong data size
and two clobbered bytes of the first sample.  `reference_quirk=True` reproduces that byte for byte; the default writes a
valid file."""
from __future__ import annotations

import struct

import numpy as np


class WaveWriter:
    def __init__(self, file_name: str, sample_rate: int, channels: int, sample_format: str = "f32", reference_quirk: bool = False):
        if sample_format not in ("f32", "s16"):
            raise ValueError("sample_format must be 'f32' or 's16'")
        self._f = open(file_name, "wb")
        self._fmt, self._quirk = sample_format, bool(reference_quirk)
        bits = 32 if sample_format == "f32" else 16
        block_align = channels * bits // 8
        self._f.write(b"RIFF\0\0\0\0WAVEfmt ")                                  # BLANK_HEADER, WaveWriter.cs:11
        self._f.write(struct.pack("<ihhiihhh", 18, 3 if sample_format == "f32" else 1, channels, sample_rate,
                                  block_align * sample_rate, block_align, bits, 0))   # WaveWriter.cs:24-41
        self._f.write(b"data\0\0\0\0")                                           # BLANK_DATA_HEADER
        self._data_size_at = self._f.tell() - 4                                  # 42

    def write_samples(self, buf: np.ndarray, offset: int, count: int):
        """WriteSamples(buf, offset, count) (WaveWriter.cs:46-52)."""
        want = np.float32 if self._fmt == "f32" else np.int16
        a = np.ascontiguousarray(buf[offset: offset + count])
        if a.dtype != want:
            raise TypeError(f"buffer must be {np.dtype(want).name}")
        self._f.write(a.astype("<" + ("f4" if self._fmt == "f32" else "i2"), copy=False).tobytes())

    def close(self):
        """Dispose (WaveWriter.cs:54-69): RIFF chunk size = length - 8, data chunk size = length - 48 (sic; kept)."""
        if self._f is None:
            return
        length = self._f.tell()
        self._f.seek(4); self._f.write(struct.pack("<I", length - 8))
        if self._quirk:
            self._f.seek(44); self._f.write(struct.pack("<I", length - 48))
        else:
            self._f.seek(self._data_size_at); self._f.write(struct.pack("<I", length - (self._data_size_at + 4)))
        self._f.close(); self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def decode_to_wav(ogg_source, wav_file: str, sample_format: str = "f32", **reader_kw) -> int:
    """TestApp/Program.cs:12-28: VorbisReader.ReadSamples in 4-second chunks into a WaveWriter.  Returns samples per channel."""
    from .reader import VorbisReader
    with VorbisReader(ogg_source, **reader_kw) as vr, WaveWriter(wav_file, vr.sample_rate, vr.channels, sample_format) as ww:
        buf = np.zeros(vr.sample_rate * vr.channels * 4, np.float32)
        total = 0
        while True:
            cnt = vr.read_samples(buf, 0, buf.size)
            if cnt <= 0:
                break
            if sample_format == "s16":          # host-side form of the device rule (round to nearest even of v * 32768, saturated)
                ww.write_samples(np.clip(np.rint(buf[:cnt] * np.float32(32768.0)), -32768, 32767).astype(np.int16), 0, cnt)
            else:
                ww.write_samples(buf, 0, cnt)
            total += cnt // vr.channels
        return total
