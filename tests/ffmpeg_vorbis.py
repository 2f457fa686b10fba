"""An INDEPENDENT Vorbis decoder for cross-checking the oracle: FFmpeg's native vorbis decoder inside the libavcodec that
ships with opencv-python-headless, driven through ctypes on already-demuxed packets (no libavformat).  Test infrastructure
only; returns None when the libraries are not there.  It proves Vorbis-correctness of the oracle (agreement to float
rounding), not bit parity with NVorbis -- FFmpeg neither clips nor trims the end of the stream."""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np


def _libs():
    try:
        import cv2
    except Exception:
        return None
    d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    out = {}
    for name in ("avutil", "swresample", "avcodec"):
        hits = sorted(glob.glob(os.path.join(d, f"lib{name}-*.so*")))
        if not hits:
            return None
        try:
            out[name] = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
        except OSError:
            return None
    return out


def decode(packets: list[bytes]) -> np.ndarray | None:
    """packets: the three Vorbis header packets followed by the audio packets.  Returns float32 [samples, channels]
    (un-clipped), or None if FFmpeg is unavailable / refuses the stream."""
    L = _libs()
    if L is None:
        return None
    au, ac = L["avutil"], L["avcodec"]
    vp, i32 = C.c_void_p, C.c_int
    ac.avcodec_find_decoder_by_name.restype = vp; ac.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    ac.avcodec_alloc_context3.restype = vp; ac.avcodec_alloc_context3.argtypes = [vp]
    ac.avcodec_parameters_alloc.restype = vp
    ac.avcodec_parameters_to_context.argtypes = [vp, vp]
    ac.avcodec_open2.argtypes = [vp, vp, vp]
    ac.av_packet_alloc.restype = vp
    ac.av_new_packet.argtypes = [vp, i32]
    ac.av_packet_unref.argtypes = [vp]
    ac.avcodec_send_packet.argtypes = [vp, vp]
    ac.avcodec_receive_frame.argtypes = [vp, vp]
    au.av_frame_alloc.restype = vp
    au.av_frame_unref.argtypes = [vp]
    au.av_mallocz.restype = vp; au.av_mallocz.argtypes = [C.c_size_t]
    codec = ac.avcodec_find_decoder_by_name(b"vorbis")
    if not codec:
        return None
    ctx = ac.avcodec_alloc_context3(codec)
    par = ac.avcodec_parameters_alloc()
    if not ctx or not par:
        return None
    # extradata in Xiph lacing: 0x02, sizes of the first two headers (255-laced), then the three headers
    def lace(n):
        return bytes([255] * (n // 255) + [n % 255])
    extra = bytes([2]) + lace(len(packets[0])) + lace(len(packets[1])) + b"".join(packets[:3])
    buf = au.av_mallocz(len(extra) + 64)
    C.memmove(buf, extra, len(extra))
    # AVCodecParameters: codec_type @0, codec_id @4, codec_tag @8, extradata @16, extradata_size @24
    C.c_int.from_address(par + 0).value = 1                                   # AVMEDIA_TYPE_AUDIO
    C.c_int.from_address(par + 4).value = C.c_int.from_address(codec + 20).value   # AVCodec.id (name, long_name, type, id)
    C.c_void_p.from_address(par + 16).value = buf
    C.c_int.from_address(par + 24).value = len(extra)
    if ac.avcodec_parameters_to_context(ctx, par) < 0 or ac.avcodec_open2(ctx, codec, None) < 0:
        return None
    pkt, frame = ac.av_packet_alloc(), au.av_frame_alloc()
    out = []
    for p in packets[3:]:
        if len(p) == 0:
            continue
        if ac.av_new_packet(pkt, len(p)) < 0:
            return None
        C.memmove(C.c_void_p.from_address(pkt + 24).value, p, len(p))        # AVPacket.data @24
        r = ac.avcodec_send_packet(ctx, pkt)
        ac.av_packet_unref(pkt)
        if r < 0:
            continue
        while ac.avcodec_receive_frame(ctx, frame) >= 0:
            n = C.c_int.from_address(frame + 112).value                       # AVFrame.nb_samples
            ext = C.c_void_p.from_address(frame + 96).value                   # AVFrame.extended_data (planar float)
            chans = []
            c = 0
            while c < 8:
                ptr = C.c_void_p.from_address(ext + 8 * c).value
                if not ptr:
                    break
                chans.append(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,)).copy())
                c += 1
            if n > 0 and chans:
                out.append(np.stack(chans, axis=1))
            au.av_frame_unref(frame)
    return np.concatenate(out) if out else np.zeros((0, 1), np.float32)
