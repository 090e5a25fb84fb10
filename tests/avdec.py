"""Independent conformant H.264 decoder for the tests: FFmpeg's native `h264` decoder inside the
libavcodec that ships with opencv-python-headless, driven through ctypes.

Returns full Y, U and V planes, so the test can demand that decoding the encoder's output
reproduces the encoder's own reconstruction bit-exactly (north star, correctness part 3).
"""
import ctypes as C
import glob
import os

import numpy as np

_libs = None


def _load():
    global _libs
    if _libs is not None:
        return _libs
    import cv2  # noqa: F401  -- makes the bundled libs' dependencies resolvable
    base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    avutil = C.CDLL(glob.glob(os.path.join(base, "libavutil-*.so*"))[0], mode=C.RTLD_GLOBAL)
    for dep in ("libswresample-*", "libvpx-*", "libaom-*"):
        for p in glob.glob(os.path.join(base, dep + ".so*")):
            try:
                C.CDLL(p, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
    avcodec = C.CDLL(glob.glob(os.path.join(base, "libavcodec-*.so*"))[0], mode=C.RTLD_GLOBAL)
    avcodec.avcodec_find_decoder.restype = C.c_void_p
    avcodec.avcodec_find_decoder.argtypes = [C.c_int]
    avcodec.avcodec_alloc_context3.restype = C.c_void_p
    avcodec.avcodec_alloc_context3.argtypes = [C.c_void_p]
    avcodec.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    avcodec.av_packet_alloc.restype = C.c_void_p
    avcodec.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_free_context.argtypes = [C.c_void_p]
    avcodec.av_packet_free.argtypes = [C.c_void_p]
    avutil.av_frame_alloc.restype = C.c_void_p
    avutil.av_frame_free.argtypes = [C.c_void_p]
    avutil.av_log_set_level.argtypes = [C.c_int]
    _libs = (avcodec, avutil)
    return _libs


def available():
    try:
        _load()
        return True
    except Exception:
        return False


class _AVFrameHead(C.Structure):
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


class _AVPacketHead(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p),
                ("size", C.c_int)]


def split_nals(stream: bytes):
    """Split an Annex-B stream into NAL units (each including its own start code)."""
    pos = []
    i = 0
    n = len(stream)
    while True:
        j = stream.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        start = j - 1 if j > 0 and stream[j - 1] == 0 else j
        pos.append((start, j + 3))
        i = j + 3
    nals = []
    for k, (s, hdr) in enumerate(pos):
        end = pos[k + 1][0] if k + 1 < len(pos) else n
        nals.append((stream[hdr] & 0x1F, stream[s:end]))
    return nals


def split_access_units(stream: bytes):
    """One access unit = parameter sets (if any) + all slices of one picture.  A slice NAL whose
    first_mb_in_slice is 0 (first payload bit set: ue(0) = '1') starts a new picture."""
    aus, cur, has_slice = [], b"", False
    for t, nal in split_nals(stream):
        if t in (1, 5):
            hdr = nal.index(b"\x00\x00\x01") + 3
            if has_slice and (nal[hdr + 1] & 0x80):
                aus.append(cur)
                cur, has_slice = b"", False
            cur += nal
            has_slice = True
        else:
            if has_slice:
                aus.append(cur)
                cur, has_slice = b"", False
            cur += nal
    if cur:
        aus.append(cur)
    return aus


def decode(stream: bytes, quiet=True):
    """Decode an Annex-B stream; returns a list of (Y, U, V) uint8 numpy planes."""
    avcodec, avutil = _load()
    if quiet:
        avutil.av_log_set_level(16)  # AV_LOG_ERROR
    codec = avcodec.avcodec_find_decoder(27)  # AV_CODEC_ID_H264
    assert codec, "no h264 decoder in bundled libavcodec"
    ctx = avcodec.avcodec_alloc_context3(codec)
    assert avcodec.avcodec_open2(ctx, codec, None) == 0
    pkt = avcodec.av_packet_alloc()
    frame = avutil.av_frame_alloc()
    frames = []

    def drain():
        while avcodec.avcodec_receive_frame(ctx, frame) == 0:
            fh = _AVFrameHead.from_address(frame)
            w, h = fh.width, fh.height
            planes = []
            for p, (pw, ph) in enumerate(((w, h), (w // 2, h // 2), (w // 2, h // 2))):
                ls = fh.linesize[p]
                buf = (C.c_uint8 * (ls * ph)).from_address(fh.data[p])
                a = np.frombuffer(buf, dtype=np.uint8).reshape(ph, ls)[:, :pw].copy()
                planes.append(a)
            frames.append(tuple(planes))

    keep = []
    for au in split_access_units(stream):
        padded = au + b"\x00" * 64
        cbuf = C.create_string_buffer(padded, len(padded))
        keep.append(cbuf)
        ph = _AVPacketHead.from_address(pkt)
        ph.data = C.addressof(cbuf)
        ph.size = len(au)
        r = avcodec.avcodec_send_packet(ctx, pkt)
        if r != 0:
            raise RuntimeError("avcodec_send_packet failed: %d" % r)
        drain()
    ph = _AVPacketHead.from_address(pkt)
    ph.data = None
    ph.size = 0
    avcodec.avcodec_send_packet(ctx, None)
    drain()
    pp = C.c_void_p(pkt)
    avcodec.av_packet_free(C.byref(pp))
    fp = C.c_void_p(frame)
    avutil.av_frame_free(C.byref(fp))
    cp = C.c_void_p(ctx)
    avcodec.avcodec_free_context(C.byref(cp))
    return frames
