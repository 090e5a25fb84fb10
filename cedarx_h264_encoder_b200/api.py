"""ctypes binding of include/cedar_b200.h (the drop-in boundary).

Field names follow struct cedar_ioctl_config of the reference (kernel/cedar_ioctl.h:12-46); the
call flow follows userspace/h264enc.c: config -> fill input buffers -> encode -> read bytestream.
"""
import ctypes as C
import os
import subprocess

# The serial CABAC stages of successive frames overlap on side streams, which need their own hardware queues; the variable
# only counts before CUDA initialises in the process, and the library leaves its host's environment alone.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libcedar_b200.so"

FORMAT_NV12, FORMAT_NV16 = 0, 1
ENTROPY_CAVLC, ENTROPY_CABAC = 0, 1


class LibraryMissing(RuntimeError):
    pass


class CedarConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "src_width", "src_height", "src_format", "dst_width", "dst_height", "profile", "level", "qp",
        "keyframe_interval", "thumbnail", "thumbnail_downscale", "entropy_coding_mode",
        "me_range", "relax_gop", "device", "gops_in_flight", "max_clip_frames", "slice_rows", "sps_crop", "auto_level", "repeat_headers", "intra4x4", "p_intra", "queue_gops")]


class CedarIO(C.Structure):
    _fields_ = [("input_luma", C.c_void_p), ("input_luma_size", C.c_int), ("input_chroma", C.c_void_p),
                ("input_chroma_size", C.c_int), ("bytestream", C.c_void_p), ("bytestream_size", C.c_int)]


MBINFO_DTYPE = np.dtype([("type", "u1"), ("i16_mode", "u1"), ("chroma_mode", "u1"), ("cbp", "u1"),
                         ("mv", "<i2", (2,)), ("mvd", "<i2", (2,)), ("pad", "<u4")])

_lib = None


def library_path():
    # CEDAR_B200_LIB: development aid (an instrumented build of the same sources, e.g. tools/libcedar_prof.so)
    return os.environ.get("CEDAR_B200_LIB") or os.path.join(PKG_DIR, LIB_NAME)


def build_library(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a build of the library and the CLI (in-tree)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", PKG_DIR, "clean"])
    subprocess.check_call(["make", "-s", "-C", PKG_DIR, "all"])
    return library_path()


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissing(
            "%s is not built; run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C %s`. "
            "There is no CPU fallback." % (path, PKG_DIR))
    L = C.CDLL(path)
    H = C.c_void_p
    L.cedar_b200_open.argtypes = [C.POINTER(CedarConfig), C.POINTER(CedarIO), C.POINTER(H)]
    L.cedar_b200_encode_frame.argtypes = [H]
    L.cedar_b200_flush.argtypes = [H]
    L.cedar_b200_close.argtypes = [H]
    L.cedar_b200_close.restype = None
    L.cedar_b200_pipe_open.argtypes = [C.POINTER(CedarConfig), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    L.cedar_b200_pipe_workers.argtypes = [H]
    L.cedar_b200_pipe_acquire.argtypes = [H, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.cedar_b200_pipe_acquire.restype = C.c_void_p
    L.cedar_b200_pipe_submit.argtypes = [H, C.c_int]
    L.cedar_b200_pipe_finish.argtypes = [H]
    L.cedar_b200_pipe_next.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int),
                                       C.POINTER(C.POINTER(C.c_double)), C.c_int]
    L.cedar_b200_pipe_next.restype = C.c_longlong
    L.cedar_b200_pipe_release.argtypes = [H]
    L.cedar_b200_pipe_close.argtypes = [H]
    L.cedar_b200_pipe_close.restype = None
    L.cedar_b200_clip_input.argtypes = [H, C.POINTER(C.c_size_t)]
    L.cedar_b200_clip_input.restype = C.c_void_p
    L.cedar_b200_clip_upload.argtypes = [H, C.c_int]
    L.cedar_b200_clip_encode.argtypes = [H, C.c_int, C.c_int]
    L.cedar_b200_clip_download.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.cedar_b200_clip_download.restype = C.c_longlong
    L.cedar_b200_stats.argtypes = [H, C.POINTER(C.c_double), C.c_int]
    L.cedar_b200_profile_enable.argtypes = [H, C.c_int]
    L.cedar_b200_profile_read.argtypes = [H, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                          C.c_int, C.c_int]
    L.cedar_b200_launch_count.argtypes = [H]
    L.cedar_b200_launch_count.restype = C.c_longlong
    L.cedar_b200_stream.argtypes = [H]
    L.cedar_b200_stream.restype = C.c_void_p
    L.cedar_b200_debug_read.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    L.cedar_b200_debug_read.restype = C.c_longlong
    L.cedar_b200_write_sps.argtypes = [C.POINTER(CedarConfig), C.c_void_p, C.c_int]
    L.cedar_b200_write_pps.argtypes = [C.POINTER(CedarConfig), C.c_void_p, C.c_int]
    L.cedar_b200_slice_header.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.cedar_b200_slice_header_mb.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64),
                                             C.POINTER(C.c_int)]
    L.cedar_b200_version.restype = C.c_char_p
    _lib = L
    return L


def align16(x):
    return (x + 15) & ~15


def make_config(width, height, qp=24, gop=25, cabac=1, fmt=FORMAT_NV12, me_range=16, profile=77, level=41,
                dst_width=None, dst_height=None, relax_gop=1, device=0, gops_in_flight=0, max_clip_frames=0,
                slice_rows=0, sps_crop=0, auto_level=0, repeat_headers=0, intra4x4=0, p_intra=0, queue_gops=0):
    """Defaults are the reference's hard-coded ones (userspace/h264enc.c:53-66)."""
    return CedarConfig(width, height, fmt, align16(width) if dst_width is None else dst_width,
                       align16(height) if dst_height is None else dst_height, profile, level, qp, gop, 0, 0,
                       cabac, me_range, relax_gop, device, gops_in_flight, max_clip_frames, slice_rows, sps_crop, auto_level, repeat_headers, intra4x4, p_intra, queue_gops)


def write_sps(cfg):
    buf = np.zeros(64, np.uint8)
    n = load_library().cedar_b200_write_sps(C.byref(cfg), buf.ctypes.data, 64)
    return buf[:n].tobytes()


def write_pps(cfg):
    buf = np.zeros(64, np.uint8)
    n = load_library().cedar_b200_write_pps(C.byref(cfg), buf.ctypes.data, 64)
    return buf[:n].tobytes()


def slice_header_bits(frame_i, frame_p_count, cabac):
    bits, n = C.c_uint32(), C.c_int()
    load_library().cedar_b200_slice_header(frame_i, frame_p_count, cabac, C.byref(bits), C.byref(n))
    return format(bits.value, "0%db" % n.value)


class Encoder:
    """open / encode_frame / close, mirroring the reference's config / encode / release flow."""

    def __init__(self, cfg: CedarConfig):
        self.L = load_library()
        self.cfg = cfg
        self.io = CedarIO()
        self.h = C.c_void_p()
        r = self.L.cedar_b200_open(C.byref(cfg), C.byref(self.io), C.byref(self.h))
        if r:
            self.h = C.c_void_p()
            raise OSError(-r, "cedar_b200_open failed: %s" % os.strerror(-r))
        self.W, self.H = cfg.dst_width, cfg.dst_height
        w, h = cfg.src_width, cfg.src_height
        self.luma_bytes = w * h
        self.chroma_bytes = w * h if cfg.src_format == FORMAT_NV16 else w * h // 2
        self._luma = np.ctypeslib.as_array(C.cast(self.io.input_luma, C.POINTER(C.c_uint8)), shape=(self.luma_bytes,))
        self._chroma = np.ctypeslib.as_array(C.cast(self.io.input_chroma, C.POINTER(C.c_uint8)),
                                             shape=(self.chroma_bytes,))
        self._bs = np.ctypeslib.as_array(C.cast(self.io.bytestream, C.POINTER(C.c_uint8)),
                                         shape=(self.io.bytestream_size,))
        self.frame_bytes = self.luma_bytes + self.chroma_bytes

    # ---- frame mode (ioctl(CEDAR_IOCTL_ENCODE) semantics) ----
    def encode(self, luma, chroma) -> bytes:
        self._luma[:] = np.asarray(luma, np.uint8).reshape(-1)
        self._chroma[:] = np.asarray(chroma, np.uint8).reshape(-1)
        n = self.L.cedar_b200_encode_frame(self.h)
        if n < 0:
            raise OSError(-n, "cedar_b200_encode_frame failed: %s" % os.strerror(-n))
        return self._bs[:n].tobytes()

    def flush(self) -> bytes:
        """Queued mode (queue_gops): the next outstanding frame, b"" when drained."""
        n = self.L.cedar_b200_flush(self.h)
        if n < 0:
            raise OSError(-n, "cedar_b200_flush failed: %s" % os.strerror(-n))
        return self._bs[:n].tobytes()

    # ---- clip mode (GOP-parallel) ----
    def clip_input(self, nframes):
        fb = C.c_size_t()
        p = self.L.cedar_b200_clip_input(self.h, C.byref(fb))
        if not p:
            raise RuntimeError("clip mode is off (max_clip_frames == 0)")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nframes, fb.value))

    def clip_upload(self, nframes):
        self._ck(self.L.cedar_b200_clip_upload(self.h, nframes), "clip_upload")

    def clip_encode(self, nframes, first_frame_index=0):
        self._ck(self.L.cedar_b200_clip_encode(self.h, nframes, first_frame_index), "clip_encode")

    def clip_download(self, nframes):
        out = C.c_void_p()
        sizes = (C.c_int * nframes)()
        n = self.L.cedar_b200_clip_download(self.h, C.byref(out), sizes)
        if n < 0:
            raise OSError(-n, "cedar_b200_clip_download failed: %s" % os.strerror(-n))
        data = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(n,))
        return data, np.frombuffer(sizes, dtype=np.int32).copy()

    def encode_clip(self, frames, first_frame_index=0):
        """frames: uint8 array [n, frame_bytes] (luma then chroma per frame).  Returns (bytes, sizes)."""
        n = len(frames)
        self.clip_input(n)[:] = frames
        self.clip_upload(n)
        self.clip_encode(n, first_frame_index)
        data, sizes = self.clip_download(n)
        return data.tobytes(), sizes

    # ---- statistics / profiling / debug ----
    def sse_y(self, nframes=1):
        a = (C.c_double * nframes)()
        self._ck(self.L.cedar_b200_stats(self.h, a, nframes), "stats")
        return np.array(a[:])

    def profile_enable(self, mode=1):
        """0 off, 1 live (overlapped streams), 2 serialised (standalone kernel times)."""
        self.L.cedar_b200_profile_enable(self.h, int(mode))

    def profile_read(self, reset=True):
        cap = 32
        names, ms, n = (C.c_char_p * cap)(), (C.c_float * cap)(), (C.c_int * cap)()
        k = self.L.cedar_b200_profile_read(self.h, names, ms, n, cap, int(reset))
        return {names[i].decode(): (ms[i], n[i]) for i in range(k)}

    def stream_ptr(self):
        return self.L.cedar_b200_stream(self.h)

    def launch_count(self):
        return self.L.cedar_b200_launch_count(self.h)

    def debug_planes(self, what):
        buf = np.empty(self.W * self.H * 3 // 2, np.uint8)
        self._ck(self.L.cedar_b200_debug_read(self.h, what, buf.ctypes.data, buf.size), "debug_read")
        ys = self.W * self.H
        cs = ys // 4
        return (buf[:ys].reshape(self.H, self.W), buf[ys:ys + cs].reshape(self.H // 2, self.W // 2),
                buf[ys + cs:].reshape(self.H // 2, self.W // 2))

    def debug_syntax(self):
        nmb = (self.W // 16) * (self.H // 16)
        mbi = np.empty(nmb, MBINFO_DTYPE)
        nnz = np.empty((nmb, 32), np.uint8)
        coef = np.empty((nmb, 26, 16), np.int16)
        self._ck(self.L.cedar_b200_debug_read(self.h, 3, mbi.ctypes.data, mbi.nbytes), "debug_read")
        self._ck(self.L.cedar_b200_debug_read(self.h, 4, nnz.ctypes.data, nnz.nbytes), "debug_read")
        self._ck(self.L.cedar_b200_debug_read(self.h, 5, coef.ctypes.data, coef.nbytes), "debug_read")
        return mbi, nnz, coef

    def debug_i4_modes(self):
        nmb = (self.W // 16) * (self.H // 16)
        modes = np.empty((nmb, 16), np.uint8)
        self._ck(self.L.cedar_b200_debug_read(self.h, 7, modes.ctypes.data, modes.nbytes), "debug_read")
        return modes

    def _ck(self, r, what):
        if r < 0:
            raise OSError(-r, "cedar_b200_%s failed: %s" % (what, os.strerror(-r)))

    def close(self):
        if self.h:
            self._luma = self._chroma = self._bs = None
            self.L.cedar_b200_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipe:
    """cedar_b200_pipe_*: ordered pipeline over several handles / GPUs (csrc/pipeline.cpp)."""

    def __init__(self, cfg: CedarConfig, devices=None, handles_per_device=0, gops_per_batch=0):
        self.L = load_library()
        self.p = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices) if devices else None
        r = self.L.cedar_b200_pipe_open(C.byref(cfg), devs, len(devices) if devices else 0, handles_per_device, gops_per_batch,
                                        C.byref(self.p))
        if r:
            self.p = C.c_void_p()
            raise OSError(-r, "cedar_b200_pipe_open failed: %s" % os.strerror(-r))

    def workers(self):
        return self.L.cedar_b200_pipe_workers(self.p)

    def acquire(self):
        """numpy view [capacity_frames, frame_bytes] of the staging buffer of the next batch."""
        fb, cap = C.c_size_t(), C.c_int()
        ptr = self.L.cedar_b200_pipe_acquire(self.p, C.byref(fb), C.byref(cap))
        if not ptr:
            raise RuntimeError("pipe_acquire: a batch is already being filled, or the stream is finished")
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(cap.value, fb.value))

    def submit(self, nframes):
        r = self.L.cedar_b200_pipe_submit(self.p, nframes)
        if r:
            raise OSError(-r, "cedar_b200_pipe_submit failed: %s" % os.strerror(-r))

    def finish(self):
        self.L.cedar_b200_pipe_finish(self.p)

    def next(self, wait=True):
        """(bytes, sizes, sse) of the next batch in submission order; None when nothing is outstanding."""
        out, sizes, sse, n = C.c_void_p(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)(), C.c_int()
        total = self.L.cedar_b200_pipe_next(self.p, C.byref(out), C.byref(sizes), C.byref(n), C.byref(sse), int(wait))
        if total == 0:
            return None
        if total < 0:
            raise OSError(-total, "cedar_b200_pipe_next: %s" % os.strerror(-total))
        res = C.string_at(out, total), [sizes[i] for i in range(n.value)], [sse[i] for i in range(n.value)]
        self.L.cedar_b200_pipe_release(self.p)  # everything is copied: the batch's worker may go on
        return res

    def encode(self, frames):
        """frames: uint8 [n, frame_bytes].  Single-threaded driver: keeps every worker busy, returns (bytes, sizes)."""
        n, done, out, sizes, inflight = len(frames), 0, [], [], 0
        W = self.workers()
        while done < n or inflight:
            while done < n and inflight < W:
                buf = self.acquire()
                k = min(len(buf), n - done)
                buf[:k] = frames[done:done + k]
                self.submit(k)
                done += k
                inflight += 1
            got = self.next(True)
            inflight -= 1
            out.append(got[0])
            sizes += got[1]
        self.finish()
        return b"".join(out), sizes

    def close(self):
        if self.p:
            self.L.cedar_b200_pipe_close(self.p)
            self.p = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
