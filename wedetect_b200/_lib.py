"""ctypes binding of libwedetect_b200.so (the C ABI declared in include/wedetect_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present the
product path raises.  (Tests that only need the symbol table can call `load(require_gpu=False)`.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WD_LIB_PATH") or os.path.join(_HERE, "libwedetect_b200.so")   # WD_LIB_PATH: A/B builds of the library

WD_OP_NI, WD_OP_NF, WD_OP_NP = 48, 8, 16
ACT_PLANE_SCALE = 4.0    # WD_ACT_PLANE_SCALE of include/wedetect_b200.h

# enum wd_op_kind
OP_GEMM, OP_LN_ROWS, OP_DWCONV_LN, OP_STEM_PATCH, OP_IM2COL_S2, OP_CAST_BF16 = 1, 2, 3, 4, 5, 6
OP_TEXT_EMBED, OP_ATTN_SMALL, OP_L2NORM_ROWS, OP_GATHER_ROWS, OP_FOLD_TEXT, OP_POSTPROCESS, OP_GATHER_EMBED = 7, 8, 9, 10, 11, 12, 13
OP_SCALE_ROWS, OP_RETR_REDUCE, OP_LETTERBOX, OP_MLP_FUSED, OP_CV_RESIZE_PAD = 14, 15, 16, 17, 18
ACT_NONE, ACT_RELU, ACT_SILU, ACT_GELU = 0, 1, 2, 3

EXPORTS = [
    "wd_last_error", "wd_version", "wd_launch_count", "wd_device_info", "wd_op_run", "wd_program_create",
    "wd_program_run", "wd_program_capture", "wd_program_replay", "wd_program_num_launches",
    "wd_program_destroy", "wd_pp_workspace_bytes", "wd_program_num_ops", "wd_program_run_timed",
    "wd_program_find_stuck_op", "wd_act_plane_scale", "wd_jpeg_open", "wd_jpeg_info", "wd_jpeg_decode", "wd_jpeg_close",
]


class WdOp(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("i", ctypes.c_int32 * WD_OP_NI),
        ("f", ctypes.c_float * WD_OP_NF),
        ("p", ctypes.c_void_p * WD_OP_NP),
    ]


class PPParams(ctypes.Structure):
    _fields_ = [
        ("B", ctypes.c_int32), ("K", ctypes.c_int32), ("nlevels", ctypes.c_int32),
        ("lvl_h", ctypes.c_int32 * 4), ("lvl_w", ctypes.c_int32 * 4), ("lvl_stride", ctypes.c_int32 * 4),
        ("ld_logit", ctypes.c_int32 * 4),
        ("logits", ctypes.c_void_p * 4), ("dist", ctypes.c_void_p * 4),
        ("score_thr", ctypes.c_float), ("nms_pre", ctypes.c_int32), ("iou_thr", ctypes.c_float),
        ("max_per_img", ctypes.c_int32), ("nms_mode", ctypes.c_int32), ("tv_numel_thr", ctypes.c_int32),
        ("multi_label", ctypes.c_int32),
        ("img_meta", ctypes.c_void_p), ("clamp_wh", ctypes.c_void_p),
        ("out_boxes", ctypes.c_void_p), ("out_scores", ctypes.c_void_p), ("out_labels", ctypes.c_void_p),
        ("out_anchor", ctypes.c_void_p), ("out_counts", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_uint64),
    ]


class WdError(RuntimeError):
    pass


_lib = None


def load(require_gpu=True, device=0):
    """Load the shared library (building it is `__graft_entry__.build()`'s job, not ours).  `device`: the ordinal to validate."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WdError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.wd_last_error.restype = ctypes.c_char_p
        lib.wd_version.restype = ctypes.c_int
        lib.wd_act_plane_scale.restype = ctypes.c_float
        lib.wd_launch_count.restype = ctypes.c_uint64
        lib.wd_device_info.argtypes = [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 3
        lib.wd_op_run.argtypes = [ctypes.POINTER(WdOp), ctypes.c_void_p]
        lib.wd_program_create.argtypes = [ctypes.POINTER(WdOp), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.wd_program_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.wd_program_capture.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.wd_program_replay.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.wd_program_num_launches.argtypes = [ctypes.c_void_p]
        lib.wd_program_num_ops.argtypes = [ctypes.c_void_p]
        lib.wd_program_run_timed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        lib.wd_program_find_stuck_op.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        lib.wd_program_destroy.argtypes = [ctypes.c_void_p]
        lib.wd_program_destroy.restype = None
        lib.wd_pp_workspace_bytes.argtypes = [ctypes.c_int] * 4
        lib.wd_pp_workspace_bytes.restype = ctypes.c_uint64
        lib.wd_jpeg_open.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
        lib.wd_jpeg_info.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64] + [ctypes.POINTER(ctypes.c_int)] * 4
        lib.wd_jpeg_decode.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
        lib.wd_jpeg_close.argtypes = [ctypes.c_void_p]
        lib.wd_jpeg_close.restype = None
        _lib = lib
    if require_gpu:
        sm, major, minor = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        rc = _lib.wd_device_info(int(device), ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor))
        if rc != 0:
            raise WdError("no CUDA device for libwedetect_b200: " + last_error())
        if major.value != 10:
            raise WdError(f"libwedetect_b200 is built for sm_100a only; found sm_{major.value}{minor.value}")
    return _lib


def last_error():
    return (_lib.wd_last_error() or b"").decode(errors="replace") if _lib is not None else ""


def check(rc, what):
    if rc != 0:
        raise WdError(f"{what} failed ({rc}): {last_error()}")


def launch_count():
    return int(load(require_gpu=False).wd_launch_count())


def run_op(op, stream=0):
    lib = load()
    check(lib.wd_op_run(ctypes.byref(op), ctypes.c_void_p(stream)), f"wd_op_run(kind={op.kind})")


class JpegDecoder:
    """nvJPEG decoder state on the current device (wd_jpeg_*): info() parses the header on the host, decode() writes
    interleaved pixels into device memory on a stream."""

    def __init__(self):
        lib = load()
        self._lib = lib
        h = ctypes.c_void_p()
        check(lib.wd_jpeg_open(ctypes.byref(h)), "wd_jpeg_open")
        self._h = h

    def info(self, data):
        w, h, nc, css = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        check(self._lib.wd_jpeg_info(self._h, data, len(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(nc), ctypes.byref(css)), "wd_jpeg_info")
        return w.value, h.value, nc.value, css.value

    def decode(self, data, dst_ptr, pitch, bgr=True, stream=0):
        check(self._lib.wd_jpeg_decode(self._h, data, len(data), ctypes.c_void_p(dst_ptr), pitch, 1 if bgr else 0, ctypes.c_void_p(stream)), "wd_jpeg_decode")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.wd_jpeg_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Program:
    """An op list compiled once (TMA descriptors prebuilt) and replayed per batch."""

    def __init__(self, ops, keepalive=()):
        lib = load()
        self._lib = lib
        self._keep = list(keepalive)
        arr = (WdOp * len(ops))(*ops)
        self._ops = arr
        h = ctypes.c_void_p()
        check(lib.wd_program_create(arr, len(ops), ctypes.byref(h)), "wd_program_create")
        self._h = h
        self._captured = False

    @property
    def num_launches(self):
        return int(self._lib.wd_program_num_launches(self._h))

    def run(self, stream=0):
        check(self._lib.wd_program_run(self._h, ctypes.c_void_p(stream)), "wd_program_run")

    def run_timed(self, stream=0):
        """Eager run with CUDA events between ops -> list of per-op milliseconds (measurement helper)."""
        n = int(self._lib.wd_program_num_ops(self._h))
        ms = (ctypes.c_float * n)()
        check(self._lib.wd_program_run_timed(self._h, ctypes.c_void_p(stream), ms), "wd_program_run_timed")
        return list(ms)

    def find_stuck_op(self, stream=0, timeout_ms=5000):
        """Debug: index of the first op that does not complete within timeout_ms (-1 if the program finishes)."""
        stuck = ctypes.c_int(-1)
        check(self._lib.wd_program_find_stuck_op(self._h, ctypes.c_void_p(stream), timeout_ms, ctypes.byref(stuck)), "wd_program_find_stuck_op")
        return stuck.value

    def capture(self, stream):
        check(self._lib.wd_program_capture(self._h, ctypes.c_void_p(stream)), "wd_program_capture")
        self._captured = True

    def replay(self, stream=0):
        if not self._captured:
            raise WdError("program not captured")
        check(self._lib.wd_program_replay(self._h, ctypes.c_void_p(stream)), "wd_program_replay")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.wd_program_destroy(self._h)
                self._h = None
        except Exception:
            pass
