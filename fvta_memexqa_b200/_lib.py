"""ctypes binding of libfvta_b200.so (include/fvta_b200.h).  No torch types cross this
boundary: only raw device pointers, sizes and the stream handle.

There is no CPU fallback: if the library is missing or fails to load this module raises.
"""
import ctypes
import os

from .build import LIB_PATH

c_float_p = ctypes.c_void_p
c_u8_p = ctypes.c_void_p
c_i32_p = ctypes.c_void_p


class FvtaAttnDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "N", "K", "T", "D", "JQ", "simi", "add_tanh", "use_time_warp", "warp_type", "window",
        "time_warp_att", "has_mask", "ws_from_fwd", "path")]


class FvtaAttnParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("att_W", "att_b", "WH_W", "WH_b", "WC_W", "WC_b")]


class FvtaAttnGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("att_W", "att_b", "WH_W", "WH_b", "WC_W", "WC_b")]


class FvtaLstmDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "J", "In", "d", "path")]


class FvtaReaderDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "ntok", "In", "d", "Jmax", "path")]


LSTM_AUTO, LSTM_FFMA, LSTM_TC = 0, 1, 2
PATH_AUTO, PATH_SMALL, PATH_FFMA, PATH_TC = 0, 1, 2, 3
PATH_NAMES = {1: "attn_small", 2: "attn_tile_fwd (FFMA)", 3: "attn_fwd_tc (tcgen05)"}

ERRORS = {0: "FVTA_OK", -1: "FVTA_E_BADARG", -2: "FVTA_E_UNSUPPORTED", -3: "FVTA_E_WORKSPACE",
          -4: "FVTA_E_LAUNCH", -5: "FVTA_E_SHAPE"}

# name -> (restype, argtypes); every symbol include/fvta_b200.h declares
SYMBOLS = {
    "fvta_version": (ctypes.c_int, []),
    "fvta_last_error": (ctypes.c_char_p, []),
    "fvta_launch_count": (ctypes.c_ulonglong, []),
    "fvta_attn_fwd_path": (ctypes.c_int, [ctypes.POINTER(FvtaAttnDesc), ctypes.c_int, ctypes.c_int]),
    "fvta_attn_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(FvtaAttnDesc)]),
    "fvta_attn_stats_floats": (ctypes.c_size_t, [ctypes.POINTER(FvtaAttnDesc)]),
    "fvta_attn_fwd": (ctypes.c_int, [ctypes.POINTER(FvtaAttnDesc), c_float_p, c_u8_p, c_float_p, c_u8_p,
                                     c_float_p, ctypes.POINTER(FvtaAttnParams), c_float_p, c_float_p,
                                     c_float_p, c_float_p, c_float_p, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_void_p]),
    "fvta_attn_bwd": (ctypes.c_int, [ctypes.POINTER(FvtaAttnDesc), c_float_p, c_u8_p, c_float_p, c_u8_p,
                                     c_float_p, ctypes.POINTER(FvtaAttnParams), c_float_p, c_float_p,
                                     c_float_p, c_float_p, c_float_p, ctypes.POINTER(FvtaAttnGrads),
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fvta_bilstm_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(FvtaLstmDesc)]),
    "fvta_bilstm_reserve_floats": (ctypes.c_size_t, [ctypes.POINTER(FvtaLstmDesc)]),
    "fvta_bilstm_fwd": (ctypes.c_int, [ctypes.POINTER(FvtaLstmDesc), c_float_p, c_i32_p, c_float_p,
                                       c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fvta_bilstm_bwd": (ctypes.c_int, [ctypes.POINTER(FvtaLstmDesc), c_float_p, c_i32_p, c_float_p,
                                       c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                       c_float_p, c_float_p, c_float_p, c_float_p,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fvta_reader_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(FvtaReaderDesc)]),
    "fvta_reader_reserve_floats": (ctypes.c_size_t, [ctypes.POINTER(FvtaReaderDesc)]),
    "fvta_reader_fwd": (ctypes.c_int, [ctypes.POINTER(FvtaReaderDesc), c_float_p, c_i32_p, c_i32_p, ctypes.c_void_p,
                                       c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fvta_reader_bwd": (ctypes.c_int, [ctypes.POINTER(FvtaReaderDesc), c_float_p, c_i32_p, c_i32_p, ctypes.c_void_p,
                                       c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                       c_float_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fvta_gemm": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_float_p, ctypes.c_int32,
                                 ctypes.c_int32, c_float_p, ctypes.c_int32, ctypes.c_int32, c_float_p,
                                 ctypes.c_int32, c_float_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "fvta_gemm_gather": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_float_p, ctypes.c_int32,
                                        ctypes.c_void_p, ctypes.c_int32, c_float_p, ctypes.c_int32, ctypes.c_int32,
                                        c_float_p, ctypes.c_int32, c_float_p, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_void_p]),
    "fvta_charcnn_fwd": (ctypes.c_int, [ctypes.c_int32] * 5 + [c_i32_p, c_float_p, c_float_p, c_float_p,
                                                                 ctypes.c_int32, c_u8_p, ctypes.c_void_p]),
    "fvta_charcnn_bwd": (ctypes.c_int, [ctypes.c_int32] * 5 + [c_i32_p, c_float_p, ctypes.c_int32, c_u8_p, c_float_p,
                                                                 ctypes.c_int32, c_float_p, c_float_p,
                                                                 ctypes.c_void_p]),
    "fvta_embed_rows": (ctypes.c_int, [ctypes.c_int32] * 4 + [ctypes.c_void_p, c_float_p, c_float_p, c_float_p,
                                                                ctypes.c_int32, ctypes.c_void_p]),
    "fvta_embed_rows_bwd": (ctypes.c_int, [ctypes.c_int32] * 3 + [ctypes.c_void_p, c_float_p, ctypes.c_int32,
                                                                    c_float_p, ctypes.c_void_p]),
    "fvta_bias_act_grad": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, c_float_p, ctypes.c_int32, c_float_p,
                                          ctypes.c_int32, ctypes.c_int32, c_float_p, ctypes.c_int32, c_float_p,
                                          ctypes.c_void_p]),
    "fvta_output_loss": (ctypes.c_int, [ctypes.c_int32] * 3 + [c_float_p] * 5 + [c_u8_p] + [c_float_p] * 8 +
                         [ctypes.c_void_p]),
    "fvta_adadelta": (ctypes.c_int, [ctypes.c_int64, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                     c_i32_p, c_float_p, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_void_p]),
    "fvta_p2p_allreduce_adadelta": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                                   c_float_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   c_float_p, c_float_p, c_float_p, c_float_p, c_i32_p,
                                                   ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
}

_lib = None


class FvtaError(RuntimeError):
    pass


def load():
    """Load the library once; raise loudly if it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FvtaError(
            "libfvta_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "-- the FVTA hot path has no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.fvta_last_error().decode() if rc == -4 else ""
        raise FvtaError("%s failed: %s %s" % (what, ERRORS.get(rc, rc), msg))
