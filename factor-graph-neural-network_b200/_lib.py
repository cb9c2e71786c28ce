"""ctypes binding of libfgnn_b200.so (C ABI in include/fgnn_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FGNN_B200_LIB") or os.path.join(HERE, "libfgnn_b200.so")   # override: debug/trace builds

# enums (include/fgnn_b200.h)
NO_EXTENSION, ORIG_WITH_NEIGHBOR, ORIG_WITH_DIFF = 0, 1, 2
AGG_MAX, AGG_SOFTMAX, AGG_MEAN, AGG_NONE = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU = 0, 1, 2
F32, BF16 = 0, 1
I64, I32 = 0, 1
KERNEL_AUTO, KERNEL_SIMT, KERNEL_TCGEN05 = 0, 1, 2
FLAG_MASK_NEGATIVE = 1
FLAG_ACCUMULATE = 2
FLAG_NO_FUSED_REDUCE = 4

OK = 0
ERR_INVALID_ARG, ERR_INDEX_RANGE, ERR_SHAPE, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA, ERR_NO_DEVICE = (
    -1, -2, -3, -4, -5, -6, -7)


class MpArgs(ctypes.Structure):
    """struct fgnn_mp_args"""
    _fields_ = [
        ("x", ctypes.c_void_p), ("idx", ctypes.c_void_p), ("etype", ctypes.c_void_p),
        ("filters", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("bn_scale", ctypes.c_void_p),
        ("bn_shift", ctypes.c_void_p), ("out", ctypes.c_void_p), ("workspace", ctypes.c_void_p),
        ("workspace_bytes", ctypes.c_size_t),
        ("x_sb", ctypes.c_int64), ("x_sc", ctypes.c_int64), ("x_sn", ctypes.c_int64),
        ("idx_sb", ctypes.c_int64), ("et_sb", ctypes.c_int64),
        ("out_sb", ctypes.c_int64), ("out_so", ctypes.c_int64), ("out_sm", ctypes.c_int64),
        ("out_sk", ctypes.c_int64),
        ("B", ctypes.c_int32), ("N", ctypes.c_int32), ("M", ctypes.c_int32), ("K", ctypes.c_int32),
        ("C", ctypes.c_int32), ("O", ctypes.c_int32), ("T", ctypes.c_int32),
        ("extension", ctypes.c_int32), ("aggregator", ctypes.c_int32), ("activation", ctypes.c_int32),
        ("dtype", ctypes.c_int32), ("idx_dtype", ctypes.c_int32), ("kernel", ctypes.c_int32),
        ("flags", ctypes.c_uint32), ("gamma", ctypes.c_float), ("act_slope", ctypes.c_float),
        ("filters_version", ctypes.c_int64),
        ("tile_slots", ctypes.c_void_p), ("out_rows", ctypes.c_void_p),
        ("sm_limit", ctypes.c_int32), ("reserved_", ctypes.c_int32),
        ("src_ptr", ctypes.c_void_p), ("slot_edge", ctypes.c_void_p), ("etype_edges", ctypes.c_void_p),
        ("messages", ctypes.c_void_p), ("n_edges", ctypes.c_int64),
        ("src_rows", ctypes.c_void_p), ("n_src_rows", ctypes.c_int64), ("src_row_cap", ctypes.c_int32),
        ("src_rows_per_batch", ctypes.c_int32), ("src_edge_slot", ctypes.c_void_p),
    ]


class ExchangeArgs(ctypes.Structure):
    """struct fgnn_exchange_args"""
    _fields_ = [
        ("raw", ctypes.c_void_p * 8), ("out", ctypes.c_void_p * 8), ("flags", ctypes.c_void_p * 8),
        ("counter", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("bn_scale", ctypes.c_void_p),
        ("bn_shift", ctypes.c_void_p),
        ("rows", ctypes.c_int64), ("row0", ctypes.c_int64), ("row1", ctypes.c_int64),
        ("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("J", ctypes.c_int32), ("O", ctypes.c_int32),
        ("activation", ctypes.c_int32), ("act_slope", ctypes.c_float), ("epoch", ctypes.c_uint32),
        ("ctas", ctypes.c_int32), ("raw_mask", ctypes.c_void_p), ("out_mask", ctypes.c_void_p),
    ]


class HaloJob(ctypes.Structure):
    """struct fgnn_halo_job"""
    _fields_ = [("src", ctypes.c_void_p * 8), ("dst", ctypes.c_void_p), ("src_row", ctypes.c_void_p), ("src_rank", ctypes.c_void_p),
                ("dst_row0", ctypes.c_int64), ("n", ctypes.c_int32), ("row_bytes", ctypes.c_int32)]


class HaloArgs(ctypes.Structure):
    """struct fgnn_halo_args"""
    _fields_ = [("jobs", HaloJob * 4), ("flags", ctypes.c_void_p * 8), ("counter", ctypes.c_void_p),
                ("n_jobs", ctypes.c_int32), ("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("epoch", ctypes.c_uint32),
                ("ctas", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


EXPORTS = {
    "fgnn_version": (ctypes.c_int, []),
    "fgnn_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "fgnn_last_cuda_error": (ctypes.c_int, []),
    "fgnn_mp_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(MpArgs)]),
    "fgnn_mp_select_kernel": (ctypes.c_int, [ctypes.POINTER(MpArgs)]),
    "fgnn_mp_forward": (ctypes.c_int, [ctypes.POINTER(MpArgs), ctypes.c_void_p]),
    "fgnn_mp_forward_host": (ctypes.c_int, [ctypes.POINTER(MpArgs)]),
    "fgnn_check_index_range": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "fgnn_check_index_range_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                    ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "fgnn_epilogue_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_int32, ctypes.c_float,
                                             ctypes.c_void_p]),
    "fgnn_epilogue_sum_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                                 ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_int32, ctypes.c_float, ctypes.c_int32, ctypes.c_void_p]),
    "fgnn_instance_norm_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                                   ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                   ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_float,
                                                   ctypes.c_int32, ctypes.c_float, ctypes.c_void_p]),
    "fgnn_instance_norm_partial": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                                   ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    "fgnn_instance_norm_apply": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                                 ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                                 ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_float,
                                                 ctypes.c_void_p]),
    "fgnn_to_node_major": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                          ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    "fgnn_mp_src_supported": (ctypes.c_int, [ctypes.POINTER(MpArgs)]),
    "fgnn_src_permute_etype": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                              ctypes.c_void_p]),
    "fgnn_emodel_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_void_p]),
    "fgnn_bwd_gather": (ctypes.c_int, [ctypes.POINTER(MpArgs), ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "fgnn_bwd_slot_values": (ctypes.c_int, [ctypes.POINTER(MpArgs), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_void_p]),
    "fgnn_bwd_aggregate": (ctypes.c_int, [ctypes.POINTER(MpArgs), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_void_p]),
    "fgnn_bwd_outer": (ctypes.c_int, [ctypes.POINTER(MpArgs), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "fgnn_bwd_scatter": (ctypes.c_int, [ctypes.POINTER(MpArgs), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_void_p]),
    "fgnn_plan_build_host": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64),
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "fgnn_locality_order_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "fgnn_comm_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "fgnn_comm_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "fgnn_comm_close": (ctypes.c_int, [ctypes.c_void_p]),
    "fgnn_comm_free": (ctypes.c_int, [ctypes.c_void_p]),
    "fgnn_halo_pull": (ctypes.c_int, [ctypes.POINTER(HaloArgs), ctypes.c_void_p]),
    "fgnn_exchange_forward": (ctypes.c_int, [ctypes.POINTER(ExchangeArgs), ctypes.c_void_p]),
    "fgnn_launch_count": (ctypes.c_uint64, []),
    "fgnn_set_programmatic_launch": (ctypes.c_int, [ctypes.c_int]),
}

_lib = None


class FgnnError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib().fgnn_strerror(status).decode()
        if status == ERR_CUDA:
            msg += f" [cudaError {lib().fgnn_last_cuda_error()}]"
        super().__init__(f"fgnn_b200 {where}: {msg} (status {status})")


def lib():
    """Load libfgnn_b200.so; raises if it has not been built (python -m ... build, or
    __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"fgnn_b200: CUDA library not built: {LIB_PATH} is missing. Run "
                f"`python factor-graph-neural-network_b200/build.py` (needs nvcc). There is no CPU or "
                f"PyTorch fallback for this path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)     # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, where):
    if status == OK:
        return
    if status == ERR_INDEX_RANGE:
        raise IndexError(f"fgnn_b200 {where}: nn_idx entry out of range")
    if status == ERR_INVALID_ARG and where.endswith("extension"):
        raise ValueError("extension must one of mp_conv_type")
    raise FgnnError(status, where)
