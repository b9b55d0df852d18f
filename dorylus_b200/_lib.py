"""ctypes binding of libdorylus_b200.so (the C ABI in include/dorylus_b200.h).

The library is loaded from the package directory (built in-tree by dorylus_b200.build).  If it is
missing this raises immediately: there is no Python / CPU fallback for any operator.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdorylus_b200.so")

DORY_ABI_VERSION = 1
DORY_MAX_LAYERS = 8
DORY_UNIQUE_ID_BYTES = 128
DORY_IPC_BLOB_BYTES = 80

OK, EINVAL, ESTATE, ECUDA, ENOMEM, ECOMM, EFORMAT, ENODEV = 0, -1, -2, -3, -4, -5, -6, -7
ERROR_NAMES = {EINVAL: "DORY_EINVAL", ESTATE: "DORY_ESTATE", ECUDA: "DORY_ECUDA", ENOMEM: "DORY_ENOMEM",
               ECOMM: "DORY_ECOMM", EFORMAT: "DORY_EFORMAT", ENODEV: "DORY_ENODEV"}

FORWARD, BACKWARD = 0, 1
GCN, GAT = 0, 1
FLAG_STRICT_MASK, FLAG_GAT_PREDICT_AH, FLAG_NO_TENSOR_CORES, FLAG_APPLY_FIRST = 0x1, 0x2, 0x4, 0x8


class DoryChunk(C.Structure):
    """== struct Chunk (reference src/common/utils.hpp:64-74)."""

    _fields_ = [("localId", C.c_uint32), ("globalId", C.c_uint32), ("lowBound", C.c_uint32),
                ("upBound", C.c_uint32), ("layer", C.c_uint32), ("dir", C.c_uint32), ("epoch", C.c_uint32),
                ("vertex", C.c_uint8)]


class DoryConfig(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("gnn_type", C.c_uint32), ("n_layers", C.c_uint32),
                ("dims", C.c_uint32 * (DORY_MAX_LAYERS + 1)), ("node_id", C.c_uint32),
                ("num_nodes", C.c_uint32), ("device", C.c_int32), ("learning_rate", C.c_float),
                ("flags", C.c_uint32)]


class DoryStats(C.Structure):
    _fields_ = [("acc_sum", C.c_float), ("loss_sum", C.c_float), ("val_rows", C.c_uint32),
                ("epochs_done", C.c_uint32), ("kernel_launches", C.c_uint64), ("edges_aggregated", C.c_uint64)]


# every symbol include/dorylus_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_u32, _u64, _f32p = C.c_uint32, C.c_uint64, C.POINTER(C.c_float)
_u32p, _u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
_chunkp = C.POINTER(DoryChunk)
SYMBOLS = {
    "dory_create": (C.c_int, [C.POINTER(_P), C.POINTER(DoryConfig)]),
    "dory_destroy": (None, [_P]),
    "dory_last_error": (C.c_char_p, [_P]),
    "dory_abi_version": (C.c_int, []),
    "dory_sync": (C.c_int, [_P]),
    "dory_set_option": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "dory_preprocess_edges": (C.c_int, [_u32p, _u32p, _u64, C.POINTER(C.c_int32), _u32, _u32, _u32, C.c_int,
                                        C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "dory_preprocess_dir": (C.c_int, [C.c_char_p, _u32, _u32, C.c_int]),
    "dory_preprocess_incident_edges": (C.c_int, [_u32p, _u32p, _u64, C.POINTER(C.c_int32), _u32, _u32, _u32, _u32p, _u64,
                                                 C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "dory_free": (None, [_P]),
    "dory_read_features": (C.c_int, [C.c_char_p, C.c_char_p, _P, C.c_size_t, _u32, _u32, _f32p, _f32p]),
    "dory_read_labels": (C.c_int, [C.c_char_p, _P, C.c_size_t, _u32, _f32p]),
    "dory_partition_edges": (C.c_int, [_u32p, _u32p, _u64, _u32, _u32, _u32, C.POINTER(C.c_int32), _u64p]),
    "dory_partition_file": (C.c_int, [C.c_char_p, _u32, C.c_char_p]),
    "dory_load_partition": (C.c_int, [_P, _P, C.c_size_t]),
    "dory_tile_info": (C.c_int, [_P, _u32, C.POINTER(C.c_double), _u32p, _u32p, _u32p]),
    "dory_graph_counts": (C.c_int, [_P, _u64p]),
    "dory_set_tensor": (C.c_int, [_P, _u32, C.c_char_p, _f32p, _u64, _u32]),
    "dory_get_tensor": (C.c_int, [_P, _u32, C.c_char_p, _f32p, _u64, _u32]),
    "dory_tensor_shape": (C.c_int, [_P, _u32, C.c_char_p, _u64p, _u32p]),
    "dory_prefetch_tensor": (C.c_int, [_P, _u32, C.c_char_p, _f32p, _u64, _u32]),
    "dory_commit_prefetch": (C.c_int, [_P]),
    "dory_tensor_device": (C.c_int, [_P, _u32, C.c_char_p, C.POINTER(_P), _u64p, _u32p, _u32p]),
    "dory_init_weights": (C.c_int, [_P]),
    "dory_set_weights": (C.c_int, [_P, _u32, C.c_char_p, _f32p, _u32, _u32]),
    "dory_get_weights": (C.c_int, [_P, _u32, C.c_char_p, _f32p, _u32, _u32]),
    "dory_get_weight_grad": (C.c_int, [_P, _u32, C.c_char_p, _f32p, _u32, _u32]),
    "dory_apply_update": (C.c_int, [_P, _u32]),
    "dory_aggregate": (C.c_int, [_P, _chunkp]),
    "dory_apply_vertex": (C.c_int, [_P, _chunkp]),
    "dory_scatter": (C.c_int, [_P, _chunkp]),
    "dory_apply_edge": (C.c_int, [_P, _chunkp]),
    "dory_predict": (C.c_int, [_P, _chunkp]),
    "dory_inc_layer": (C.c_int, [_P, _chunkp]),
    "dory_layer_schedule": (C.c_int, [_P, _u32, C.POINTER(C.c_int)]),
    "dory_forward": (C.c_int, [_P, _u32]),
    "dory_backward": (C.c_int, [_P, _u32]),
    "dory_epoch": (C.c_int, [_P, C.POINTER(DoryStats)]),
    "dory_get_stats": (C.c_int, [_P, C.POINTER(DoryStats)]),
    "dory_stats_enqueue": (C.c_int, [_P, _u32]),
    "dory_stats_collect": (C.c_int, [_P, _u32, C.POINTER(DoryStats)]),
    "dory_comm_unique_id": (C.c_int, [_P]),
    "dory_comm_init": (C.c_int, [_P, _P]),
    "dory_comm_set_recv_slots": (C.c_int, [_P, _u32, _u32, _u32p, _u32]),
    "dory_comm_send_gvids": (C.c_int, [_P, _u32, _u32, _u32p, _u32p]),
    "dory_ghost_slots": (C.c_int, [_P, C.c_size_t, _u32, _P, C.c_size_t, _u32, _u32p, _u32p]),
    "dory_comm_set_send_slots": (C.c_int, [_P, _u32, _u32, _u32p, _u32]),
    "dory_comm_ipc_export": (C.c_int, [_P, _u32, C.c_char_p, _P]),
    "dory_comm_ipc_import": (C.c_int, [_P, _u32, C.c_char_p, _u32, _P]),
    "dory_event_record": (C.c_int, [_P, _u32]),
    "dory_event_elapsed_ms": (C.c_int, [_P, _u32, _u32, _f32p]),
    "dory_flush_l2": (C.c_int, [_P, C.c_size_t]),
    "dory_measure_fma_peak": (C.c_int, [_P, _f32p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the native library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -m dorylus_b200.build` (nvcc, sm_100a). "
                "dorylus_b200 has no CPU / pure-Python fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.dory_abi_version() != DORY_ABI_VERSION:
            raise ImportError("libdorylus_b200.so ABI %d != binding %d" % (lib.dory_abi_version(), DORY_ABI_VERSION))
        _lib = lib
    return _lib


class DoryError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("%s (%d): %s" % (ERROR_NAMES.get(code, "DORY_E?"), code, message))
        self.code = code
