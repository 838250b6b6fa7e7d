"""ctypes binding of libmauve_b200.so (the C ABI declared in include/mauve_b200.h).

The library is the product: there is no Python or CPU implementation of the path behind it.
Loading fails loudly if the shared object has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAUVE_B200_LIB") or os.path.join(_HERE, "libmauve_b200.so")  # the override is a tuning aid

MB_OK = 0
MODE_UNIQUE, MODE_SEED_ENUM, MODE_UNIQUE_COUNT, MODE_PAIRWISE, MODE_REPEAT = 0, 1, 2, 3, 4
SOLID_SEED = 2 ** 31 - 1
CODING_SEED = 3


class MbParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("direct_only", C.c_int32), ("min_multi", C.c_uint64), ("max_multi", C.c_uint64),
                ("nway_mask", C.c_uint64)]


class MbResult(C.Structure):
    _fields_ = [("n_matches", C.c_uint64), ("n_comps", C.c_uint64), ("length", C.POINTER(C.c_uint32)),
                ("comp_off", C.POINTER(C.c_uint64)), ("comp_seq", C.POINTER(C.c_uint32)), ("comp_start", C.POINTER(C.c_int64)),
                ("unique_mers", C.c_uint64), ("unique_mers_per_seq", C.POINTER(C.c_uint64)), ("nseq", C.c_uint32)]


class MbResultCompact(C.Structure):
    _fields_ = [("n_matches", C.c_uint64), ("n_comps", C.c_uint64), ("length", C.POINTER(C.c_uint32)),
                ("comp_off", C.POINTER(C.c_uint64)), ("comp_seq", C.POINTER(C.c_uint8)), ("comp_start", C.POINTER(C.c_int32)),
                ("unique_mers", C.c_uint64), ("unique_mers_per_seq", C.POINTER(C.c_uint64)), ("nseq", C.c_uint32)]


class MbBatchResult(C.Structure):
    _fields_ = [("n_problems", C.c_uint64), ("n_matches", C.c_uint64), ("n_comps", C.c_uint64), ("match_off", C.POINTER(C.c_uint64)),
                ("length", C.POINTER(C.c_uint32)), ("comp_off", C.POINTER(C.c_uint64)), ("comp_seq", C.POINTER(C.c_uint32)),
                ("comp_start", C.POINTER(C.c_int64))]


class MbStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_seeds", "n_runs", "n_buckets", "n_candidates", "n_extended", "n_matches", "n_comps")] + \
               [(n, C.c_uint32) for n in ("radix_passes", "record_bytes", "dedup_batches", "dedup_iters")] + \
               [(n, C.c_float) for n in ("ms_h2d", "ms_pack", "ms_extract", "ms_sort", "ms_bucket", "ms_dedup", "ms_output", "ms_d2h",
                                         "ms_total_device")] + \
               [(n, C.c_uint64) for n in ("h2d_bytes", "d2h_bytes", "kernel_launches")] + \
               [("ms_radix_kernels", C.c_float), ("radix_launches", C.c_uint32)]


EXPORTS = ["mb_ctx_create", "mb_ctx_destroy", "mb_set_stream", "mb_add_sequence", "mb_add_sequence_device", "mb_clear_sequences", "mb_accumulate",
           "mb_set_seed", "mb_find", "mb_find_device", "mb_fetch_result", "mb_get_sml", "mb_get_mers", "mb_get_stats", "mb_strerror",
           "mb_last_cuda_error", "mb_device_count", "mb_version",
           "mb_dist_extract", "mb_dist_extract_records", "mb_dist_enum_local", "mb_dist_extract_count", "mb_dist_partition", "mb_dist_p2p_recv_array",
           "mb_dist_use_p2p_recv", "mb_ipc_export", "mb_ipc_import", "mb_ipc_close", "mb_dist_recv_buffer", "mb_dist_local", "mb_dist_rows_pack", "mb_dist_push", "mb_dist_resolve", "mb_dist_accept", "mb_dist_match_pack", "mb_dist_match_partition", "mb_dist_output", "mb_dist_stage_ms", "mb_find_multi", "mb_debug_radix", "mb_find_batch", "mb_set_segments", "mb_position_table", "mb_fetch_result_compact", "mb_find_compact", "mb_get_packed_device", "mb_add_sequence_device_packed", "mb_copy_packed_device"]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C mauvealigner_b200/csrc` "
                          "(or __graft_entry__.build()); there is no fallback implementation")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.mb_ctx_create.argtypes = [C.POINTER(vp), i32]
    L.mb_ctx_destroy.argtypes = [vp]
    L.mb_set_stream.argtypes = [vp, vp]
    L.mb_add_sequence.argtypes = [vp, vp, u64, i32, C.POINTER(i32)]
    L.mb_add_sequence_device.argtypes = [vp, vp, u64, C.POINTER(i32)]
    L.mb_add_sequence_device_packed.argtypes = [vp, vp, u64, C.POINTER(i32)]
    L.mb_get_packed_device.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(u64)]
    L.mb_copy_packed_device.argtypes = [vp, i32, vp]
    L.mb_clear_sequences.argtypes = [vp]
    L.mb_accumulate.argtypes = [vp, i32]
    L.mb_set_seed.argtypes = [vp, u64]
    L.mb_find.argtypes = [vp, C.POINTER(MbParams), C.POINTER(C.POINTER(MbResult))]
    L.mb_find_device.argtypes = [vp, C.POINTER(MbParams)]
    L.mb_fetch_result.argtypes = [vp, C.POINTER(C.POINTER(MbResult))]
    L.mb_fetch_result_compact.argtypes = [vp, C.POINTER(C.POINTER(MbResultCompact))]
    L.mb_find_compact.argtypes = [vp, C.POINTER(MbParams), C.POINTER(C.POINTER(MbResultCompact))]
    L.mb_get_sml.argtypes = [vp, i32, vp, u64, C.POINTER(u64)]
    L.mb_get_mers.argtypes = [vp, i32, vp, u64, C.POINTER(u64)]
    L.mb_get_stats.argtypes = [vp, C.POINTER(MbStats)]
    L.mb_strerror.argtypes = [i32]
    L.mb_strerror.restype = C.c_char_p
    L.mb_last_cuda_error.argtypes = [vp]
    L.mb_last_cuda_error.restype = C.c_char_p
    L.mb_version.restype = C.c_char_p
    pu64 = C.POINTER(u64)
    L.mb_dist_extract.argtypes = [vp, i32, i32, C.POINTER(vp), pu64]
    L.mb_dist_extract_records.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(vp), pu64]
    L.mb_dist_enum_local.argtypes = [vp, C.POINTER(MbParams), u64]
    L.mb_dist_extract_count.argtypes = [vp, i32, i32, pu64]
    L.mb_dist_partition.argtypes = [vp, C.POINTER(vp), pu64, C.POINTER(vp)]
    L.mb_dist_p2p_recv_array.argtypes = [vp, u64, C.POINTER(vp)]
    L.mb_dist_use_p2p_recv.argtypes = [vp, i32]
    L.mb_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.mb_ipc_import.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.mb_ipc_close.argtypes = [vp, vp]
    L.mb_dist_recv_buffer.argtypes = [vp, i32, u64, C.POINTER(vp)]
    L.mb_dist_local.argtypes = [vp, C.POINTER(MbParams), u64, pu64]
    L.mb_dist_push.argtypes = [vp, vp, pu64, C.c_uint32, C.POINTER(vp), pu64]
    L.mb_dist_rows_pack.argtypes = [vp, C.POINTER(vp), pu64, C.POINTER(vp)]
    L.mb_dist_match_pack.argtypes = [vp, C.POINTER(vp), pu64, C.POINTER(vp), pu64, C.POINTER(vp), C.POINTER(vp)]
    L.mb_dist_resolve.argtypes = [vp, u64, C.POINTER(vp)]
    L.mb_dist_accept.argtypes = [vp, C.POINTER(vp)]
    L.mb_dist_match_partition.argtypes = [vp, pu64, pu64]
    L.mb_dist_output.argtypes = [vp, u64, u64]
    L.mb_find_multi.argtypes = [C.POINTER(vp), i32, C.POINTER(MbParams)]
    L.mb_dist_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.mb_debug_radix.argtypes = [vp, u64, i32, i32, i32, C.POINTER(C.c_float)]
    L.mb_find_batch.argtypes = [vp, C.POINTER(MbParams), C.c_uint32, C.c_uint32, C.POINTER(vp), pu64, C.POINTER(C.POINTER(MbBatchResult))]
    L.mb_set_segments.argtypes = [vp, C.c_uint32, pu64]
    L.mb_position_table.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint32)), pu64]
    _lib = L
    return L


class MauveError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = lib().mb_strerror(code).decode()
        super().__init__(f"mauve_b200 error {code}: {msg}{(' — ' + detail) if detail else ''}")
