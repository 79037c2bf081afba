"""ctypes binding of libkmerjs_b200.so (include/kmerjs_b200.h).

This is the same C ABI the Node.js N-API addon binds (node/addon.cc); Python is
the host language here only because the build image has no Node.js.  There is no CPU fallback:
if the library cannot be loaded, or the machine has no sm_100 device, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

KJ_OK = 0
KJ_E_INVALID, KJ_E_NO_SM100, KJ_E_CUDA, KJ_E_NOMEM, KJ_E_IO = -1, -2, -3, -4, -5
KJ_E_TABLE_FULL, KJ_E_NO_HITS, KJ_E_NO_WINNER, KJ_E_RANGE, KJ_E_STATE = -6, -7, -8, -9, -10
KJ_MEM_HOST, KJ_MEM_DEVICE = 0, 1
KJ_F_NO_ORDER, KJ_F_FORCE_GENERIC, KJ_F_FORWARD_ONLY, KJ_F_NO_LINE_GATE, KJ_F_COUNT_BASES = 1, 2, 4, 8, 16
KJ_VEC_SCORES, KJ_VEC_FIRST_ORD, KJ_VEC_FIRST_IDX = 0, 1, 2
KJ_DB_AUTO, KJ_DB_KMER_DOCS, KJ_DB_TEMPLATE_DOCS, KJ_DB_KMERFINDER_MAP, KJ_DB_PACKED = 0, 1, 2, 3, 4
KJ_ABI_VERSION = 1

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class KjError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{message} (kj error {code})")
        self.code = code
        self.message = message


class kj_count_params(C.Structure):
    _fields_ = [("prefix", C.c_char_p), ("prefix_len", C.c_uint32), ("k", C.c_uint32),
                ("step", C.c_uint32), ("flags", C.c_uint32), ("base_line", C.c_uint64),
                ("base_col", C.c_uint64), ("capacity_hint", C.c_uint64)]


class kj_db_desc(C.Structure):
    _fields_ = [("n_kmers", C.c_uint64), ("kmer_bytes", C.c_void_p), ("kmer_len", C.c_void_p),
                ("list_off", C.c_void_p), ("tmpl_ids", C.c_void_p), ("n_templates", C.c_uint32),
                ("lengths", C.c_void_p), ("ulengths", C.c_void_p),
                ("summary_templates", C.c_uint64), ("summary_unique_lens", C.c_uint64),
                ("summary_total_len", C.c_uint64), ("part", C.c_uint32), ("n_parts", C.c_uint32)]


class kj_row(C.Structure):
    _fields_ = [("template_id", C.c_uint32), ("reserved", C.c_uint32), ("score", C.c_uint64),
                ("expected", C.c_double), ("z", C.c_double), ("probability", C.c_double),
                ("frac_q", C.c_double), ("frac_d", C.c_double), ("depth", C.c_double),
                ("kmers_template", C.c_uint64), ("total_frac_q", C.c_double),
                ("total_frac_d", C.c_double), ("total_temp_cover", C.c_double),
                ("tscore", C.c_uint64), ("hits", C.c_uint64), ("z_device", C.c_double),
                ("probability_device", C.c_double)]


class kj_synth_params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_uint64), ("read_len", C.c_uint32),
                ("first_read", C.c_uint64), ("genome", C.c_void_p), ("genome_len", C.c_uint64),
                ("sub_rate", C.c_double), ("n_rate", C.c_double), ("lead_n_rate", C.c_double)]


vp = C.c_void_p
# name -> (restype, argtypes): every symbol include/kmerjs_b200.h declares
SIGNATURES = {
    "kj_init": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "kj_destroy": (None, [vp]),
    "kj_last_error": (C.c_char_p, [vp]),
    "kj_abi_version": (C.c_int, []),
    "kj_launch_count": (C.c_uint64, [vp]),
    "kj_scan_kernel_ms": (C.c_double, [vp, u64p]),
    "kj_verify_kernel_ms": (C.c_double, [vp]),
    "kj_scan_kernel_bytes": (C.c_uint64, [vp]),
    "kj_reset_timers": (None, [vp]),
    "kj_set_stage_chunk": (C.c_int, [vp, C.c_uint64]),
    "kj_enable_timers": (None, [vp, C.c_int]),
    "kj_counts_create": (C.c_int, [vp, C.POINTER(kj_count_params), C.POINTER(vp)]),
    "kj_counts_add_buffer": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int]),
    "kj_counts_add_file": (C.c_int, [vp, C.c_char_p]),
    "kj_counts_finish": (C.c_int, [vp]),
    "kj_counts_size": (C.c_uint64, [vp]),
    "kj_counts_lines": (C.c_uint64, [vp]),
    "kj_counts_bases": (C.c_uint64, [vp]),
    "kj_counts_bytes_read": (C.c_uint64, [vp]),
    "kj_counts_occurrences": (C.c_uint64, [vp]),
    "kj_counts_export": (C.c_int, [vp, vp, vp, vp]),
    "kj_counts_alive": (C.c_int, [vp, vp]),
    "kj_counts_free": (None, [vp]),
    "kj_count_newlines": (C.c_int, [vp, vp, C.c_uint64, C.c_int, u64p, u64p]),
    "kj_counts_partition": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), u64p]),
    "kj_counts_merge_records": (C.c_int, [vp, vp, C.c_uint64]),
    "kj_counts_merge_host_records": (C.c_int, [vp, vp, C.c_uint64]),
    "kj_counts_irregular_size": (C.c_uint64, [vp]),
    "kj_counts_irregular_export": (C.c_int, [vp, vp]),
    "kj_counts_irregular_merge": (C.c_int, [vp, vp, C.c_uint64]),
    "kj_counts_irregular_merge_part": (C.c_int, [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32]),
    "kj_segment_bytes": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "kj_counts_partition_segments": (C.c_int, [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]),
    "kj_counts_merge_segments": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "kj_counts_set_totals": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "kj_owner": (C.c_uint32, [C.c_char_p, C.c_uint32, C.c_uint32]),
    "kj_db_create": (C.c_int, [vp, C.POINTER(kj_db_desc), C.POINTER(vp)]),
    "kj_db_load": (C.c_int, [vp, C.c_char_p, C.c_int, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "kj_db_save_packed": (C.c_int, [C.c_char_p, C.POINTER(kj_db_desc), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]),
    "kj_db_template": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), u64p, u64p]),
    "kj_db_summary": (C.c_int, [vp, u64p, u64p, u64p]),
    "kj_db_free": (None, [vp]),
    "kj_db_n_kmers": (C.c_uint64, [vp]),
    "kj_db_n_pairs": (C.c_uint64, [vp]),
    "kj_db_n_templates": (C.c_uint32, [vp]),
    "kj_first_match": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
    "kj_first_match_local": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
    "kj_match_vec_len": (C.c_uint64, [vp, C.c_int]),
    "kj_match_get": (C.c_int, [vp, C.c_int, vp]),
    "kj_match_set": (C.c_int, [vp, C.c_int, vp]),
    "kj_match_get_async": (C.c_int, [vp, C.c_int, vp]),
    "kj_match_set_async": (C.c_int, [vp, C.c_int, vp]),
    "kj_match_commit": (C.c_int, [vp]),
    "kj_match_set_query_size": (C.c_int, [vp, C.c_uint64]),
    "kj_match_hits": (C.c_uint64, [vp]),
    "kj_match_n_matched": (C.c_uint32, [vp]),
    "kj_match_scores": (C.c_int, [vp, vp, vp, vp]),
    "kj_match_template_kmers": (C.c_int, [vp, C.c_uint32, vp, C.c_uint64, u64p]),
    "kj_match_free": (None, [vp]),
    "kj_wta_next": (C.c_int, [vp, C.POINTER(kj_row)]),
    "kj_match_set_max_hits": (C.c_int, [vp, C.c_uint32]),
    "kj_match_defer_rows": (C.c_int, [vp, C.c_int]),
    "kj_wta_row": (C.c_int, [vp, C.POINTER(kj_row)]),
    "kj_wta_all": (C.c_int, [vp, C.POINTER(kj_row), C.c_uint32, u32p, C.POINTER(C.c_int)]),
    "kj_match_matched_size": (C.c_int, [vp, u64p, u64p]),
    "kj_match_export_matched": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64]),
    "kj_match_from_matched": (C.c_int, [vp, vp, C.c_uint32, u64p, u64p, u64p, u64p, C.c_uint64, C.POINTER(vp)]),
    "kj_matched_segment_bytes": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "kj_match_export_segment": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64]),
    "kj_counts_export_matched_segment": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32]),
    "kj_match_from_segments": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "kj_match_query_size": (C.c_uint64, [vp]),
    "kj_match_segment_sizes": (C.c_int, [vp, u64p, u64p]),
    "kj_standard_scoring": (C.c_int, [vp, C.POINTER(kj_row), C.c_uint32, u32p]),
    "kj_set_rounding_mode": (C.c_int, [vp, C.c_int]),
    "kj_stats_zscore": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, f64p,
                                  C.c_char_p, C.c_uint64]),
    "kj_stats_fastp_text": (C.c_int, [C.c_char_p, f64p]),
    "kj_stats_zscore_device": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp, vp]),
    "kj_stats_row": (C.c_int, [C.c_int] + [C.c_uint64] * 10 + [C.POINTER(kj_row), C.POINTER(C.c_int)]),
    "kj_synth_size": (C.c_int, [vp, C.POINTER(kj_synth_params), u64p]),
    "kj_synth_generate": (C.c_int, [vp, C.POINTER(kj_synth_params), vp, C.c_uint64]),
    "kj_synth_genome": (C.c_int, [vp, C.c_uint64, vp, C.c_uint64]),
}

_lib = None


def library_path() -> str:
    return _build.LIB


def lib():
    """Load (building first if the sources are newer and nvcc is present) the native library."""
    global _lib
    if _lib is None:
        path = _build.LIB
        try:
            if _build.needs_build():
                _build.build()
        except Exception as exc:  # no nvcc (GPU box): the prebuilt library must be there
            if not os.path.exists(path):
                raise RuntimeError(
                    f"libkmerjs_b200.so is missing and cannot be built ({exc}); "
                    "kmerjs_b200 has no CPU fallback") from exc
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError = header and library disagree
            fn.restype = res
            fn.argtypes = args
        if L.kj_abi_version() != KJ_ABI_VERSION:
            raise RuntimeError("libkmerjs_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, ctx=None) -> int:
    if rc < 0:
        msg = lib().kj_last_error(ctx)
        raise KjError(rc, (msg or b"").decode("utf-8", "replace") or f"kj error {rc}")
    return rc
