"""ctypes loader for libswiftover_cuda.so (the C ABI in include/swiftover_cuda.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
present when a context is created, this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SWO_LIB: alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("SWO_LIB") or os.path.join(HERE, "libswiftover_cuda.so")

SWO_OK, SWO_E_IO, SWO_E_PARSE, SWO_E_CUDA, SWO_E_OOM, SWO_E_ARG, SWO_E_STATE, SWO_E_CAPACITY = 0, -1, -2, -3, -4, -5, -6, -7


class SwoRow(C.Structure):
    _fields_ = [("qcid_strand", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("src", C.c_uint32)]


class SwoThick(C.Structure):
    _fields_ = [("thick_start", C.c_int32), ("thick_end", C.c_int32)]


class SwoLifted(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("n_unmatched", C.c_uint64), ("row_offset", C.POINTER(C.c_uint32)),
                ("rows", C.POINTER(SwoRow)), ("thick", C.POINTER(SwoThick))]


class SwoTextOut(C.Structure):
    _fields_ = [("lifted", C.c_void_p), ("lifted_len", C.c_size_t), ("unmatched", C.c_void_p),
                ("unmatched_len", C.c_size_t), ("n_records", C.c_uint64), ("n_rows", C.c_uint64),
                ("n_unmatched", C.c_uint64)]


class SwoMafStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_records", "n_matched", "n_unmatched", "n_multiple", "n_refchg", "n_neither")]


_lib = None


def lib():
    """Load the library once.  Raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C swiftover_b200/csrc` "
            "(or __graft_entry__.build()).  swiftover_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, i64, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_int64, C.c_uint64
    sig = {
        "swo_create": (i32, [C.POINTER(vp), C.POINTER(C.c_int), i32]),
        "swo_create_inspect": (i32, [C.POINTER(vp)]),
        "swo_destroy": (None, [vp]),
        "swo_last_error": (C.c_char_p, [vp]),
        "swo_version": (C.c_char_p, []),
        "swo_load_chain": (i32, [vp, C.c_char_p]),
        "swo_load_chain_mem": (i32, [vp, C.c_char_p, sz]),
        "swo_num_tcontigs": (i32, [vp]),
        "swo_tcontig_name": (C.c_char_p, [vp, i32]),
        "swo_tcontig_id": (i32, [vp, C.c_char_p, sz]),
        "swo_tcontig_nblocks": (i64, [vp, i32]),
        "swo_tcontig_size": (i64, [vp, i32]),
        "swo_num_blocks": (i64, [vp]),
        "swo_table_is_disjoint": (i32, [vp]),
        "swo_num_qcontigs": (i32, [vp]),
        "swo_qcontig_name": (C.c_char_p, [vp, i32]),
        "swo_num_qsizes": (i32, [vp]),
        "swo_qsize_name": (C.c_char_p, [vp, i32]),
        "swo_qsize_value": (i64, [vp, i32]),
        "swo_get_blocks": (i32, [vp, i32, i64, i64, vp, vp, vp, vp, vp]),
        "swo_lift_intervals": (i32, [vp, sz, vp, vp, vp, vp, vp, vp, C.POINTER(SwoLifted)]),
        "swo_lift_intervals_dev": (i32, [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp]),
        "swo_lift_points": (i32, [vp, sz, vp, vp, vp, vp, vp]),
        "swo_lift_bed_text": (i32, [vp, vp, sz, C.POINTER(SwoTextOut)]),
        "swo_text_num_segments": (i32, [vp]),
        "swo_text_segment": (i32, [vp, i32, i32, C.POINTER(vp), C.POINTER(sz)]),
        "swo_lift_bed_text_dev": (i32, [vp, vp, sz, vp, sz, vp, sz, vp, vp]),
        "swo_load_genome_2bit": (i32, [vp, vp, sz, vp, vp, i32]),
        "swo_lift_vcf_points_dev": (i32, [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "swo_load_genome_fasta": (i32, [vp, C.c_char_p]),
        "swo_genome_has_seq": (i32, [vp, C.c_char_p]),
        "swo_genome_seq_len": (i64, [vp, C.c_char_p]),
        "swo_lift_vcf_text": (i32, [vp, C.c_char_p, sz, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(SwoTextOut)]),
        "swo_lift_vcf_points": (i32, [vp, sz, vp, vp, vp, vp, vp, sz, vp, vp, vp, vp, vp]),
        "swo_lift_maf_records": (i32, [vp, sz, vp, vp, vp, vp, vp, vp, sz, vp, sz, vp, vp, vp, vp, vp, vp, vp]),
        "swo_lift_maf_text": (i32, [vp, C.c_char_p, sz, C.c_char_p, C.POINTER(SwoTextOut), C.POINTER(SwoMafStats)]),
        "swo_host_alloc": (vp, [sz]),
        "swo_host_free": (None, [vp]),
        "swo_set_option": (i32, [vp, C.c_char_p, i64]),
        "swo_launch_count": (u64, [vp]),
        "swo_synth_bed_lengths_dev": (i32, [vp, sz, i32, u64, vp, vp, vp, vp, vp, vp, vp, vp]),
        "swo_synth_bed_write_dev": (i32, [vp, sz, i32, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError if the symbol is missing
        f.restype, f.argtypes = res, args
    L._signatures = sig
    _lib = L
    return L
