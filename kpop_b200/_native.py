"""ctypes binding of include/kpopcount.h.  Loading fails loudly: there is no fallback of any kind."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkpopcount_gpu.so")

KPC_DNA_SS, KPC_DNA_DS, KPC_PROTEIN = 0, 1, 2
KPC_FASTA, KPC_FASTQ_SE, KPC_FASTQ_PE = 0, 1, 2
KPC_OK = 0
KPC_E_ARG, KPC_E_K_RANGE, KPC_E_MALFORMED_FASTQ, KPC_E_QUOTES_IN_NAME, KPC_E_IO = -1, -2, -3, -4, -5
KPC_E_CUDA, KPC_E_NOMEM, KPC_E_STATE, KPC_E_UNSUPPORTED, KPC_E_PE_MISMATCH = -6, -7, -8, -9, -10

SINK_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_char), ctypes.c_size_t)

# name -> (restype, argtypes); every symbol include/kpopcount.h declares
PROTOTYPES = {
    "kpc_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_longlong,
                                  ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "kpc_destroy": (None, [ctypes.c_void_p]),
    "kpc_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "kpc_set_sink": (ctypes.c_int, [ctypes.c_void_p, SINK_FN, ctypes.c_void_p]),
    "kpc_staging_slots": (ctypes.c_int, [ctypes.c_void_p]),
    "kpc_staging": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "kpc_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "kpc_feed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "kpc_set_record_base": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_ulonglong]),
    "kpc_hash_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_ulonglong)]),
    "kpc_hash_import": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_int]),
    "kpc_bucket_count": (ctypes.c_ulonglong, [ctypes.c_void_p]),
    "kpc_count_newlines": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_ulonglong)]),
    "kpc_dense_has_hi": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "kpc_reset_label": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "kpc_feed_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "kpc_set_pair_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong]),
    "kpc_set_single_pass": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "kpc_complete_pairs": (ctypes.c_longlong, [ctypes.c_void_p]),
    "kpc_end": (ctypes.c_int, [ctypes.c_void_p]),
    "kpc_finish": (ctypes.c_int, [ctypes.c_void_p]),
    "kpc_kmers_counted": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]),
    "kpc_dense_table": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.POINTER(ctypes.c_ulonglong)]),
    "kpc_dense_max": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]),
    "kpc_dense_promote": (ctypes.c_int, [ctypes.c_void_p]),
    "kpc_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "kpc_discard_text": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "kpc_text_bytes": (ctypes.c_ulonglong, [ctypes.c_void_p]),
    "kpc_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "kpc_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "kpc_kernel_launches": (ctypes.c_ulonglong, [ctypes.c_void_p]),
    "kpc_backend": (ctypes.c_char_p, []),
    "kpc_set_sink_buffer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "kpc_sink_buffer_used": (ctypes.c_ulonglong, [ctypes.c_void_p]),
    "kpc_profile_enable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "kpc_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                        ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]),
    "kpc_synth_fastq": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_ulonglong,
                                       ctypes.c_ulonglong]),
    "kpc_synth_offset": (ctypes.c_ulonglong, [ctypes.c_ulonglong]),
}

_lib = None


def load(path=None):
    """Load libkpopcount_gpu.so (built in-tree by kpop_b200/csrc/Makefile) and bind every entry point."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `make -C kpop_b200/csrc` (or __graft_entry__.build()). "
            "kpop_b200 has no CPU fallback.")
    lib = ctypes.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means the library and the header disagree
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
