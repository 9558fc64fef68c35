"""ctypes binding of libpsi_b200.so (include/psi_b200.h).  One function per exported symbol;
status codes become PsiError.  No computation happens on this side."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LIMBS = 8

# BatchedFHEPSIClient.cpp:23-38
PLAINTEXT_MODULUS = {16: 65537, 32: 4296540161, 40: 1099579260929, 48: 281474981953537}


def depth_for_E(E):
    """Multiplicative depth the client requests (BatchedFHEPSIClient.cpp:46-57)."""
    if E < 500:
        return 3
    if E < 5000:
        return 5
    return 10


class PsiError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("psi_b200 status %d: %s" % (status, message))
        self.status = status
        self.message = message


PSI_OK, PSI_ERR_INVALID, PSI_ERR_NO_DEVICE, PSI_ERR_CUDA, PSI_ERR_STATE = 0, 1, 2, 3, 4

_U64x8 = ctypes.c_uint64 * MAX_LIMBS
_U64x9 = ctypes.c_uint64 * (MAX_LIMBS + 1)
_F64x8 = ctypes.c_double * MAX_LIMBS


class PsiParams(ctypes.Structure):
    """struct psi_params"""
    _fields_ = [
        ("N", ctypes.c_uint32), ("L", ctypes.c_uint32), ("Lp", ctypes.c_uint32),
        ("mult_technique", ctypes.c_uint32), ("ks_technique", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
        ("t", ctypes.c_uint64),
        ("q", _U64x8), ("p", _U64x8), ("psi_q", _U64x8), ("psi_p", _U64x8), ("psi_t", ctypes.c_uint64),
        ("QHatInvModq", _U64x8), ("QHatModp", _U64x8 * MAX_LIMBS), ("alphaQModp", _U64x8 * (MAX_LIMBS + 1)),
        ("qInv", _F64x8),
        ("negPQHatInvModq", _U64x8), ("qInvModp", _U64x8 * MAX_LIMBS), ("PHatInvModp", _U64x8),
        ("PHatModq", _U64x8 * MAX_LIMBS), ("alphaPModq", _U64x8 * (MAX_LIMBS + 1)), ("pInv", _F64x8),
        ("tQSHatInvModsDivsModq", _U64x9 * MAX_LIMBS), ("tQSHatInvModsDivsFrac", _F64x8),
        ("fp_contract", ctypes.c_uint32), ("ks_num_parts", ctypes.c_uint32), ("Lk", ctypes.c_uint32),
        ("reserved2", ctypes.c_uint32), ("pk", _U64x8), ("psi_pk", _U64x8),
        ("tPSHatInvModsDivsModp", _U64x9 * MAX_LIMBS), ("tPSHatInvModsDivsFrac", _F64x8),
    ]


_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_vp = ctypes.c_void_p
_vpp = ctypes.POINTER(ctypes.c_void_p)
_pp = ctypes.POINTER(PsiParams)
_u32, _u64, _int, _sz = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, ctypes.c_size_t

# every symbol include/psi_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "psi_params_generate": (_int, [_u32, _u64, _u32, _u32, _pp]),
    "psi_params_generate_ex": (_int, [_u32, _u64, _u32, _u32, _u32, _u32, _u32, _pp]),
    "psi_ctx_create": (_int, [_pp, _int, _vpp]),
    "psi_ctx_destroy": (_int, [_vp]),
    "psi_set_relin_key": (_int, [_vp, _u64p, _u64p]),
    "psi_db_load_limbs": (_int, [_vp, _u32, _u32, _u32, _u64p, _u64p]),
    "psi_db_encode_slots": (_int, [_vp, _u32, _u32, _u32, _u32, _i64p, _i64p]),
    "psi_db_get_limbs": (_int, [_vp, _u64p, _u64p]),
    "psi_query_set": (_int, [_vp, _u64p, _u64p, _vp]),
    "psi_query_upload": (_int, [_vp, _u64p, _u64p, _vp]),
    "psi_query_commit": (_int, [_vp, _vp]),
    "psi_query_landing_ptr": (_int, [_vp, _u32, _vpp, ctypes.POINTER(_sz), _vpp, ctypes.POINTER(_sz)]),
    "psi_query_next_landing": (_int, [_vp, _u32p]),
    "psi_query_uploaded": (_int, [_vp, _u32]),
    "psi_run": (_int, [_vp, _vp]),
    "psi_query_run_streamed": (_int, [_vp, _u64p, _u64p, _u64p, _vp]),
    "psi_run_phases": (_int, [_vp, _u32, _vp]),
    "psi_result_get": (_int, [_vp, _u64p, _vp]),
    "psi_stream_sync": (_int, [_vp]),
    "psi_run_launch_count": (_int, [_vp, _u32p]),
    "psi_result_device_ptr": (_int, [_vp, _vpp, ctypes.POINTER(_sz)]),
    "psi_debug_ntt": (_int, [_vp, _u64p, _u32p, _u32, _int]),
    "psi_debug_mul_ctct": (_int, [_vp, _u64p, _u64p, _u64p]),
    "psi_debug_set_tuning": (_int, [_vp, _int, _int]),
    "psi_bench_imad_peak": (_int, [_int, ctypes.POINTER(ctypes.c_double)]),
    "psi_bench_pipe_peak": (_int, [_int, _int, ctypes.POINTER(ctypes.c_double)]),
    "psi_hct_create": (_int, [_u64, _u32, _u64, _u32, _u64, _u64, _u64, _int, _int, _u64, _vpp]),
    "psi_hct_insert_all": (_int, [_vp, _u64p, _sz]),
    "psi_hct_get_cells": (_int, [_vp, _u64p]),
    "psi_hct_destroy": (_int, [_vp]),
    "psi_hct_build_device": (_int, [_vp, _u64, _u32, _u64, _u32, _u64, _u64, _u64, _u64p, _sz, _u64p]),
    "psi_db_build_from_items": (_int, [_vp, _u64, _u32, _u64, _u32, _u64, _u64, _u64, _u64p, _sz, _u64, _u64]),
    "psi_hash_index": (_int, [_u64, _u32, _u64p, _sz, _u32, _u32, _u64p]),
    "psi_client_table": (_int, [_u64, _u32, _u64, _u32, _u64p, _sz, _u64, _u64p]),
    "psi_random_data_input": (_int, [_sz, _sz, _sz, _u64, _u64, _u64p, _u64p, _u64p]),
    "psi_pie_create": (_int, [_vp, _pp, _vp, _u64, _u64, _int, _vpp]),
    "psi_pie_dims": (_int, [_vp, _u32p, _u32p, _u32p, _u32p]),
    "psi_pie_get_slots": (_int, [_vp, _i64p, _i64p]),
    "psi_pie_set_query": (_int, [_vp, _u64p, _u64p]),
    "psi_pie_run": (_int, [_vp]),
    "psi_pie_get_result_list": (_int, [_vp, _u64p]),
    "psi_pie_destroy": (_int, [_vp]),
    "psi_set_encode_lift": (_int, [_vp, _u32]),
    "psi_set_packing_cofactor": (_int, [_vp, _u32]),
    "psi_params_from_moduli": (_int, [_u32, _u64, _u32, _u64p, _u64p, _u32, _u64p, _u64p, _u64, _pp]),
    "psi_query_upload_limbs": (_int, [_vp, ctypes.POINTER(_u64p), ctypes.POINTER(_u64p), _vp]),
    "psi_result_get_limbs": (_int, [_vp, ctypes.POINTER(_u64p), _vp]),
    "psi_query_run_streamed_limbs": (_int, [_vp, ctypes.POINTER(_u64p), ctypes.POINTER(_u64p), ctypes.POINTER(_u64p), _vp]),
    "psi_set_host_threads": (_int, [_vp, _int]),
    "psi_db_load_limbs_shard": (_int, [_vp, _u32, _u32, _u32, _u32, _u32, _u64p, _u64p]),
    "psi_db_encode_slots_shard": (_int, [_vp, _u32, _u32, _u32, _u32, _u32, _u32, _i64p, _i64p]),
    "psi_db_build_from_items_shard": (_int, [_vp, _u64, _u32, _u64, _u32, _u64, _u64, _u64, _u64p, _sz, _u64, _u64, _u32, _u32]),
    "psi_db_get_bin_limbs": (_int, [_vp, _u32, _u64p, _u64p]),
    "psi_multi_create": (_int, [_pp, ctypes.POINTER(_int), _u32, _vpp]),
    "psi_multi_destroy": (_int, [_vp]),
    "psi_multi_device_count": (_int, [_vp, _u32p]),
    "psi_multi_bin_range": (_int, [_vp, _u32, _u32p, _u32p]),
    "psi_multi_ctx": (_int, [_vp, _u32, _vpp]),
    "psi_multi_set_encode_lift": (_int, [_vp, _u32]),
    "psi_multi_set_host_threads": (_int, [_vp, _int]),
    "psi_multi_set_relin_key": (_int, [_vp, _u64p, _u64p]),
    "psi_multi_db_load_limbs": (_int, [_vp, _u32, _u32, _u32, _u64p, _u64p]),
    "psi_multi_db_encode_slots": (_int, [_vp, _u32, _u32, _u32, _u32, _i64p, _i64p]),
    "psi_multi_db_build_from_items": (_int, [_vp, _u64, _u32, _u64, _u32, _u64, _u64, _u64, _u64p, _sz, _u64, _u64]),
    "psi_multi_query_set": (_int, [_vp, _u64p, _u64p]),
    "psi_multi_query_set_limbs": (_int, [_vp, ctypes.POINTER(_u64p), ctypes.POINTER(_u64p)]),
    "psi_multi_run": (_int, [_vp]),
    "psi_multi_result_get": (_int, [_vp, _u64p]),
    "psi_multi_result_get_limbs": (_int, [_vp, ctypes.POINTER(_u64p)]),
    "psi_multi_query_run_limbs": (_int, [_vp, ctypes.POINTER(_u64p), ctypes.POINTER(_u64p), ctypes.POINTER(_u64p)]),
    "psi_multi_sync": (_int, [_vp]),
    "psi_multi_run_launch_count": (_int, [_vp, _u32p]),
    "psi_pie_create_multi": (_int, [_vp, _pp, _vp, _u64, _u64, _int, _vpp]),
    "psi_nb_eval_sum_indices": (_int, [_u32, _u32, _u64p, _u32p]),
    "psi_nb_rotation_index": (_int, [_u32, ctypes.c_int64, _u64p]),
    "psi_nb_set_automorphism_keys": (_int, [_vp, _u32, _u64p, _u64p, _u64p]),
    "psi_nb_db_load_limbs": (_int, [_vp, _u32, _u32, _u32, _u64p, _u64p, _u64p]),
    "psi_nb_db_encode_slots": (_int, [_vp, _u32, _u32, _u32, _u32, _i64p, _i64p]),
    "psi_nb_db_get_limbs": (_int, [_vp, _u64p, _u64p, _u64p]),
    "psi_nb_run": (_int, [_vp, _u32, _u32, _u64p, _u64p, _vp]),
    "psi_nb_launch_count": (_int, [_vp, _u32p]),
    "psi_multi_nb_set_automorphism_keys": (_int, [_vp, _u32, _u64p, _u64p, _u64p]),
    "psi_multi_nb_db_encode_slots": (_int, [_vp, _u32, _u32, _u32, _u32, _i64p, _i64p]),
    "psi_multi_nb_db_load_limbs": (_int, [_vp, _u32, _u32, _u32, _u64p, _u64p, _u64p]),
    "psi_multi_nb_pie_range": (_int, [_vp, _u32, _u32p, _u32p]),
    "psi_multi_nb_run": (_int, [_vp, _u64p, _u64p]),
    "psi_device_count": (_int, [ctypes.POINTER(_int)]),
    "psi_last_error": (ctypes.c_char_p, []),
    "psi_version": (ctypes.c_char_p, []),
}

_LIB = None


def lib_path():
    # PSI_B200_LIB selects an experiment build of the same library (Makefile: BUILD= LIB= EXTRA=)
    return os.environ.get("PSI_B200_LIB") or os.path.join(_HERE, "libpsi_b200.so")


def build_library(force=False):
    """nvcc build of libpsi_b200.so for sm_100a (cross-compiles without a GPU)."""
    args = ["make", "-C", _HERE, "-s", "-j8"]
    if force:
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean"])
    subprocess.check_call(args)
    return lib_path()


def lib():
    """The loaded library.  Missing .so is an error, never a fallback."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise PsiError(PSI_ERR_NO_DEVICE, "libpsi_b200.so is not built (run __graft_entry__.build()); "
                                              "there is no CPU implementation to fall back to")
        L = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != PSI_OK:
        raise PsiError(rc, lib().psi_last_error().decode("utf-8", "replace"))


MULT_HPS, MULT_HPSPOVERQ = 0, 1
KS_BV, KS_HYBRID = 0, 1
FP_SEPARATE, FP_FMA = 0, 1


def params_generate(N, t, depth, L_override=0, mult_technique=MULT_HPSPOVERQ, ks_technique=KS_BV, fp_contract=FP_SEPARATE):
    """psi_params_generate(_ex): stand-alone BFV-RNS context tables (host only, no device needed)."""
    p = PsiParams()
    check(lib().psi_params_generate_ex(N, t, depth, L_override, mult_technique, ks_technique, fp_contract, ctypes.byref(p)))
    return p
