"""ctypes binding of libbader_b200.so (the C ABI in include/bader_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
present, every compute entry raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libbader_b200.so")

_i64 = ctypes.c_int64
_f64 = ctypes.c_double
_int = ctypes.c_int
_p = ctypes.c_void_p
_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (argtypes, restype); every symbol include/bader_b200.h declares
SIGNATURES = {
    "bdr_last_error": ([], ctypes.c_char_p),
    "bdr_version": ([], _int),
    "bdr_device_count": ([ctypes.POINTER(_int)], _int),
    "bdr_create": ([_int, _i64, _i64, _i64, _pp], _int),
    "bdr_destroy": ([_p], _int),
    "bdr_synchronize": ([_p], _int),
    "bdr_upload_density": ([_p, _int, _p], _int),
    "bdr_download_density": ([_p, _int, _p], _int),
    "bdr_alias_density": ([_p, _int, _int], _int),
    "bdr_copy_density": ([_p, _int, _int], _int),
    "bdr_upload_labels": ([_p, _int, _p, _int], _int),
    "bdr_download_labels": ([_p, _int, _p, _int], _int),
    "bdr_download_known": ([_p, _p], _int),
    "bdr_clear_labels": ([_p, _int], _int),
    "bdr_vacuum_assign": ([_p, _f64, _f64, _int, ctypes.POINTER(_f64), ctypes.POINTER(_f64)], _int),
    "bdr_vacuum_count": ([_p, ctypes.POINTER(_i64)], _int),
    "bdr_bader_calc": ([_p, _int, _p, _p, ctypes.POINTER(_i64)], _int),
    "bdr_get_maxima": ([_p, _p, _i64], _int),
    "bdr_refine": ([_p, _int, _int, _i64, _p, _p, ctypes.POINTER(_i64), _p, _i64], _int),
    "bdr_edge_find": ([_p, _int, ctypes.POINTER(_i64)], _int),
    "bdr_charge_sum": ([_p, _int, _int, _f64, _i64, _p, _p], _int),
    "bdr_assign_atoms": ([_p, _p, _i64, _p, _i64, _p, _p, _p], _int),
    "bdr_surface_distance": ([_p, _int, _p, _p, _i64, _p, ctypes.POINTER(_int)], _int),
    "bdr_volume_mask": ([_p, _int, _int, _i64, _p], _int),
    "bdr_run": ([_p, _p, _f64, _f64, _int, _int, _i64, _p, _p, _p, _int,
                 ctypes.POINTER(_i64), _p, _i64, _p, _p], _int),
    "bdr_profile_enable": ([_p, _int], _int),
    "bdr_profile_reset": ([_p], _int),
    "bdr_profile_get": ([_p, _int, ctypes.POINTER(_f64), ctypes.POINTER(_i64)], _int),
    "bdr_launch_count": ([_p, ctypes.POINTER(_i64)], _int),
    "bdr_sync_count": ([_p, ctypes.POINTER(_i64)], _int),
    "bdr_timer_start": ([_p], _int),
    "bdr_timer_stop": ([_p, ctypes.POINTER(_f64)], _int),
    "bdr_trace_steps": ([_p, ctypes.POINTER(_i64), ctypes.POINTER(_i64)], _int),
    "bdr_synth_separable": ([_p, _int, _p, _p, _p, _i64], _int),
    "bdr_synth_general": ([_p, _int, _p, _p, _p, _p, _i64], _int),
    "bdr_slab_create": ([_int, _i64, _i64, _i64, _int, _pp], _int),
    "bdr_slab_seed": ([_p, _p, ctypes.POINTER(_i64), ctypes.POINTER(_i64)], _int),
    "bdr_slab_roots": ([_p, _p, _i64], _int),
    "bdr_slab_first_voxel": ([_p, _i64, _p], _int),
    "bdr_slab_apply_rank": ([_p, _p], _int),
    "bdr_slab_first_voxel_labels": ([_p, _i64, _p], _int),
    "bdr_slab_relabel": ([_p, _int, _p], _int),
    "bdr_slab_first_pass": ([_p, _int, ctypes.POINTER(_i64)], _int),
    "bdr_slab_trace": ([_p, _int, _p, _p, _int, ctypes.POINTER(_i64)], _int),
    "bdr_slab_requeue": ([_p, _int, _p, _i64, ctypes.POINTER(_i64)], _int),
    "bdr_slab_surface_distance": ([_p, _int, _p, _p, _i64, _p, _p, ctypes.POINTER(_i64)], _int),
    "bdr_slab_ipc_export": ([_p, _p], _int),
    "bdr_slab_ipc_attach": ([_p, _int, _int, _p, _p, _i64], _int),
    "bdr_edge_pass": ([_p, _int, ctypes.POINTER(_i64)], _int),
    "bdr_trace_pass": ([_p, _int, _p, _p, ctypes.POINTER(_i64), ctypes.POINTER(_i64)], _int),
    "bdr_trace_pass_list": ([_p, _int, _p, _p, _int, ctypes.POINTER(_i64), ctypes.POINTER(_i64)], _int),
    "bdr_slab_ec_begin": ([_p, _int], _int),
    "bdr_slab_ec_round": ([_p, ctypes.POINTER(_i64)], _int),
    "bdr_slab_ec_finish": ([_p, _int, ctypes.POINTER(_i64)], _int),
    "bdr_set_stream": ([_p, _p], _int),
    "bdr_slab_comm_id": ([_p], _int),
    "bdr_slab_comm_init": ([_p, _int, _int, _p], _int),
    "bdr_slab_exchange": ([_p, _int], _int),
    "bdr_slab_rounds": ([_p, _int, _p, _p, _i64, _p, _i64, ctypes.POINTER(_i64), ctypes.POINTER(_int)], _int),
    "bdr_slab_refine": ([_p, _int, _int, _i64, _p, _p, ctypes.POINTER(_i64), _p, _i64], _int),
    "bdr_selftest_div": ([_p, _i64, ctypes.c_uint64, ctypes.POINTER(_i64)], _int),
    "bdr_set_option": ([_p, _int, _i64], _int),
    "bdr_device_ptr": ([_p, _int, _pp], _int),
    "bdr_parse_text": ([_int, _p, _i64, _i64, _i64, _i64, _i64, _int, _int, _f64, _p,
                        ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(_i64), _p, _i64], _int),
    "bdr_parse_token_host": ([ctypes.c_char_p, _i64, ctypes.POINTER(_f64)], _int),
    "bdr_parse_release": ([_int], _int),
    "bdr_format_grid": ([ctypes.c_char_p, _p, _i64, _i64, _i64, _int, _i64, _int, _int, _int], _int),
    "bdr_host_alloc": ([_i64], _p),
    "bdr_host_free": ([_p], _int),
    "bdr_host_hash": ([_p, _i64, _int, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(_int)], _int),
}

_lib = None


class BaderB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise BaderB200Error(
                f"{SO_PATH} is missing: build it with `python -m pybader_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(SO_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise BaderB200Error(load().bdr_last_error().decode("utf-8", "replace"))
