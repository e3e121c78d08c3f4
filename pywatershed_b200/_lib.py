"""ctypes binding of libpwsb200.so (include/pws_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).
There is no CPU path: importing works anywhere (so host-side logic can be
tested), but creating an engine without the library or without a CUDA device
raises.
"""
from __future__ import annotations

import ctypes as C
import pathlib as pl

import os

PKG = pl.Path(__file__).resolve().parent
# PWS_B200_LIB: development knob to load an alternative build of the same
# library (kernel tuning variants); the default is the in-tree build.
LIB_PATH = pl.Path(os.environ.get("PWS_B200_LIB", PKG / "libpwsb200.so"))
_lib = None


class PwsLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PwsLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc -gencode arch=compute_100a,code=sm_100a). "
            "pywatershed_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, cp, i32, i64, dbl = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "pws_abi_version": (i32, []),
        "pws_n_hru_vars": (i32, []),
        "pws_hru_var_name": (cp, [i32]),
        "pws_hru_var_kind": (i32, [i32]),
        "pws_n_seg_vars": (i32, []),
        "pws_seg_var_name": (cp, [i32]),
        "pws_n_hru_state": (i32, []),
        "pws_hru_state_name": (cp, [i32]),
        "pws_create": (i32, [C.POINTER(vp), i32, i64, i64, i32]),
        "pws_destroy": (i32, [vp]),
        "pws_last_error": (cp, [vp]),
        "pws_set_param_f64": (i32, [vp, cp, vp, i64]),
        "pws_set_param_i32": (i32, [vp, cp, vp, i64]),
        "pws_set_option": (i32, [vp, cp, dbl]),
        "pws_init": (i32, [vp]),
        "pws_get_soltab": (i32, [vp, i32, vp]),
        "pws_select_outputs": (i32, [vp, i32, vp, i32, vp]),
        "pws_n_selected": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "pws_run_device": (i32, [vp, i32] + [vp] * 8),
        "pws_run_host": (i32, [vp, i32] + [vp] * 8),
        "pws_run_host_f32": (i32, [vp, i32] + [vp] * 8),
        "pws_run_channel_device": (i32, [vp, i32, vp, vp]),
        "pws_n_inputs": (i32, []),
        "pws_input_name": (cp, [i32]),
        "pws_run_host_inputs": (i32, [vp, i32, vp, i32, vp, vp, C.c_uint32, vp, vp, vp, vp]),
        "pws_sync": (i32, [vp]),
        "pws_get_state": (i32, [vp, cp, vp]),
        "pws_set_state": (i32, [vp, cp, vp]),
        "pws_launch_count": (i64, [vp]),
        "pws_last_timing": (i32, [vp, C.POINTER(dbl), C.POINTER(dbl)]),
        "pws_column_kernel_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "pws_timer_mark": (i32, [vp, i32]),
        "pws_timer_elapsed_ms": (i32, [vp, C.POINTER(dbl)]),
        "pws_last_launches": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "pws_last_blocks": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "pws_select_accumulations": (i32, [vp, i32, vp, i32, vp]),
        "pws_get_accumulations": (i32, [vp, vp, vp]),
        "pws_clear_accumulations": (i32, [vp]),
        "pws_et_host": (i32, [i32, i32, i64] + [vp] * 11),
        "pws_n_seg_state": (i32, []),
        "pws_seg_state_name": (cp, [i32]),
        "pws_reset_state": (i32, [vp]),
        "pws_output_tile": (i32, [vp, C.POINTER(i32), C.POINTER(i64)]),
        "pws_set_flow_graph": (i32, [vp, vp, vp, i32]),
        "pws_set_flow_node_params": (i32, [vp, vp]),
        "pws_run_flow_graph_host": (i32, [vp, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "pws_abi_version", "pws_n_hru_vars", "pws_hru_var_name", "pws_hru_var_kind",
    "pws_n_seg_vars", "pws_seg_var_name", "pws_n_hru_state", "pws_hru_state_name",
    "pws_create", "pws_destroy", "pws_last_error", "pws_set_param_f64",
    "pws_set_param_i32", "pws_set_option", "pws_init", "pws_get_soltab",
    "pws_select_outputs", "pws_n_selected", "pws_run_device", "pws_run_host",
    "pws_run_channel_device", "pws_n_inputs", "pws_input_name", "pws_run_host_inputs", "pws_sync", "pws_get_state", "pws_set_state",
    "pws_launch_count", "pws_last_timing", "pws_column_kernel_info",
    "pws_timer_mark", "pws_timer_elapsed_ms", "pws_last_launches", "pws_reset_state",
    "pws_output_tile", "pws_run_host_f32", "pws_last_blocks", "pws_n_seg_state",
    "pws_seg_state_name", "pws_select_accumulations", "pws_get_accumulations",
    "pws_clear_accumulations", "pws_et_host", "pws_set_flow_graph", "pws_set_flow_node_params", "pws_run_flow_graph_host",
)
