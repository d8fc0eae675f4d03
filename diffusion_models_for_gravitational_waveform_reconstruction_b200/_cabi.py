"""ctypes binding of include/gwb200.h.  There is no fallback: if the library is missing or a call fails we raise."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_PKG, "libgwb200.so")
_lib: Optional[C.CDLL] = None

GW_F32, GW_BF16 = 0, 1


class StepParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("cfg_both", C.c_int), ("selfcond", C.c_int), ("pred_x0", C.c_int),
                ("eps_scale", C.c_float), ("dc_weight", C.c_float), ("y_dc", C.c_void_p),
                ("seed", C.c_ulonglong), ("sample0", C.c_long), ("advance", C.c_void_p), ("rng", C.c_void_p)]


class ConvTcShape(C.Structure):
    _fields_ = [("n_src", C.c_int), ("pair", C.c_int), ("B", C.c_int), ("L", C.c_int), ("C0", C.c_int),
                ("L0", C.c_int), ("C1", C.c_int), ("Cout", C.c_int)]


_P, _I, _L, _F, _U64 = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_ulonglong

_SIGS = {
    "gw_version": ([], _I),
    "gw_last_error": ([], C.c_char_p),
    "gw_device_info": ([C.POINTER(_I)] * 3, _I),
    "gw_film_vectors": ([_P, _I, _I, _F, _P, _P, _P, _P, _I, _I, _P, _P, _P], _I),
    "gw_cond_pyramid": ([_P, _I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_P), _P], _I),
    "gw_conv_in": ([_P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P], _I),
    "gw_conv_in_block": ([_P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _I, _P, _P], _I),
    "gw_conv3_simt": ([_P, _I, _I, _I, _P, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P], _I),
    "gw_gn_apply": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _I, _P], _I),
    "gw_gn_apply_stream": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P], _I),
    "gw_final_step": ([_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, C.POINTER(StepParams), _P, _P, _P, _P, _P, _P], _I),
    "gw_step_advance": ([_P, _I, _P], _I),
    "gw_philox_normal": ([_U64, _L, C.c_uint, _I, _I, _P, _P], _I),
    "gw_q_sample": ([_P, _P, _P, _P, _P, _I, _U64, _L, C.c_uint, _F, _P, _I, _I, _I, _P], _I),
    "gw_conv_tc_packed_elems": ([C.POINTER(ConvTcShape)], _L),
    "gw_conv_tc_pack": ([C.POINTER(ConvTcShape), _P, _P, _P], _I),
    "gw_conv_tc_n_part": ([C.POINTER(ConvTcShape)], _I),
    "gw_conv_tc": ([C.POINTER(ConvTcShape), _P, _P, _P, _P, _P, _P, _I, _P], _I),
    "gw_conv_in_gn_group": ([_I, _I, _I, _I], _I),
    "gw_conv_in_gn": ([_P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P], _I),
    "gw_conv_in_direct_ws_floats": ([_I, _I, _I, _I, _I], _L),
    "gw_conv_in_direct": ([_P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P], _I),
    "gw_conv_gn_group": ([C.POINTER(ConvTcShape), _I, _I], _I),
    "gw_conv_gn_sync_bytes": ([_I], _L),
    "gw_conv_gn": ([C.POINTER(ConvTcShape), _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P,
                    _P], _I),
    "gw_conv_gn2": ([C.POINTER(ConvTcShape), _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P,
                     _P, _P, _P], _I),
    "gw_conv_gn3": ([C.POINTER(ConvTcShape), _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P,
                     _P, _P, _P, _I, _P, _P], _I),
}
# training step: backward.cu / optim.cu
_SIGS.update({
    "gw_reduce_rows": ([_P, _I, _L, _F, _P, _I, _P], _I),
    "gw_loss": ([_P, _P, _P, _P, _I, _I, _I, _F, _F, _P, _P, _P, _P], _I),
    "gw_final_bwd": ([_P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P], _I),
    "gw_gn_bwd_scratch_elems": ([_I, _I, _I, _I], _L),
    "gw_gn_bwd": ([_P, _P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _L, _P, _P, _P, _P, _I, _P, _P, _L, _P, _P, _P, _P,
                   _P, _P, _P], _I),
    "gw_gn_bwd_phase": ([_P, _P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _L, _P, _P, _P, _P, _I, _P, _P, _L, _P, _P, _P, _P,
                         _P, _P, _I, _P], _I),
    "gw_gn_bwd2": ([_P, _P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _L, _P, _P, _P, _P, _I, _P, _P, _L, _P, _P, _P, _P,
                    _P, _P, _P, _P], _I),
    "gw_gn_bwd_sync_bytes": ([_I], _L),
    "gw_gn_bwd_fused_group": ([_I, _I, _I, _I, _I], _I),
    "gw_gen_conv": ([_P, _I, _I, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P], _I),
    "gw_gen_weight_dgrad": ([_P, _I, _I, _I, _P, _P], _I),
    "gw_gen_gn_stats": ([_P, _I, _I, _I, _I, _I, _P, _P], _I),
    "gw_gen_gn_apply": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _L, _L, _P, _P, _P, _I, _P], _I),
    "gw_gen_final": ([_P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, C.POINTER(StepParams), _P, _P, _P, _P, _P, _P], _I),
    "gw_gen_final_bwd": ([_P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "gw_gen_gn_bwd_scratch_floats": ([_I, _I], _L),
    "gw_gen_gn_bwd": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _I, _L, _P, _P, _I, _P, _P, _L, _P, _P, _P, _P, _P,
                       _P, _P], _I),
    "gw_gen_wgrad": ([_P, _I, _I, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "gw_gen_split_cat": ([_P, _I, _I, _I, _I, _I, _P, _P, _I, _P], _I),
    "gw_weight_dgrad": ([_P, _I, _I, _P, _P], _I),
    "gw_split_cat_grad": ([_P, _I, _I, _I, _I, _I, _P, _P, _I, _P], _I),
    "gw_wgrad3_simt": ([_P, _I, _I, _I, _P, _I, _P, _I, _I, _I, _I, _P, _L, _P, _P], _I),
    "gw_wgrad_in": ([_P, _I, _I, _I, _P, _I, _I, _P, _L, _P, _P], _I),
    "gw_wgrad_tc_scratch_elems": ([_I, _I, _I, _I, _I], _L),
    "gw_wgrad_tc_finish": ([_I, _I, _I, _I, _I, _I, _I, _P, _P, _P], _I),
    "gw_wgrad_tc": ([_I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _L, _P, _I, _P], _I),
    "gw_film_bwd": ([_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P], _I),
    "gw_batch_prepare": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    "gw_train_draws": ([_U64, _P, _L, _I, _I, _I, _F, _P, _P, _P], _I),
    "gw_train_pack": ([_P, _P, _I, _P, _P, _P, _P, _P, _I, _U64, _L, _P, _F, _I, _I, _P, _I, _I, _I, _P], _I),
    "gw_selfcond_x0": ([_P, _P, _P, _P, _I, _I, _I, _P], _I),
    "gw_opt_scratch_doubles": ([], _I),
    "gw_grad_sumsq": ([_P, _L, _P, _P], _I),
    "gw_adamw_ema": ([_P, _P, _P, _P, _P, _L, _P, _P, _P, _I, _P, C.c_double, C.c_double, _F, _P, _P], _I),
    "gw_bucket_reset": ([_P, _L, _P, _P], _I),
    "gw_train_advance": ([_P, _P, _P, _P], _I),
    "gw_loss_weight": ([_P, _P, _F, _P, _I, _P], _I),
})
_SIGS["gw_score_batch"] = ([_P, _P, _P, _I, _I, C.c_double, C.c_double, _I, C.c_double, _P, _P], _I)
_SIGS["gw_set_option"] = ([C.c_char_p, _I], _I)
_EXTRA_SIGS = {}


def exported_symbols():
    return list(_SIGS) + list(_EXTRA_SIGS)


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libgwb200.so (built in-tree by build.py).  Raises if absent: the CUDA path is the only path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(_LIB_PATH)
    for name, (args, res) in {**_SIGS, **_EXTRA_SIGS}.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    # A/B switches for measurements, e.g. GWB200_OPTIONS="pdl=0,gn_bwd_stats_fast=0" (see gw_set_option in include/gwb200.h)
    for kv in filter(None, os.environ.get("GWB200_OPTIONS", "").split(",")):
        k, v = kv.split("=")
        if lib.gw_set_option(k.strip().encode(), int(v)) != 0:
            raise RuntimeError("GWB200_OPTIONS: " + lib.gw_last_error().decode())
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().gw_last_error().decode()
        raise RuntimeError(f"gwb200 {what} failed ({rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
