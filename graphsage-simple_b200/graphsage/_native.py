"""ctypes binding of libgsage_sm100.so (include/gsage.h).

This is the stub a maintainer of the reference would add: the reference is
pure Python, so its "FFI" for the hot path is exactly this file.  Loading fails loudly --
there is no fallback implementation anywhere in the package.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSAGE_LIB", os.path.join(_HERE, "lib", "libgsage_sm100.so"))

_i32, _i64, _u32, _u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
_f32, _ptr = ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); kept in the order of include/gsage.h
SIGNATURES = {
    "gs_abi_version": (_i32, []),
    "gs_strerror": (ctypes.c_char_p, [_i32]),
    "gs_sample_csr": (_i32, [_ptr, _ptr, _i32, _ptr, _i32, _ptr, _i32, _i32, _i32, _u64, _i64, _ptr,
                             _u32, _u32, _i32, _ptr, _ptr, _ptr]),
    "gs_dedup_scratch_ints": (_i32, [_i32]),
    "gs_dedup_remap": (_i32, [_ptr, _ptr, _i32, _ptr, _i32, _i32, _ptr, _ptr, _i32, _ptr, _ptr, _ptr]),
    "gs_gather_mean_fwd": (_i32, [_ptr, _i64, _i32, _ptr, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _i64, _i32, _ptr]),
    "gs_scatter_mean_bwd": (_i32, [_ptr, _i64, _i32, _i32, _ptr, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _i64, _ptr]),
    "gs_encoder_fwd": (_i32, [_ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr, _ptr, _i64, _ptr]),
    "gs_encoder_bwd_ws_floats": (_i64, [_i32, _i32, _i32]),
    "gs_encoder_bwd": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr,
                              _ptr, _ptr, _i64, _ptr, _i64, _ptr, _ptr]),
    "gs_encoder_dgrad": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr, _i64, _ptr]),
    "gs_encoder_tc_supported": (_i32, [_i32, _i32]),
    "gs_encoder_fwd_tc_ws_floats": (_i64, [_i32, _i32]),
    "gs_encoder_fwd_tc": (_i32, [_ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr, _ptr, _i64, _ptr, _ptr]),
    "gs_encoder_wgrad_tc_ws_floats": (_i64, [_i32, _i32, _i32]),
    "gs_encoder_wgrad_tc": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr,
                                   _ptr, _i64, _ptr, _ptr]),
    "gs_sage_encoder_fwd_tc": (_i32, [_ptr, _i64, _ptr, _i32, _ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _ptr,
                                      _ptr, _i64, _ptr, _ptr]),
    "gs_sage_encoder_wgrad_tc": (_i32, [_ptr, _i64, _ptr, _i32, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i32, _i32,
                                        _i32, _ptr, _ptr, _i64, _ptr, _ptr]),
    "gs_classifier_ws_floats": (_i64, [_i32, _i32, _i32]),
    "gs_classifier_xent": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i32, _i32, _i32, _f32, _ptr, _i64, _ptr,
                                  _ptr, _i64, _ptr, _i64, _ptr, _ptr]),
    "gs_head_supported": (_i32, [_i32, _i32, _i32, _i32]),
    "gs_head_ws_floats": (_i64, [_i32, _i32, _i32]),
    "gs_head_fwd_bwd": (_i32, [_ptr, _i64, _i32, _ptr, _ptr, _i32, _ptr, _ptr, _i64, _i32, _i32, _ptr, _i64, _i32,
                               _ptr, _i32, _f32, _ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64,
                               _ptr, _i64, _ptr, _ptr]),
    "gs_head_rows": (_i32, [_ptr, _i64, _i32, _ptr, _ptr, _i32, _ptr, _ptr, _i64, _i32, _i32, _ptr, _i64, _i32,
                            _ptr, _i32, _f32, _ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _ptr]),
    "gs_head_wgrad": (_i32, [_ptr, _i64, _ptr, _i64, _i32, _i32, _i32, _i32, _i32, _ptr, _ptr, _i64, _ptr, _i64,
                             _ptr, _ptr]),
    "gs_sgd_step": (_i32, [_ptr, _ptr, _f32, _i64, _ptr]),
    "gs_allreduce_sgd_blocks": (_i32, [_i64]),
    "gs_allreduce_sgd": (_i32, [_ptr, _ptr, _i64, _f32, _ptr, _ptr, _i32, _i32, _ptr, _ptr]),
    "gs_gather_rows": (_i32, [_ptr, _i64, _i32, _ptr, _i32, _ptr, _ptr, _i64, _ptr]),
    "gs_bucket_scratch_ints": (_i32, [_i32, _i32]),
    "gs_bucket_by_owner": (_i32, [_ptr, _i32, _ptr, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "gs_gather_rows_peer": (_i32, [_ptr, _i32, _i64, _i32, _ptr, _i32, _ptr, _ptr, _i64, _ptr]),
    "gs_gather_mean_fwd_peer": (_i32, [_ptr, _i32, _i64, _i32, _ptr, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _i64, _i32, _ptr]),
    "gs_sample_csr_peer": (_i32, [_ptr, _ptr, _i32, _i32, _ptr, _i32, _ptr, _i32, _i32, _i32, _u64, _i64, _ptr,
                                  _u32, _u32, _i32, _ptr, _ptr, _ptr]),
    "gs_take_all_count": (_i32, [_ptr, _ptr, _i32, _ptr, _i32, _i32, _ptr, _ptr, _ptr]),
    "gs_take_all_fill": (_i32, [_ptr, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _ptr]),
    "gs_gather_mean_ragged": (_i32, [_ptr, _i64, _i32, _ptr, _ptr, _i32, _ptr, _i64, _ptr]),
    "gs_scatter_mean_ragged": (_i32, [_ptr, _i64, _i32, _ptr, _ptr, _i32, _ptr, _i64, _ptr]),
    "gs_advance_step": (_i32, [_ptr, _ptr]),
    "gs_remap_ids": (_i32, [_ptr, _ptr, _i32, _ptr, _i32, _ptr, _ptr]),
    "gs_stage_next": (_i32, [_ptr, _i64, _i64, _ptr, _ptr, _ptr]),
}

_lib = None


def load():
    """dlopen the library once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libgsage_sm100.so not found at %s -- build it with "
                "`python graphsage-simple_b200/build.py` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the header and the .so disagree
            fn.restype = res
            fn.argtypes = args
        if lib.gs_abi_version() != 1:
            raise RuntimeError("libgsage_sm100.so ABI version mismatch")
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(code, what):
    if code != 0:
        msg = load().gs_strerror(code).decode()
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg, code))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("graphsage (B200 build) only runs on CUDA tensors; got a %s tensor" % t.device)
