"""ctypes binding of libmbseg.so (the C ABI declared in include/mbseg.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is
raised (the reference's callers catch RuntimeError around net(), src/inference/infer.py:352-356).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmbseg.so")
_lib = None

# Kernels of this library write parameters and BatchNorm buffers through raw pointers, which does not bump
# torch's tensor version counters.  Every such writer (TrainEngine.forward_backward, Ranger.step, the fused Adam
# step) calls note_raw_write(); caches keyed on parameter versions (the eval engine, unets._NetBase.engine)
# include this epoch so they are rebuilt after training steps.
_raw_write_epoch = 0


def note_raw_write():
    global _raw_write_epoch
    _raw_write_epoch += 1


def raw_write_epoch():
    return _raw_write_epoch

c_void_p, c_int, c_float, c_size_t, c_int64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t,
                                               ctypes.c_int64)


class ConvDesc(ctypes.Structure):
    """mirror of mbs_conv_desc (include/mbseg.h)"""
    _fields_ = [("mode", c_int), ("N", c_int), ("H", c_int), ("W", c_int),
                ("src0", c_void_p), ("C0", c_int), ("ld0", c_int), ("coff0", c_int),
                ("src1", c_void_p), ("C1", c_int), ("ld1", c_int), ("coff1", c_int),
                ("weight", c_void_p), ("Cout", c_int),
                ("bias", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("act", c_int),
                ("dst", c_void_p), ("ldd", c_int), ("coffd", c_int),
                ("head_w", c_void_p), ("head_b", c_float * 4), ("head_n", c_int), ("head_out", c_void_p)]


class ConvRefDesc(ctypes.Structure):
    """mirror of mbs_convref_desc (include/mbseg.h): fp32 check mode"""
    _fields_ = [("mode", c_int), ("N", c_int), ("H", c_int), ("W", c_int),
                ("src0", c_void_p), ("C0", c_int), ("src1", c_void_p), ("C1", c_int),
                ("weight", c_void_p), ("Cout", c_int),
                ("bias", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("act", c_int), ("dst", c_void_p)]


class RangerTensor(ctypes.Structure):
    """mirror of mbs_ranger_tensor (include/mbseg.h)"""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("slow", c_void_p),
                ("numel", ctypes.c_longlong), ("rows", c_int), ("row_len", c_int), ("gc", c_int), ("row_start", c_int)]


class WgradDesc(ctypes.Structure):
    """mirror of mbs_wgrad_desc (include/mbseg.h)"""
    _fields_ = [("kind", c_int), ("N", c_int), ("Ho", c_int), ("Wo", c_int),
                ("a", c_void_p), ("Cm", c_int), ("lda", c_int), ("coffa", c_int),
                ("b", c_void_p), ("Cn", c_int), ("ldb", c_int), ("coffb", c_int),
                ("out", c_void_p), ("out_ld", c_int), ("out_coff", c_int), ("partial", c_int)]


class PackJob(ctypes.Structure):
    """mirror of mbs_pack_job (include/mbseg.h)"""
    _fields_ = [("w", c_void_p), ("fwd", c_void_p), ("dgrad", c_void_p), ("cout", c_int), ("cin", c_int), ("kind", c_int),
                ("tile0", c_int)]


_SIGS = {
    "mbs_last_error": (ctypes.c_char_p, []),
    "mbs_version": (c_int, []),
    "mbs_launch_count": (c_int64, [c_int]),
    "mbs_debug_flags": (c_int, [c_int]),
    "mbs_first_conv": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "mbs_first_conv_halo64": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(ConvDesc), c_void_p]),
    "mbs_frame_minmax": (c_int, [c_void_p, c_int, ctypes.c_longlong, c_void_p, c_void_p, c_void_p]),
    "mbs_conv_gemm": (c_int, [ctypes.POINTER(ConvDesc), c_void_p]),
    "mbs_pack_conv3x3_weight": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mbs_pack_convT2x2_weight": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mbs_postproc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mbs_distance_postprocessing": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p,
                                            c_void_p, c_size_t, c_void_p, c_void_p]),
    "mbs_distance_postprocessing_sweep": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                                  c_void_p, c_size_t, c_void_p, c_void_p]),
    "mbs_boundary_postprocessing": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "mbs_softmax3_hwc": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mbs_pp_front": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                             c_void_p]),
    "mbs_pp_label8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_labels_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mbs_boundary_border_labels": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mbs_j4_labels": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mbs_labels_max_mal": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_distance_labels": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_distance_labels_ex": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_bn_train_fwd": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mbs_pack_conv3x3_dgrad": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mbs_pack_train_weights": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "mbs_unpack_conv3x3_grad": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mbs_bn_train_bwd": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mbs_bn_scratch_floats": (c_size_t, [c_int]),
    "mbs_sample_group_norm": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                      c_void_p]),
    "mbs_head_fwd": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mbs_smoothl1": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_void_p]),
    "mbs_regression_loss": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    "mbs_ce_dice_loss": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mbs_head_bwd": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mbs_maxpool2x2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mbs_zero_insert_up2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mbs_add3_bf16": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_void_p]),
    "mbs_first_conv_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mbs_conv_wgrad": (c_int, [ctypes.POINTER(WgradDesc), c_void_p]),
    "mbs_conv_wgrad_splits": (c_int, [ctypes.POINTER(WgradDesc)]),
    "mbs_wgrad_reduce": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mbs_ranger_step": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_int, c_int, c_float,
                                c_void_p]),
    "mbs_adam_step": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_float, c_int, c_void_p]),
    "mbs_label8_instances": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_pp_watershed": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p,
                                 c_int, c_void_p]),
    "mbs_instance_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_conv_ref_f32": (c_int, [ctypes.POINTER(ConvRefDesc), c_void_p]),
    "mbs_aug_warp": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mbs_aug_workspace_bytes": (c_size_t, [c_int]),
    "mbs_aug_contrast": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_aug_blur": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mbs_aug_noise_normalize": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, ctypes.c_ulonglong, c_float, c_float, c_void_p,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    "mbs_contour_first": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mbs_contour_trace": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}


def lib():
    """Load libmbseg.so; raise RuntimeError (never fall back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). microbeseg_b200 has no CPU or library fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().mbs_last_error().decode(errors="replace")
        raise RuntimeError(f"libmbseg {what} failed (code {rc}): {msg}")


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)
