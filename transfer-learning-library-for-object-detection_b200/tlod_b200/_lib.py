"""ctypes binding of libtlod_b200.so (include/tlod_b200.h).

There is no fallback: if the shared library is missing or does not export a
symbol declared in the header, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtlod_b200.so")

P = c_void_p  # every pointer crosses the boundary as a raw address

# name -> (restype, argtypes); mirrors include/tlod_b200.h one to one
SIGNATURES = {
    "tlod_version": (c_int, []),
    "tlod_error_string": (ctypes.c_char_p, [c_int]),
    "tlod_launch_count": (ctypes.c_ulonglong, []),
    "tlod_profile_enable": (None, [c_int]),
    "tlod_profile_reset": (None, []),
    "tlod_profile_collect": (c_int, []),
    "tlod_profile_get": (c_int, [c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_double),
                                 ctypes.POINTER(c_longlong)]),
    "tlod_roi_align_plan_bytes": (c_size_t, [c_int, c_int]),
    "tlod_roi_align_plan": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, c_size_t, P]),
    "tlod_roi_align_forward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P,
                                       c_size_t, P]),
    "tlod_roi_align_backward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P,
                                        c_size_t, P]),
    "tlod_avgpool2x2_forward": (c_int, [P, P, c_longlong, c_int, c_int, P]),
    "tlod_avgpool2x2_backward": (c_int, [P, P, c_longlong, c_int, c_int, P]),
    "tlod_roi_align_avg_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tlod_roi_align_avg_forward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P,
                                           c_size_t, P, c_size_t, P]),
    "tlod_roi_align_avg_backward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P,
                                            c_size_t, P, c_size_t, P]),
    "tlod_roi_pool_forward": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "tlod_roi_pool_backward": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "tlod_roi_crop_forward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "tlod_roi_crop_backward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "tlod_roi_crop_pool_workspace_bytes": (c_size_t, [c_int]),
    "tlod_roi_crop_pool_forward": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P,
                                           c_size_t, P]),
    "tlod_roi_crop_pool_backward": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P,
                                            c_size_t, P]),
    "tlod_nms_workspace_bytes": (c_size_t, [c_int]),
    "tlod_nms": (c_int, [P, c_int, c_int, c_float, c_int, P, P, P, c_size_t, P]),
    "tlod_class_nms_padded_rows": (c_int, [c_int]),
    "tlod_class_nms_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tlod_class_nms": (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, c_float, P, P, P, P, P, P, P, c_size_t,
                               P]),
    "tlod_proposals_n_sorted": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "tlod_proposals_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "tlod_proposals": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                               P, P, P, P, c_size_t, P]),
    "tlod_bbox_transform_inv_clip": (c_int, [P, c_int, P, P, P, c_int, c_int, P]),
    "tlod_clip_boxes": (c_int, [P, P, c_int, c_int, c_int, P]),
    "tlod_bbox_overlaps_batch": (c_int, [P, c_int, c_int, c_int, P, c_int, P, c_int, c_int, c_int, P]),
    "tlod_bbox_transform_batch": (c_int, [P, c_int, P, P, c_int, c_int, P]),
    "tlod_anchor_labels_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tlod_anchor_labels": (c_int, [P, P, c_int, P, P, P, c_int, c_int, c_int, c_float, c_float, c_int,
                                   P, c_size_t, P]),
    "tlod_anchor_subsample_host": (c_int, [P, c_int, c_int, c_int, c_int, P, ctypes.POINTER(c_int),
                                           ctypes.POINTER(c_int)]),
    "tlod_anchor_subsample_host_ahead": (c_int, [P, c_int, c_int, c_int, c_int, P, ctypes.POINTER(c_int), P, c_int,
                                                 ctypes.POINTER(c_int)]),
    "tlod_mt_pregen": (c_int, [P, P, c_int]),
    "tlod_numpy_permutation": (c_int, [P, ctypes.POINTER(c_int), c_longlong, P]),
    "tlod_anchor_targets_finalize": (c_int, [P, P, P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_float, c_float, c_float, P]),
    "tlod_anchor_targets_finalize_dev": (c_int, [P, P, P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_int,
                                                 c_int, c_int, P, P]),
    "tlod_roi_gt_assign": (c_int, [P, c_int, c_int, P, c_int, P, P, P, c_int, c_int, c_int, P]),
    "tlod_proposal_sample_host": (c_int, [P, c_int, c_int, c_int, c_int, c_float, c_float, c_float, P,
                                          ctypes.POINTER(c_int), P, P]),
    "tlod_proposal_targets": (c_int, [P, c_int, c_int, P, c_int, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int,
                                      c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float),
                                      ctypes.POINTER(c_float), c_int, P]),
    "tlod_rpn_loss_workspace_bytes": (c_size_t, []),
    "tlod_rpn_loss_forward": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P, c_size_t, P]),
    "tlod_rpn_loss_backward": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "tlod_space_to_depth_forward": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "tlod_space_to_depth_backward": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "tlod_instance_labels": (c_int, [P, P, c_int, c_int, c_int, c_float, P]),
    "tlod_grl_backward": (c_int, [P, P, c_float, c_longlong, P]),
    "tlod_grl_backward_weighted": (c_int, [P, P, P, c_float, c_int, c_int, P]),
    "tlod_da_loss_workspace_bytes": (c_size_t, []),
    "tlod_da_loss_forward": (c_int, [P, P, P, c_int, P, c_int, c_int, c_int, c_int, P, c_size_t, P]),
    "tlod_da_image_loss_workspace_bytes": (c_size_t, []),
    "tlod_da_image_loss_forward": (c_int, [c_int, P, P, P, P, P, c_int, c_int, P, P, c_size_t, P]),
    "tlod_da_image_loss_backward": (c_int, [c_int, P, P, P, P, P, c_int, c_int, P, P, P, P, P]),
    "tlod_da_loss_backward": (c_int, [P, P, P, c_int, P, P, c_float, c_float, c_float, P, P, c_int, c_int,
                                      c_int, c_int, P]),
}


ERR_UNSUPPORTED = -3  # TLOD_ERR_UNSUPPORTED
ERR_WORKSPACE = -4    # TLOD_ERR_WORKSPACE


class TlodError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libtlod_b200.so is not built (%s). Run `python __graft_entry__.py build` or "
            "`make -C transfer-learning-library-for-object-detection_b200/csrc`. "
            "There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.tlod_error_string(int(rc))
        raise TlodError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(lib.tlod_launch_count())


def profile(enable: bool) -> None:
    """Turn per-kernel event timing on/off (bench.py)."""
    lib.tlod_profile_enable(1 if enable else 0)


def profile_reset() -> None:
    lib.tlod_profile_reset()


def profile_read() -> dict:
    """{kernel name: (total_ms, launches)}; synchronises the recorded events."""
    n = lib.tlod_profile_collect()
    out = {}
    for i in range(n):
        name = ctypes.c_char_p()
        ms = ctypes.c_double()
        cnt = c_longlong()
        check(lib.tlod_profile_get(i, ctypes.byref(name), ctypes.byref(ms), ctypes.byref(cnt)), "tlod_profile_get")
        out[name.value.decode()] = (ms.value, cnt.value)
    return out
