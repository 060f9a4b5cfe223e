"""ctypes binding of libveto_b200.so (C ABI: include/veto_b200.h).

PyTorch is used for device memory and streams only: every call passes raw ``data_ptr()`` values and the
current CUDA stream.  There is no CPU fallback — if the library is missing or the device is not an
sm_100 part, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

from . import build as _build

VETO_MAX_LAYERS = 16
ABI_VERSION = 4     # include/veto_b200.h VETO_ABI_VERSION
PREC_FP32, PREC_BF16X3, PREC_BF16, PREC_F16C8, PREC_F16 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16, "f16c8": PREC_F16C8, "f16": PREC_F16}
# f16c8 / f16 are inference modes of the encoder (include/veto_b200.h): the training step and the depth backbone of a
# module configured with them run in the bf16 mode of the same storage class
TRAIN_PRECISION = {"fp32": "fp32", "bf16x3": "bf16x3", "bf16": "bf16", "f16c8": "bf16x3", "f16": "bf16"}

_fp = c_void_p  # device pointers travel as integers


class VetoConfig(Structure):
    _fields_ = [(n, c_int32) for n in ("dim", "layers", "heads", "mlp_dim", "channels", "pool", "patch",
                                      "num_obj", "num_out", "precision")]


_W_SCALARS = ["obj_embed", "class_proj_w", "class_proj_b", "bn_weight", "bn_bias", "bn_mean", "bn_var", "pos_w",
              "pos_b", "loc_proj_w", "loc_proj_b", "cls_token", "pos_embedding", "proj_d_w", "proj_d_b",
              "proj_v_w", "proj_v_b"]
_W_LAYERS = ["ln1_w", "ln1_b", "qkv_w", "out_w", "out_b", "ln2_w", "ln2_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b"]


class VetoWeights(Structure):
    _fields_ = ([(n, c_void_p) for n in _W_SCALARS] + [(n, c_void_p * VETO_MAX_LAYERS) for n in _W_LAYERS] +
                [("rel_out_w", c_void_p), ("rel_out_b", c_void_p)])


class VetoInputs(Structure):
    _fields_ = [("n_boxes", c_int32), ("n_pairs", c_int64), ("boxes", c_void_p), ("labels", c_void_p),
                ("obj_logits", c_void_p), ("roi_rgb", c_void_p), ("roi_depth", c_void_p), ("subj", c_void_p),
                ("obj", c_void_p), ("freq_bias", c_void_p)]


class VetoOutputs(Structure):
    _fields_ = [("rel_logits", c_void_p), ("rel_features", c_void_p), ("tokens", c_void_p)]


class VetoTrainInputs(Structure):
    _fields_ = [("rel_labels", c_void_p), ("class_weight", c_void_p), ("rel_offsets", c_void_p),
                ("box_offsets", c_void_p), ("n_images", c_int32), ("p_pos_dropout", c_float),
                ("p_emb_dropout", c_float), ("p_attn_dropout", c_float), ("seed", ctypes.c_uint64),
                ("bn_momentum", c_float), ("bn_running_mean", c_void_p), ("bn_running_var", c_void_p),
                ("n_heads", c_int32), ("head_offsets", c_void_p), ("head_labels", c_void_p)]


class VetoGrads(VetoWeights):
    """veto_grads: the field order of veto_weights, non-const pointers."""


class VetoTrainOutputs(Structure):
    _fields_ = [("loss", c_void_p), ("rel_logits", c_void_p), ("grad_roi_depth", c_void_p), ("grad_roi_rgb", c_void_p)]


DEPTH_CONVS = 15


class VetoDepthWeights(Structure):
    _fields_ = [(n, c_void_p * DEPTH_CONVS) for n in ("conv_w", "bn_w", "bn_b", "bn_mean", "bn_var")]


class VetoDepthGrads(Structure):
    _fields_ = [(n, c_void_p * DEPTH_CONVS) for n in ("conv_w", "bn_w", "bn_b")]


class VetoError(RuntimeError):
    """A negative return code from libveto_b200 (the reference raises RuntimeError from AT_ERROR)."""


_PROTOS = {
    "veto_abi_version": (c_int, []),
    "veto_last_error": (c_char_p, []),
    "veto_device_check": (c_int, []),
    "veto_pairs_capacity": (c_int, [POINTER(c_int32), c_int, c_int, POINTER(c_int64)]),
    "veto_pairs_enumerate": (c_int, [POINTER(c_int32), c_int, _fp, _fp, c_int, c_int, _fp, _fp, _fp, c_void_p]),
    "veto_pairs_globalize": (c_int, [_fp, c_int64, _fp, _fp, c_int, _fp, _fp, c_void_p]),
    "veto_roi_align_forward": (c_int, [_fp, c_int, c_int, c_int, c_int, _fp, c_int, c_float, c_int, c_int, c_int, _fp,
                                       c_void_p]),
    "veto_roi_align_backward": (c_int, [_fp, _fp, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _fp,
                                        c_void_p]),
    "veto_roi_gather_forward": (c_int, [POINTER(c_void_p), POINTER(c_int32), POINTER(c_int32), POINTER(c_float), c_int,
                                        c_int, c_int, _fp, c_int, c_int, c_float, c_int, c_int, _fp, _fp, c_int, c_int,
                                        c_int, c_int, _fp, _fp, _fp, c_void_p]),
    "veto_packed_bytes": (c_size_t, [POINTER(VetoConfig)]),
    "veto_pack_weights": (c_int, [POINTER(VetoConfig), POINTER(VetoWeights), _fp, c_size_t, c_void_p]),
    "veto_workspace_bytes": (c_size_t, [POINTER(VetoConfig), c_int32, c_int64, c_int32]),
    "veto_relation_forward": (c_int, [POINTER(VetoConfig), POINTER(VetoWeights), _fp, POINTER(VetoInputs),
                                      POINTER(VetoOutputs), _fp, c_size_t, c_int32, c_void_p]),
    "veto_train_workspace_bytes": (c_size_t, [POINTER(VetoConfig), c_int32, c_int64]),
    "veto_relation_train_step": (c_int, [POINTER(VetoConfig), POINTER(VetoWeights), _fp, POINTER(VetoInputs),
                                         POINTER(VetoTrainInputs), POINTER(VetoGrads), POINTER(VetoTrainOutputs), _fp,
                                         c_size_t, c_void_p]),
    "veto_last_launch_count": (c_int64, []),
    "veto_profile_begin": (c_int, [c_void_p]),
    "veto_profile_end": (c_int, [POINTER(ctypes.c_double), POINTER(c_int64)]),
    "veto_profile_tag_name": (c_char_p, [c_int]),
    "veto_postprocess": (c_int, [_fp, c_int, _fp, _fp, _fp, _fp, c_int, c_int64, _fp, _fp, _fp, _fp, c_void_p]),
    "veto_sgg_match": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_int, c_float, _fp, _fp, c_void_p]),
    "veto_relsample_detect": (c_int, [_fp] * 10 + [POINTER(c_int32), POINTER(c_int32), c_int, c_float, c_int, c_int, c_int, c_int,
                                      ctypes.c_uint64, _fp, _fp, _fp, _fp, _fp, c_void_p]),
    "veto_meet_group_labels": (c_int, [_fp, c_int64, _fp, _fp, _fp, c_int, c_int, c_int, ctypes.c_uint64, _fp, _fp, _fp, c_void_p]),
    "veto_relsample_gtbox": (c_int, [_fp, _fp, _fp, POINTER(c_int32), c_int, c_int, c_int, ctypes.c_uint64, _fp, _fp, _fp, _fp,
                                     c_void_p]),
    "veto_postprocess_meet": (c_int, [_fp, c_int, _fp, c_int, _fp, c_int, _fp, _fp, _fp, _fp, c_int, c_int64, _fp, _fp, _fp, _fp,
                                      c_void_p]),
    "veto_postprocess_meet_vote": (c_int, [_fp, c_int, _fp, c_int, _fp, c_int, c_int, _fp, _fp, _fp, _fp, c_int, c_int64, _fp, _fp,
                                           _fp, _fp, _fp, c_void_p]),
    "veto_obj_nms_per_cls": (c_int, [_fp, _fp, _fp, POINTER(c_int32), c_int, c_int, c_float, c_int, _fp, c_void_p]),
    "veto_depth_backbone_out_size": (None, [c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "veto_depth_backbone_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "veto_depth_backbone_forward": (c_int, [c_int, POINTER(VetoDepthWeights), _fp, c_int, c_int, c_int, c_int, c_float, _fp,
                                            _fp, c_size_t, c_void_p]),
    "veto_depth_backbone_backward": (c_int, [c_int, POINTER(VetoDepthWeights), _fp, c_int, c_int, c_int,
                                             POINTER(VetoDepthGrads), _fp, c_size_t, c_void_p]),
    "veto_test_gemm": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, _fp, c_size_t, c_void_p]),
    "veto_test_gemm_tn": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, POINTER(ctypes.c_uint32), _fp, c_size_t,
                                  c_void_p]),
    "veto_test_layernorm": (c_int, [_fp, _fp, _fp, _fp, c_int64, c_void_p]),
    "veto_test_attention": (c_int, [_fp, _fp, c_int64, c_void_p]),
    "veto_test_attention_tc": (c_int, [_fp, _fp, _fp, c_int64, c_int, c_void_p]),
}

EXPORTS = tuple(_PROTOS)
_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libveto_b200.so (building it first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if build_if_missing:
        try:
            if _build.needs_build():
                _build.build_library_locked()   # every rank of a torchrun job gets here at once on a fresh checkout
        except Exception as e:  # no nvcc on the box: fall through to whatever was shipped
            if not os.path.exists(path):
                raise RuntimeError(f"libveto_b200.so is missing and cannot be built: {e}") from e
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `python -m veto_b200.build` (there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.veto_abi_version() != ABI_VERSION:
        raise RuntimeError("libveto_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().veto_last_error().decode("utf-8", "replace")
        raise VetoError(f"{what or 'libveto_b200'} failed ({rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


_device_ok = False


def require_device() -> None:
    """Raise unless the current device can run the sm_100a kernels."""
    global _device_ok
    if _device_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("veto_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    check(load().veto_device_check(), "veto_device_check")
    _device_ok = True
