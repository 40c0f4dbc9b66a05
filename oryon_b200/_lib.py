"""ctypes binding of liboryon_b200.so (include/oryon_b200.h).

There is no CPU or PyTorch fallback behind these calls: if the library is missing, cannot be loaded,
or the device is not sm_100, the call raises.  PyTorch is used by the callers for device memory and
streams only.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p
from typing import Dict, Optional

from . import build as _build

ABI_VERSION = 3

MATCH_TC_REFINED = 0
MATCH_EXACT_FP32 = 1

DEPTH_I32, DEPTH_F32, DEPTH_I16, DEPTH_U16 = 0, 1, 2, 3



class PointDSCConfig(ctypes.Structure):
    """oryon_pointdsc_config"""
    _fields_ = [("in_dim", c_int32), ("num_layers", c_int32), ("num_channels", c_int32), ("num_iterations", c_int32),
                ("k", c_int32), ("reserved", c_int32), ("ratio", c_double), ("sigma_d", c_double), ("sigma", c_double),
                ("nms_radius", c_double), ("inlier_threshold", c_double)]


class PointDSCDebug(ctypes.Structure):
    """oryon_pointdsc_debug"""
    _fields_ = [("conf", c_void_p), ("features", c_void_p), ("seeds", c_void_p), ("fitness", c_void_p), ("seeds_cap", c_int32),
                ("reserved", c_int32), ("initial_trans", c_void_p), ("best_seed", c_void_p)]


class BackboneConfig(ctypes.Structure):
    """oryon_backbone_config"""
    _fields_ = [("vis_layers", c_int32), ("txt_layers", c_int32), ("precision", c_int32), ("max_pairs_per_pass", c_int32)]


class BackboneDebug(ctypes.Structure):
    """oryon_backbone_debug"""
    _fields_ = [("B", c_int32), ("reserved", c_int32), ("clip_tokens", c_void_p), ("guid1", c_void_p), ("guid2", c_void_p),
                ("guid3", c_void_p), ("fusion", c_void_p)]


# name -> (restype, argtypes): every symbol include/oryon_b200.h declares
SIGNATURES = {
    "oryon_abi_version": (c_int, []),
    "oryon_last_error": (c_char_p, []),
    "oryon_create": (c_int, [c_int, POINTER(c_void_p)]),
    "oryon_destroy": (c_int, [c_void_p]),
    "oryon_workspace_bytes": (c_int64, [c_void_p]),
    "oryon_profile_enable": (c_int, [c_void_p, c_int]),
    "oryon_profile_read": (c_int, [c_void_p, POINTER(c_double), POINTER(c_int64), c_int]),
    "oryon_match_nn": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               POINTER(c_int32), POINTER(c_int32), c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "oryon_match_last_stats": (c_int, [c_void_p, POINTER(c_int64), c_void_p]),
    "oryon_match_set_hist": (c_int, [c_void_p, c_int]),
    "oryon_match_list_hist": (c_int, [c_void_p, POINTER(c_int64), c_void_p]),
    "oryon_match_plan": (c_int, [POINTER(c_int32), POINTER(c_int32), c_int, c_int, c_int, POINTER(c_int32), c_int, POINTER(c_int32),
                                 POINTER(c_int32)]),
    "oryon_mask_to_roi": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "oryon_corrs_to_pcd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, POINTER(c_double), POINTER(c_double), c_void_p, c_void_p, c_void_p, c_void_p]),
    "oryon_select_lift": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "oryon_lift_pcd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_double), c_void_p, c_void_p, c_int,
                               c_void_p, c_void_p]),
    "oryon_gemm_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_float, c_int, c_void_p]),
    "oryon_backbone_set_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "oryon_backbone_finalize": (c_int, [c_void_p, POINTER(BackboneConfig), c_void_p]),
    "oryon_text_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "oryon_backbone_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       POINTER(BackboneDebug), c_void_p]),
    "oryon_gemm_counters": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), POINTER(c_double)]),
    "oryon_mask_postproc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "oryon_stage_inputs": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p]),
    "oryon_eval_set_object": (c_int, [c_void_p, c_int, POINTER(c_double), c_int, POINTER(c_double), c_int]),
    "oryon_eval_pose_errors": (c_int, [c_void_p, c_int, POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "oryon_eval_set_object_mesh": (c_int, [c_void_p, c_int, POINTER(c_int32), c_int]),
    "oryon_eval_vsd": (c_int, [c_void_p, c_int, POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double,
                               POINTER(c_double), c_int, POINTER(c_double), c_void_p, c_void_p]),
    "oryon_pointdsc_load": (c_int, [c_void_p, POINTER(PointDSCConfig), c_void_p, c_int64, c_void_p]),
    "oryon_pointdsc_pose": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int32), c_int, c_int, c_void_p,
                                    POINTER(PointDSCDebug), c_void_p]),
}

_lock = threading.Lock()
_cdll: Optional[ctypes.CDLL] = None
_handles: Dict[int, c_void_p] = {}


class OryonError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """dlopen the library (building it first if the .so is absent and nvcc exists) and bind every symbol."""
    global _cdll
    with _lock:
        if _cdll is not None:
            return _cdll
        path = library_path()
        if not os.path.exists(path) or not _build.is_current():
            # a library built from other sources than the ones in the tree (edited csrc/, stale .so) would silently run old
            # kernels: rebuild it (under the build's file lock, so concurrent ranks are safe) or refuse
            what = "is missing" if not os.path.exists(path) else "is stale (csrc/ or the header changed since it was built)"
            if not build_if_missing:
                raise OryonError(f"{path} {what}: run `python -m oryon_b200.build` (no fallback path exists)")
            try:
                _build.build()
            except RuntimeError as e:
                raise OryonError(f"{path} {what} and cannot be rebuilt: {e}") from e
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        got = lib.oryon_abi_version()
        if got != ABI_VERSION:
            raise OryonError(f"liboryon_b200.so ABI {got} != binding ABI {ABI_VERSION}; rebuild the library")
        _cdll = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().oryon_last_error()
        raise OryonError(f"liboryon_b200 error {rc}: {msg.decode() if msg else '?'}")


def handle(device_index: int) -> c_void_p:
    """One context per (process, device), created lazily; raises on non-sm_100 devices."""
    lib = load()
    with _lock:
        h = _handles.get(device_index)
        if h is None:
            h = c_void_p()
            rc = lib.oryon_create(int(device_index), ctypes.byref(h))
            if rc != 0:
                msg = lib.oryon_last_error()
                raise OryonError(f"oryon_create({device_index}) failed ({rc}): {msg.decode() if msg else '?'}")
            _handles[device_index] = h
        return h


def destroy_all() -> None:
    lib = load()
    with _lock:
        for h in _handles.values():
            lib.oryon_destroy(h)
        _handles.clear()


KERNEL_IDS = {"prep_rows": 0, "match_tc": 1, "refine_rows": 2, "exact_rows": 3, "mask_to_roi": 4, "lift": 5,
              "pointdsc_sc": 6, "pointdsc_net": 7, "pointdsc_seeds": 8, "pointdsc_refine": 9,
              "gemm_tc": 10, "attention": 11, "norm": 12, "eltwise": 13, "im2col": 14,
              "attn_tc": 15, "transpose_v": 16}


def profile_enable(device_index: int, enable: bool) -> None:
    check(load().oryon_profile_enable(handle(device_index), int(bool(enable))))


def profile_read(device_index: int, n_ids: int = 32):
    """{kernel name or id: (total_ms, launches)} since the last read (waits for the recorded events)."""
    ms = (c_double * n_ids)()
    cnt = (c_int64 * n_ids)()
    check(load().oryon_profile_read(handle(device_index), ms, cnt, n_ids))
    names = {v: k for k, v in KERNEL_IDS.items()}
    return {names.get(i, i): (ms[i], cnt[i]) for i in range(n_ids) if cnt[i]}
