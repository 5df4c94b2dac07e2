"""ctypes view of the C ABI in include/msmgpu.h (newmsm_b200/lib/libmsmgpu.so).

This is the only way Python reaches the CUDA path; there is no CPU fallback. Loading fails
loudly when the library has not been built (`python -m newmsm_b200.build`), and every compute
call returns MSMGPU_ERR_CUDA -> :class:`MsmGpuError` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmsmgpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "msmgpu.h")

OK, ERR_CUDA, ERR_INVALID, ERR_OUT_OF_BOX, ERR_NO_TRIANGLE, ERR_CAPACITY = range(6)

_vp, _i, _d, _i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
_pp = C.POINTER(C.c_void_p)


class MsmGpuError(RuntimeError):
    """Mirrors newresampler::MeshException / newmeshreg::MeshregException: carries the reference's message."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[msmgpu status {status}] {message}")
        self.status = status
        self.message = message


_SIGNATURES = {
    # name: (restype, argtypes)
    "msmgpu_last_error": (C.c_char_p, []),
    "msmgpu_version": (C.c_char_p, []),
    "msmgpu_device_count": (_i, []),
    "msmgpu_launch_count": (C.c_ulonglong, []),
    "msmgpu_debug_take_cuda_error": (C.c_char_p, []),
    "msmgpu_set_query_group": (_i, [_i]),
    "msmgpu_get_query_group": (_i, []),
    "msmgpu_set_tuning": (_i, [C.c_char_p, _i]),
    "msmgpu_mean_vertex_distance": (_i, [_i, _vp, _i, _vp, _vp]),
    "msmgpu_rigid_create": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _d, _pp]),
    "msmgpu_rigid_destroy": (None, [_vp]),
    "msmgpu_rigid_cost": (_i, [_vp, _vp, _d, _d, _d, _vp]),
    "msmgpu_fwd_apply_batch_f32_dev": (_i, [_vp, _vp, _i, _vp, _vp]),
    "msmgpu_ctx_create": (_i, [_i, _vp, _pp]),
    "msmgpu_ctx_destroy": (None, [_vp]),
    "msmgpu_ctx_sync": (_i, [_vp]),
    "msmgpu_ctx_stream": (_vp, [_vp]),
    "msmgpu_device_malloc": (_i, [_vp, C.c_size_t, _pp]),
    "msmgpu_host_alloc": (_i, [_vp, C.c_size_t, _pp]),
    "msmgpu_resample_batch_host_f32": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _i]),
    "msmgpu_variance_normalise": (_i, [_vp, _i, _i, _vp, _vp]),
    "msmgpu_host_free": (None, [_vp, _vp]),
    "msmgpu_device_free": (None, [_vp, _vp]),
    "msmgpu_device_download": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "msmgpu_device_copy_peer": (_i, [_vp, _vp, _vp, _vp, C.c_size_t]),
    "msmgpu_mesh_create": (_i, [_vp, _i, _vp, _i, _vp, _pp]),
    "msmgpu_mesh_create_dev": (_i, [_vp, _i, _vp, _i, _vp, _pp]),
    "msmgpu_mesh_create_view_batch": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "msmgpu_mesh_set_coords": (_i, [_vp, _vp]),
    "msmgpu_mesh_destroy": (None, [_vp]),
    "msmgpu_mesh_shape": (_i, [_vp, _vp, _vp]),
    "msmgpu_mesh_set_features_f32": (_i, [_vp, _i, _vp]),
    "msmgpu_mesh_bary_resample_f32": (_i, [_vp, _i, _vp, _vp]),
    "msmgpu_mesh_metric_resample_f32": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "msmgpu_mesh_vertex_areas": (_i, [_vp, _vp]),
    "msmgpu_mesh_set_area_source": (_i, [_vp, _vp]),
    "msmgpu_mesh_set_triangle_areas": (_i, [_vp, _vp]),
    "msmgpu_octree_build": (_i, [_vp, _pp]),
    "msmgpu_octree_build_batch": (_i, [_vp, _i, _vp, _vp]),
    "msmgpu_octree_destroy": (None, [_vp]),
    "msmgpu_octree_stats": (_i, [_vp, _vp, _vp, _vp]),
    "msmgpu_octree_dump": (_i, [_vp, _vp, _vp, _vp]),
    "msmgpu_nearest_triangle": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_nearest_triangle_dev": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_bary_weights": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_bary_weights_dev": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "msmgpu_adaptive_weights": (_i, [_vp, _vp, _pp]),
    "msmgpu_adaptive_weights_ex": (_i, [_vp, _vp, _vp, _vp, _pp]),
    "msmgpu_adaptive_weights_batch": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "msmgpu_weights_apply_batch_f32_dev": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "msmgpu_weights_shape": (_i, [_vp, _vp, _vp, _vp]),
    "msmgpu_weights_export": (_i, [_vp, _vp, _vp, _vp]),
    "msmgpu_weights_destroy": (None, [_vp]),
    "msmgpu_weights_apply_f32_dev": (_i, [_vp, _i, _vp, _vp]),
    "msmgpu_metric_resample": (_i, [_vp, _vp, _i, _vp, _vp]),
    "msmgpu_metric_resample_f32": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "msmgpu_bary_resample_f32_dev": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp]),
    "msmgpu_bary_resample_batch_f32_dev": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "msmgpu_bary_resample_f32": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "msmgpu_bary_resample": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "msmgpu_sphere_project_warp": (_i, [_vp, _vp, _i, _vp, _vp]),
    "msmgpu_surface_resample": (_i, [_vp, _vp, _i, _vp, _vp]),
    "msmgpu_nn_resample": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "msmgpu_rotation_matrices": (_i, [_vp, _i, _vp, _vp, _vp]),
    "msmgpu_smooth_data": (_i, [_vp, _i, _vp, _vp, _d, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "msmgpu_adaptive_weights_excl": (_i, [_vp, _vp, _vp, _pp]),
    "msmgpu_metric_resample_excl": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_nn_resample_excl": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_smooth_neighbourhoods": (_i, [_vp, _i, _vp, _vp, _d, _vp, C.c_int64, _vp, _vp]),
    "msmgpu_costfn_create": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _pp]),
    "msmgpu_costfn_destroy": (None, [_vp]),
    "msmgpu_costfn_reset_source": (_i, [_vp, _vp]),
    "msmgpu_costfn_set_percentile": (_i, [_vp, _d]),
    "msmgpu_fwd_create": (_i, [_vp, _i, _i, _pp]),
    "msmgpu_fwd_destroy": (None, [_vp]),
    "msmgpu_bary_resample_batch_f32_dev_keep": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_adaptive_weights_batch_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "msmgpu_costfn_set_cpgrid": (_i, [_vp, _i, _vp, _vp, _d, _i, _vp, _vp]),
    "msmgpu_costfn_patches": (_i, [_vp, _vp, _vp]),
    "msmgpu_costfn_unary_table": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_costfn_unary_table_dev": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_weights_apply_batch_f64_dev": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "msmgpu_device_pow_enabled": (_i, []),
    "msmgpu_debug_device_pow": (_i, [_vp, _i, _vp, _vp, _vp]),
    "msmgpu_group_fields": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "msmgpu_group_create": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _pp]),
    "msmgpu_group_destroy": (None, [_vp]),
    "msmgpu_group_set_mask": (_i, [_vp, _vp]),
    "msmgpu_group_pair_costs": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "msmgpu_group_pair_batch": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "msmgpu_group_set_pairs": (_i, [_vp, _i, _vp]),
    "msmgpu_group_pair_batch_dev": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "msmgpu_costfn_set_cpgrid_ho": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "msmgpu_costfn_triplet_costs": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "msmgpu_costfn_triplet_batch": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "msmgpu_group_triplet_costs": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _d, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "msmgpu_costfn_set_anatomical": (_i, [_vp, _i, _vp]),
    "msmgpu_triplet_plan_create": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _pp]),
    "msmgpu_triplet_plan_destroy": (None, [_vp]),
    "msmgpu_triplet_plan_batch": (_i, [_vp, _vp, _d, _i, _i, _i, _vp, _i, _vp]),
    "msmgpu_triplet_plan_batch_dev": (_i, [_vp, _vp, _d, _i, _i, _i, _vp, _i, _vp]),
    "msmgpu_group_triplet_batch": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _d, _i, _vp, _i, _vp]),
}


class RegParams(C.Structure):
    """msmgpu_reg_params (include/msmgpu.h)"""
    _fields_ = [("lambda_", C.c_double), ("shear_modulus", C.c_double), ("bulk_modulus", C.c_double),
                ("k_exponent", C.c_double), ("exponent", C.c_double), ("rmode", C.c_int)]


class Anatomical(C.Structure):
    """msmgpu_anatomical (include/msmgpu.h): the anatomical meshes and maps of regoption 4/5"""
    _fields_ = [("n_av", C.c_int), ("asource_xyz", C.c_void_p), ("n_at", C.c_int), ("asource_tri", C.c_void_p),
                ("n_hv", C.c_int), ("thi_xyz", C.c_void_p), ("n_ht", C.c_int), ("thi_tri", C.c_void_p), ("atarget_xyz", C.c_void_p),
                ("face_ptr", C.c_void_p), ("face_ids", C.c_void_p), ("bary_ptr", C.c_void_p), ("bary_key", C.c_void_p), ("bary_w", C.c_void_p)]


_lib = None


def declared_symbols():
    """Every function name include/msmgpu.h declares (parsed from the header text)."""
    import re
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(msmgpu_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m newmsm_b200.build` "
                               "(there is no CPU fallback for the CUDA path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)   # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


_DEBUG = bool(os.environ.get("MSMGPU_DEBUG"))


def check(status: int) -> None:
    if status != OK:
        raise MsmGpuError(status, lib().msmgpu_last_error().decode(errors="replace"))
    if _DEBUG:
        pending = lib().msmgpu_debug_take_cuda_error()
        if pending:
            raise MsmGpuError(ERR_CUDA, "pending CUDA error after a successful call: " + pending.decode())


def ptr(a):
    """numpy array / torch tensor / int / None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(_vp)
    if isinstance(a, int):
        return _vp(a)
    if hasattr(a, "data_ptr"):
        return _vp(a.data_ptr())
    raise TypeError(type(a))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def handle_array(handles):
    arr = (C.c_void_p * len(handles))(*[h.value if isinstance(h, C.c_void_p) else h for h in handles])
    return arr
