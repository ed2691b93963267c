"""ctypes loader for libvradcuda.so (the C-ABI in include/vrad_cuda.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or no
CUDA device is present when a handle is created, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

TRI48_DTYPE = np.dtype([("n", "<f4", 3), ("d", "<f4"), ("id", "<i4"), ("e", "<f4", 6),
                        ("sel0", "u1"), ("sel1", "u1"), ("flags", "u1"), ("unused", "u1")])
assert TRI48_DTYPE.itemsize == 48

# every symbol include/vrad_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vrad_env_create", "vrad_env_destroy", "vrad_last_error", "vrad_env_set_stream", "vrad_env_set_async",
    "vrad_env_last_timing", "vrad_host_alloc", "vrad_host_free", "vrad_env_add_triangles", "vrad_env_build",
    "vrad_env_upload_tree", "vrad_env_stats", "vrad_env_download_tree", "vrad_trace4", "vrad_trace_rays",
    "vrad_test_lines", "vrad_patches_upload", "vrad_build_transfers", "vrad_transfers_upload", "vrad_transfers_info",
    "vrad_transfers_download", "vrad_transfers_download_rows", "vrad_set_sky_dirs", "vrad_direct_light", "vrad_bounce", "vrad_comm_unique_id",
    "vrad_comm_init", "vrad_version",
    "vrad_env_set_triangle_colors", "vrad_bsp_upload", "vrad_point_leafnum", "vrad_cluster_from_point",
    "vrad_sky_cameras_set", "vrad_sky_cameras_get", "vrad_test_lines_sky", "vrad_leafs_trace_to_sky",
    "vrad_decompress_vis", "vrad_pvs_from_vis_lump", "vrad_patches_subdivide", "vrad_patches_set_hierarchy", "vrad_patches_set_windings", "vrad_set_light_trace_flags",
    "vrad_light_for_string", "vrad_lights_from_entities", "vrad_lights_from_patches",
    "vrad_bump_normals", "vrad_patches_set_bump", "vrad_bounce_bump_totals",
    "vrad_env_build_fast", "vrad_kd_build_binned_host",
    "vrad_points_upload", "vrad_test_lines_indexed", "vrad_env_set_option", "vrad_env_create_multi", "vrad_transfers_layout",
]

# == vrad_face_patch in include/vrad_cuda.h
FACE_PATCH_DTYPE = np.dtype([("first_point", "<i4"), ("n_points", "<i4"), ("normal", "<f4", 3), ("plane_dist", "<f4"),
                             ("lux_scale", "<f4"), ("chop", "<f4"), ("sky", "u1"), ("no_subdivide", "u1"),
                             ("has_base_light", "u1"), ("pad", "u1")])
assert FACE_PATCH_DTYPE.itemsize == 36

# == vrad_light_entity in include/vrad_cuda.h
LIGHT_ENTITY_DTYPE = np.dtype([("classname", "<i4"), ("origin", "<f4", 3), ("light_ok", "<i4"), ("light", "<f4", 3),
                               ("has_target", "<i4"), ("target_origin", "<f4", 3), ("angles", "<f4", 3), ("pitch", "<f4"), ("angle", "<f4"),
                               ("inner_cone", "<f4"), ("cone", "<f4"), ("exponent", "<f4"),
                               ("fifty_percent_distance", "<f4"), ("zero_percent_distance", "<f4"), ("hardfalloff", "<i4"),
                               ("constant_attn", "<f4"), ("linear_attn", "<f4"), ("quadratic_attn", "<f4"), ("distance", "<f4"),
                               ("ambient_ok", "<i4"), ("ambient", "<f4", 3)])
assert LIGHT_ENTITY_DTYPE.itemsize == 124


class VradConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("flags", C.c_int)]


class VradMultiConfig(C.Structure):
    _fields_ = [("n_devices", C.c_int), ("devices", C.c_int * 8), ("flags", C.c_int)]


class VradError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"vrad status {status}: {message}")
        self.status = status


_lib = None


def load(build_if_missing: bool = True):
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing:
        path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m vrad_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.vrad_last_error.restype = C.c_char_p
    lib.vrad_version.restype = C.c_char_p
    lib.vrad_host_alloc.restype = C.c_void_p
    lib.vrad_host_alloc.argtypes = [C.c_size_t]
    lib.vrad_host_free.argtypes = [C.c_void_p]
    lib.vrad_env_destroy.argtypes = [C.c_void_p]
    lib.vrad_env_destroy.restype = None
    lib.vrad_decompress_vis.restype = C.c_int64
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise VradError(rc, load().vrad_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array, a torch tensor (host or CUDA), an int address or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class PinnedArray:
    """numpy view over a vrad_host_alloc buffer."""

    def __init__(self, shape, dtype):
        self.lib = load()
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        self.addr = self.lib.vrad_host_alloc(C.c_size_t(max(nbytes, 1)))
        if not self.addr:
            raise VradError(-3, self.lib.vrad_last_error().decode())
        buf = (C.c_char * max(nbytes, 1)).from_address(self.addr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.addr:
            self.array = None
            self.lib.vrad_host_free(C.c_void_p(self.addr))
            self.addr = None
