"""Loader of the C-ABI shared library (include/hrbf_b200.h).

The product path is the CUDA library and nothing else: if libhrbf_b200.so is missing or
cannot be loaded this raises -- there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HRBF_B200_LIB") or os.path.join(_HERE, "libhrbf_b200.so")      # override: development builds with other tuning macros
_lib = None


class HrbfError(RuntimeError):
    pass


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class IcpOptions(C.Structure):
    _fields_ = [("use_search", C.c_int), ("search_radius", C.c_int), ("use_weight", C.c_int),
                ("dist_thres", C.c_float), ("angle_thres", C.c_float)]


class TrackStats(C.Structure):
    _fields_ = [("lastICPError", C.c_float), ("lastICPCount", C.c_float), ("lastRGBError", C.c_float),
                ("lastRGBCount", C.c_float), ("lastSO3Error", C.c_float), ("lastSO3Count", C.c_float),
                ("lastA", C.c_double * 36), ("lastb", C.c_double * 6), ("icp_iterations_run", C.c_int),
                ("kernel_launches", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HrbfError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.hrbf_last_error.restype = C.c_char_p
        L.hrbf_version.restype = C.c_char_p
        L.hrbf_launch_count.restype = C.c_ulonglong
        L.hrbf_reduce_workspace_bytes.restype = C.c_size_t
        L.hrbf_odometry_map.restype = C.c_void_p
        L.hrbf_odometry_image.restype = C.c_void_p
        L.hrbf_odometry_depth.restype = C.c_void_p
        L.hrbf_odometry_gradient.restype = C.c_void_p
        L.hrbf_odometry_candidates.restype = C.c_void_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise HrbfError(f"hrbf error {rc}: {lib().hrbf_last_error().decode()}")


def ptr(t):
    """device (or host) pointer of a torch tensor as c_void_p; None -> NULL"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)
