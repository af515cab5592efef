"""ctypes binding of libslide_b200.so (the C ABI declared in include/slide_b200.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLIDE_B200_LIB") or os.path.join(_HERE, "libslide_b200.so")  # override: kernel A/B runs

_lib = None

ERRORS = {-1: "SLIDE_ERR_INVALID", -2: "SLIDE_ERR_CUDA", -3: "SLIDE_ERR_UNSUPPORTED"}

# every symbol include/slide_b200.h and include/slide_sap.h declare (tests check that the library exports all of them)
SYMBOLS = [
    "slide_abi_version", "slide_last_cuda_error", "slide_launch_count", "slide_reset_launch_count",
    "slide_furthest_point_sampling", "slide_gather_points", "slide_gather_points_grad", "slide_ball_query",
    "slide_group_points", "slide_group_points_grad", "slide_three_nn", "slide_three_interpolate",
    "slide_three_interpolate_grad", "slide_knn_points", "slide_sample_farthest_points",
    "slide_furthest_point_sampling_ws", "slide_sample_farthest_points_ws", "slide_fps_resident_max_points",
    "slide_program_create", "slide_program_destroy", "slide_program_arena", "slide_program_weights",
    "slide_program_run", "slide_program_capture", "slide_program_replay", "slide_program_launches",
    "slide_program_set_gemm_backend", "slide_tc_error", "slide_tc_reset_error", "slide_tc_reload_tuning",
    "slide_program_set_resident", "slide_program_use_resident", "slide_philox_normal_slice",
    # include/slide_sap.h
    "slide_sap_mirror_concat", "slide_sap_unit_cube", "slide_dpsr_workspace_bytes", "slide_dpsr_forward",
    "slide_mc_workspace_bytes", "slide_mc_count", "slide_mc_emit",
]


class SlideError(RuntimeError):
    pass


def load():
    """Load libslide_b200.so; raises if it has not been built (python -m slide_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SlideError(
                "libslide_b200.so not found at %s -- build it with `python -m slide_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.slide_last_cuda_error.restype = ctypes.c_char_p
        lib.slide_launch_count.restype = ctypes.c_longlong
        lib.slide_program_arena.restype = ctypes.c_void_p
        lib.slide_program_weights.restype = ctypes.c_void_p
        lib.slide_program_arena.argtypes = [ctypes.c_void_p]
        lib.slide_program_weights.argtypes = [ctypes.c_void_p]
        lib.slide_program_destroy.argtypes = [ctypes.c_void_p]
        lib.slide_program_destroy.restype = None
        lib.slide_program_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        lib.slide_program_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.slide_program_capture.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_void_p]
        lib.slide_program_replay.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.slide_program_launches.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.slide_program_set_gemm_backend.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.slide_program_set_resident.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.slide_program_use_resident.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.slide_philox_normal_slice.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_ulonglong,
                                                  ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                                                  ctypes.c_void_p]
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = ERRORS.get(rc, str(rc))
        if rc == -2:
            msg += ": " + load().slide_last_cuda_error().decode()
        raise SlideError("%s failed: %s" % (what, msg))


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_of(t):
    """The current torch CUDA stream of t's device, as the cudaStream_t the C ABI expects."""
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def launch_count():
    return int(load().slide_launch_count())


def reset_launch_count():
    load().slide_reset_launch_count()
