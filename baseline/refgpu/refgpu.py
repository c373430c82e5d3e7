"""ctypes binding of baseline/refgpu/librefgpu.so (the reference's GPU design restated for timing; measurement tooling,
never imported by turbo_metrics_b200)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "librefgpu.so")
_lib = None


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", HERE])
    return SO_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("baseline/refgpu/librefgpu.so missing: run make -C baseline/refgpu (needs NPP)")
        L = C.CDLL(SO_PATH)
        L.refgpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
        L.refgpu_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.refgpu_info.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.refgpu_destroy.argtypes = [C.c_void_p]
        L.refgpu_destroy.restype = None
        _lib = L
    return _lib


class RefGpu:
    """One instance per (width, height, bit depth), like the reference's Ssimulacra2::new."""

    def __init__(self, w: int, h: int, bits: int):
        self._h = C.c_void_p()
        if lib().refgpu_create(C.byref(self._h), w, h, bits) != 0:
            raise RuntimeError("refgpu_create failed")

    def compute(self, ref, dis, pitch: int, coded_height: int):
        """ref / dis: torch CUDA tensors holding NV12 / P016 frames.  Synchronous; -> (score, norms[108])."""
        score = C.c_double()
        norms = np.zeros(108, dtype=np.float64)
        rc = lib().refgpu_compute(self._h, C.c_void_p(ref.data_ptr()), C.c_void_p(dis.data_ptr()), pitch, coded_height, C.byref(score),
                                  norms.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise RuntimeError("refgpu_compute failed")
        return score.value, norms

    def info(self):
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        lib().refgpu_info(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"graph_nodes": a.value, "kernel_nodes": b.value, "bytes": c.value}

    def close(self):
        if self._h:
            lib().refgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
