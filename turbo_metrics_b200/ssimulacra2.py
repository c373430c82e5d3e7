"""Host-side mirror of the reference's metric op `Ssimulacra2`
(/root/reference/crates/ssimulacra2-cuda/src/lib.rs:27-291) over the C ABI.

Same call sequence as the reference -- construct once per (width, height), `compute` per pair,
`get_score` afterwards -- with the colour front-end (cuda-colorspace/src/lib.rs:33-169) folded
in: frames are handed over in the layout the decoder / NPP produces (NV12, P016, packed sRGB or
packed linear f32) instead of being converted to linear f32 by a separate op first.

PyTorch is only plumbing here: tensors own device memory, `torch.cuda.current_stream()` supplies
the stream handle.  All arithmetic happens in libssimu2_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from enum import IntEnum
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Config, Frame, Info, check


class PixelFormat(IntEnum):
    """turbo-metrics/src/lib.rs:125-130 `HwFrame` variants (+ the op's own linear input)."""
    NV12 = 0
    P016 = 1
    SRGB8 = 2
    SRGB16 = 3
    SRGBF32 = 4
    LINEARF32 = 5


class ColorMatrix(IntEnum):
    """cuda-colorspace/src/lib.rs:8-13."""
    BT709 = 0
    BT601_525 = 1
    BT601_625 = 2


@dataclass
class DeviceFrame:
    """A borrowed frame: device (or host, for compute_from_cpu) addresses + pitch in bytes.
    `keepalive` holds whatever owns the memory (a torch tensor, a numpy array)."""
    plane0: int
    plane1: int
    pitch: int
    keepalive: object = None
    nbytes: int = 0

    def c(self) -> Frame:
        f = Frame()
        f.plane[0] = self.plane0
        f.plane[1] = self.plane1
        f.pitch = self.pitch
        return f

    @staticmethod
    def yuv420(buf, pitch: int, coded_height: int) -> "DeviceFrame":
        """NVDEC layout (cudarse-video/src/dec.rs:299-366): Y at buf, CbCr at buf + pitch*coded_height.
        buf: 1-D uint8 torch tensor (device or pinned host) or numpy array."""
        base = buf.data_ptr() if hasattr(buf, "data_ptr") else buf.ctypes.data
        nbytes = buf.numel() * buf.element_size() if hasattr(buf, "numel") else buf.nbytes
        return DeviceFrame(base, base + pitch * coded_height, pitch, buf, nbytes)

    @staticmethod
    def packed(img, pitch: Optional[int] = None) -> "DeviceFrame":
        """Packed RGB image (H, W, 3) of u8 / u16 / f32: NPP `Image<_, C<3>>` layout."""
        base = img.data_ptr() if hasattr(img, "data_ptr") else img.ctypes.data
        if pitch is None:
            if hasattr(img, "stride"):
                pitch = img.stride(0) * img.element_size()
            else:
                pitch = img.strides[0]
        nbytes = img.numel() * img.element_size() if hasattr(img, "numel") else img.nbytes
        return DeviceFrame(base, 0, pitch, img, nbytes)


class Ssimulacra2:
    """`Ssimulacra2::new` (lib.rs:48-107): an instance is valid for one width x height x format."""

    def __init__(self, width: int, height: int, fmt: PixelFormat = PixelFormat.LINEARF32,
                 matrix: ColorMatrix = ColorMatrix.BT709, full_range: bool = False, device: int = 0,
                 batch: int = 0, ring: int = 0, pipeline: Optional[str] = None):
        """pipeline (development / tests): None = the library default ("hv": front-end, fused H+V kernel,
        finalize); "split" = four kernels with the H-pass planes in HBM (debug_read(what=1));
        "fh" = front-end fused with the previous batch's H pass + separate V pass.  Passed to the library
        through the SSIMU2_PIPELINE environment variable it reads in ssimu2_create."""
        self._h = C.c_void_p()
        cfg = Config(width, height, int(fmt), int(matrix), int(full_range), device, batch, ring)
        import os
        old = os.environ.get("SSIMU2_PIPELINE")
        if pipeline is not None:
            os.environ["SSIMU2_PIPELINE"] = pipeline
        try:
            check(_lib.lib().ssimu2_create(C.byref(self._h), C.byref(cfg)), "ssimu2_create")
        finally:
            if pipeline is not None:
                if old is None:
                    os.environ.pop("SSIMU2_PIPELINE", None)
                else:
                    os.environ["SSIMU2_PIPELINE"] = old
        self.width, self.height, self.format = width, height, PixelFormat(fmt)
        self._keep = {}

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().ssimu2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def mem_usage(self) -> int:
        """`Ssimulacra2::mem_usage` (lib.rs:110-138)."""
        n = C.c_size_t()
        check(_lib.lib().ssimu2_mem_usage(self._h, C.byref(n)), "ssimu2_mem_usage")
        return n.value

    def info(self) -> Info:
        i = Info()
        check(_lib.lib().ssimu2_get_info(self._h, C.byref(i)), "ssimu2_get_info")
        return i

    # -- scoring --------------------------------------------------------------------------
    @staticmethod
    def _stream_handle(stream) -> int:
        if stream is None:
            import torch
            return torch.cuda.current_stream().cuda_stream
        if hasattr(stream, "cuda_stream"):
            return stream.cuda_stream
        return int(stream)

    def compute(self, ref: DeviceFrame, dis: DeviceFrame, stream=None) -> int:
        """`Ssimulacra2::compute` (lib.rs:283-287): asynchronous; returns the pair's ticket."""
        t = C.c_uint64()
        a, b = ref.c(), dis.c()
        check(_lib.lib().ssimu2_submit(self._h, C.byref(a), C.byref(b), C.c_void_p(self._stream_handle(stream)),
                                       C.byref(t)), "ssimu2_submit")
        self._keep[t.value] = (ref, dis)
        return t.value

    def compute_batch(self, refs: Sequence[DeviceFrame], diss: Sequence[DeviceFrame], stream=None) -> range:
        n = len(refs)
        assert n == len(diss)
        A = (Frame * n)(*[f.c() for f in refs])
        B = (Frame * n)(*[f.c() for f in diss])
        t = C.c_uint64()
        check(_lib.lib().ssimu2_submit_batch(self._h, n, A, B, C.c_void_p(self._stream_handle(stream)), C.byref(t)),
              "ssimu2_submit_batch")
        for i in range(n):
            self._keep[t.value + i] = (refs[i], diss[i])
        return range(t.value, t.value + n)

    def compute_from_cpu(self, ref: DeviceFrame, dis: DeviceFrame) -> int:
        """`Ssimulacra2::compute_from_cpu_srgb_sync` (lib.rs:232-250) without the sync: frames live in
        host memory; the library stages them on the device."""
        t = C.c_uint64()
        a, b = ref.c(), dis.c()
        assert ref.nbytes and ref.nbytes == dis.nbytes
        check(_lib.lib().ssimu2_submit_host(self._h, C.byref(a), C.byref(b), ref.nbytes, C.byref(t)),
              "ssimu2_submit_host")
        self._keep[t.value] = (ref, dis)
        return t.value

    def compute_from_cpu_batch(self, refs: Sequence[DeviceFrame], diss: Sequence[DeviceFrame]) -> range:
        """n host pairs in one call (`ssimu2_submit_host_batch`)."""
        n = len(refs)
        assert n == len(diss) and n > 0 and all(f.nbytes == refs[0].nbytes for f in list(refs) + list(diss))
        A = (Frame * n)(*[f.c() for f in refs])
        B = (Frame * n)(*[f.c() for f in diss])
        t = C.c_uint64()
        check(_lib.lib().ssimu2_submit_host_batch(self._h, n, A, B, refs[0].nbytes, C.byref(t)), "ssimu2_submit_host_batch")
        for i in range(n):
            self._keep[t.value + i] = (refs[i], diss[i])
        return range(t.value, t.value + n)

    def flush(self):
        check(_lib.lib().ssimu2_flush(self._h), "ssimu2_flush")

    def get_score(self, ticket: Optional[int] = None) -> float:
        """`Ssimulacra2::get_score` (lib.rs:289-291); defaults to the last submitted pair."""
        if ticket is None:
            ticket = max(self._keep) if self._keep else 0
        s = C.c_double()
        check(_lib.lib().ssimu2_get_score(self._h, ticket, C.byref(s)), "ssimu2_get_score")
        self._release(ticket)
        return s.value

    def get_scores(self, tickets: range) -> np.ndarray:
        """Scores of consecutive tickets (`ssimu2_get_scores`): the per-frame score stream in submission order."""
        n = len(tickets)
        out = np.zeros(n, np.float64)
        if n:
            assert tickets.step == 1
            check(_lib.lib().ssimu2_get_scores(self._h, tickets.start, n, out.ctypes.data_as(C.POINTER(C.c_double))), "ssimu2_get_scores")
            self._release(tickets[-1])
        return out

    def get_norms(self, ticket: int) -> np.ndarray:
        out = np.zeros(108, np.float64)
        check(_lib.lib().ssimu2_get_norms(self._h, ticket, out.ctypes.data_as(C.POINTER(C.c_double))),
              "ssimu2_get_norms")
        return out

    def compute_sync(self, ref: DeviceFrame, dis: DeviceFrame, stream=None) -> float:
        """`Ssimulacra2::compute_sync` (lib.rs:271-279)."""
        return self.get_score(self.compute(ref, dis, stream))

    def compute_from_cpu_sync(self, ref: DeviceFrame, dis: DeviceFrame) -> float:
        return self.get_score(self.compute_from_cpu(ref, dis))

    def _release(self, upto: int):
        for k in [k for k in self._keep if k <= upto]:
            del self._keep[k]

    # -- device-side results / introspection ---------------------------------------------------
    def scores_device(self):
        p, cap = C.c_uint64(), C.c_uint64()
        check(_lib.lib().ssimu2_scores_device(self._h, C.byref(p), C.byref(cap)), "ssimu2_scores_device")
        return p.value, cap.value

    def debug_read(self, ticket: int, what: int, scale: int) -> np.ndarray:
        i = self.info()
        w, h = i.width[scale], i.height[scale]
        planes = 6 if what == 0 else 15
        out = np.zeros((planes, h, w), np.float32)
        check(_lib.lib().ssimu2_debug_read(self._h, ticket, what, scale, out.ctypes.data_as(C.POINTER(C.c_float)),
                                           out.size), "ssimu2_debug_read")
        return out

    def kernel_ms(self, reset: bool = False):
        """-> (ms_total[4] for frontend/hpass/vpass/finalize, batches, pairs) since the last reset."""
        ms = (C.c_double * 4)()
        b, p = C.c_uint64(), C.c_uint64()
        check(_lib.lib().ssimu2_kernel_ms(self._h, ms, C.byref(b), C.byref(p), int(reset)), "ssimu2_kernel_ms")
        return list(ms), b.value, p.value

    def last_batch_ms(self):
        ms = (C.c_float * 4)()
        check(_lib.lib().ssimu2_last_batch_ms(self._h, ms), "ssimu2_last_batch_ms")
        return list(ms)
