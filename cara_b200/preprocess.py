"""GPU input pipeline of the entry point.

The reference builds every batch on the CPU (image_classification/vtab.py:79-82, 4 DataLoader workers):

    transforms.Resize((224, 224), interpolation=3) -> transforms.ToTensor() -> transforms.Normalize(mean, std)

on a PIL image.  ``Resize(interpolation=3)`` is Pillow's two-pass antialiased bicubic (libImaging/Resample.c of the
pinned pillow, an un-vendored dependency -- restated here, pinned by tests against Pillow itself).  ``GpuPreprocessor``
keeps the JPEG decode on the CPU (uint8 HWC arrays), ships 3 bytes per pixel instead of 12 and runs resize + ToTensor
+ Normalize in ``cara_resize_normalize``.  Pillow's per-(in, out) integer tap tables are built on the host by
``resample_tables`` with the same double-precision operations in the same order, so the result is bit-identical.
"""
import math

import numpy as np
import torch

from . import _lib as L

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_tables(in_size, out_size):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the bicubic filter (support 2) over the whole
    input range: -> (bounds int32 [out,2] = (first tap, number of taps), kk int32 [out, ksize], ksize)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def resize_reference(img, out_h=224, out_w=224):
    """numpy restatement of Pillow's two passes on a uint8 [H,W,3] array (test infrastructure for the tables)."""
    def one_pass(a, bounds, kk, axis):
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((bounds.shape[0],) + a.shape[1:], dtype=np.uint8)
        for o in range(bounds.shape[0]):
            lo, n = bounds[o]
            s = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :n].astype(np.int64), a[lo:lo + n], axes=(0, 0))
            out[o] = np.clip(s >> PRECISION_BITS, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)
    H, W = img.shape[:2]
    if W != out_w:
        b, k, _ = resample_tables(W, out_w)
        img = one_pass(img, b, k, 1)
    if H != out_h:
        b, k, _ = resample_tables(H, out_h)
        img = one_pass(img, b, k, 0)
    return img


class GpuPreprocessor:
    """uint8 HWC images -> normalised fp32 [B,3,S,S] on the device; tables cached per input size."""

    def __init__(self, device, size=224, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        self.device = torch.device(device)
        self.size = size
        self.mean = np.asarray(mean, dtype=np.float32)
        self.std = np.asarray(std, dtype=np.float32)
        self._tables = {}

    def _table(self, n):
        if n == self.size:
            return None
        t = self._tables.get(n)
        if t is None:
            b, k, ks = resample_tables(n, self.size)
            t = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), ks)
            self._tables[n] = t
        return t

    def __call__(self, images, return_uint8=False):
        """``images``: uint8 tensor / array [B,H,W,3] (one size) or a list of [H,W,3] arrays of mixed sizes."""
        if isinstance(images, (list, tuple)):
            sizes = {}
            for i, im in enumerate(images):
                sizes.setdefault(tuple(im.shape[:2]), []).append(i)
            out = torch.empty((len(images), 3, self.size, self.size), device=self.device, dtype=torch.float32)
            for idx in sizes.values():
                out[torch.as_tensor(idx, device=self.device)] = self(np.stack([np.asarray(images[i]) for i in idx]))
            return out
        src = torch.as_tensor(images)
        if src.dtype != torch.uint8 or src.ndim != 4 or src.shape[3] != 3:
            raise ValueError("expected uint8 [B,H,W,3] images")
        if src.device != self.device:
            src = (src.pin_memory() if src.device.type == "cpu" else src).to(self.device, non_blocking=True)
        if not src.is_cuda:
            raise L.CaraLibraryError("GpuPreprocessor needs a CUDA device: there is no CPU fallback")
        src = src.contiguous()
        B, H, W, _ = src.shape
        S = self.size
        tx, ty = self._table(W), self._table(H)
        tmp = torch.empty((B, H, S, 3), device=self.device, dtype=torch.uint8) if tx is not None else None
        out = torch.empty((B, 3, S, S), device=self.device, dtype=torch.float32)
        u8 = torch.empty((B, S, S, 3), device=self.device, dtype=torch.uint8) if return_uint8 else None
        p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
        st = torch.cuda.current_stream(self.device).cuda_stream
        L.lib().cara_set_device(self.device.index or 0)
        L.check(L.lib().cara_resize_normalize(
            src.data_ptr(), B, H, W, p(tx and tx[0]), p(tx and tx[1]), tx[2] if tx else 0,
            p(ty and ty[0]), p(ty and ty[1]), ty[2] if ty else 0, p(tmp), out.data_ptr(), p(u8), S, S,
            self.mean.ctypes.data, self.std.ctypes.data, st), "cara_resize_normalize")
        return (out, u8) if return_uint8 else out
