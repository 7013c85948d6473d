"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the reference's chest-X-ray image preprocessing - SURVEY.md 8f row 2.

Restates, in numpy, what ``demo.py`` / ``ReportDataset.py`` do to one grey-scale image before ``forward_image``:

  remap_to_uint8                      demo.py:173-203   (float64: x -= min; x /= max; x *= 255; truncate to uint8)
  Image.fromarray(..).convert("L")    demo.py:218
  Resize(512)  (smaller edge, PIL bilinear with antialiasing)   ReportDataset.py:104, torchvision 0.14 ``F.resize`` on PIL
  CenterCrop(448)                     ReportDataset.py:104  (torchvision: top = int(round((H - 448) / 2.)))
  ToTensor()                          uint8 / 255 -> float32
  ExpandChannels()                    ReportDataset.py:80-94  (the channel three times)

The resize is third-party code absent from /root/reference: Pillow (unpinned by the reference; 12.2.0 in this image),
``src/libImaging/Resample.c``: ``precompute_coeffs`` (double), ``normalize_coeffs_8bpc`` (PRECISION_BITS = 32 - 8 - 2),
``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc`` with an 8-bit intermediate image - restated here
with integer arithmetic so that it is bit-exact.  Pinned against Pillow + torchvision themselves by
``oracle/make_golden_preprocess.py`` (fixtures in tests/golden/preprocess_*.npz) and, where Pillow is importable, live in
``tests/test_preprocess_oracle.py``.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def remap_to_uint8(array: np.ndarray) -> np.ndarray:
    """demo.py:184,200-203 (percentiles=None branch)."""
    a = array.astype(float)
    a -= a.min()
    a /= a.max()
    a *= 255
    return a.astype(np.uint8)


def resized_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision ``_compute_resized_output_size`` for an int size: smaller edge -> size, the other int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)     # (new_h, new_w)


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c ``precompute_coeffs`` for the bilinear filter (support 1.0) over the whole axis, then
    ``normalize_coeffs_8bpc``.  Returns (bounds [out,2] int32 = (xmin, count), kk [out, ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        n = xmax - xmin
        k = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(n):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0 else v
            wgt = 1.0 - v if v < 1.0 else 0.0
            k[x] = wgt
            ww += wgt
        if ww != 0.0:
            k[:n] = k[:n] / ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, n)
    return bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit pass of Resample.c along ``axis`` (1 = horizontal)."""
    src = img if axis == 1 else img.T
    bounds, kk = precompute_coeffs(src.shape[1], out_size)
    out = np.empty((src.shape[0], out_size), np.uint8)
    s64 = src.astype(np.int64)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (s64[:, xmin:xmin + n] * kk[xx, :n].astype(np.int64)).sum(1) + (1 << (PRECISION_BITS - 1))
        out[:, xx] = _clip8(acc)
    return out if axis == 1 else out.T


def pil_resize_bilinear(img_u8: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """``Image.resize((new_w, new_h), BILINEAR)`` on a mode-L image: horizontal pass, then vertical pass on the 8-bit
    intermediate (ImagingResample: each pass only if that axis changes size)."""
    out = img_u8
    if new_w != img_u8.shape[1]:
        out = resample_axis(out, new_w, 1)
    if new_h != img_u8.shape[0]:
        out = resample_axis(out, new_h, 0)
    return out


def center_crop_offsets(h: int, w: int, crop: int) -> Tuple[int, int]:
    """torchvision ``center_crop`` (image at least as large as the crop on both axes)."""
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def preprocess(array: np.ndarray, resize: int = 512, crop: int = 448) -> np.ndarray:
    """[H,W] grey image (any numeric dtype) -> float32 [3,crop,crop] in [0,1] exactly as the reference pipeline."""
    assert array.ndim == 2
    u8 = remap_to_uint8(array)
    nh, nw = resized_size(u8.shape[0], u8.shape[1], resize)
    r = pil_resize_bilinear(u8, nh, nw)
    top, left = center_crop_offsets(nh, nw, crop)
    c = r[top:top + crop, left:left + crop]
    t = c.astype(np.float32) / np.float32(255.0)
    return np.repeat(t[None], 3, axis=0)
